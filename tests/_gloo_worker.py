"""Worker of test_two_rank_gloo_gather_and_reduce (gloo backend, CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import dae_scipy_b200 as br
from dae_scipy_b200 import dist as brdist
from _util import run_c_oracle

dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
B = 37                                           # odd on purpose: shards of 19 and 18
ens = br.ensembles.pendulum_ensemble(B, t_end=0.5, n_t_eval=6)
mine = brdist.shard_ensemble(ens, world, rank)
out = run_c_oracle(mine)                         # stand-in for the GPU kernel on this shard
local = {"y_eval": torch.as_tensor(out["y"]).permute(2, 1, 0).contiguous(),      # [T, n, B_local]
         "status": torch.as_tensor(out["status"]), "counters": torch.as_tensor(out["counters"]).t().contiguous()}
csum, hist = brdist.reduce_statistics(local["counters"].sum(dim=1, dtype=torch.int64),
                                      brdist.status_histogram(local["status"]), dist)
gathered = brdist.gather_results(local, B, dist, dst=0)
if rank == 0:
    full = run_c_oracle(ens)
    assert gathered["y_eval"].shape == (6, 5, B)
    assert np.array_equal(gathered["y_eval"].permute(2, 1, 0).numpy(), full["y"])
    assert np.array_equal(gathered["status"].numpy(), full["status"])
    assert np.array_equal(gathered["counters"].t().numpy(), full["counters"])
    assert np.array_equal(csum.numpy(), full["counters"].sum(axis=0))
    assert hist.tolist() == [B, 0, 0, 0, 0]
    print("rank0 ok")
else:
    assert gathered is None
dist.barrier()
dist.destroy_process_group()
