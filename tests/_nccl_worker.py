"""Worker of test_gpu_two_rank_nccl_sharded_solve (torchrun, one process per GPU, nccl backend): the Robertson
ensemble of BASELINE.json configs[2], sharded over the ranks by dae_scipy_b200.dist.solve_ivp_sharded, gathered onto
rank 0 over NCCL and compared with the same ensemble integrated by rank 0 alone."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # the repository root
import dae_scipy_b200 as br
from dae_scipy_b200 import dist as brdist

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
B = int(os.environ.get("BRDAE_TEST_B", "200001"))            # odd: shards differ by one system
ens = br.ensembles.robertson_ensemble(B, t_end=1e3, n_t_eval=8)
out = brdist.solve_ivp_sharded(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"],
                               gather_dst=0, device=dev, **ens["options"])
lo, hi = out["shard"]
assert (lo, hi) == brdist.shard_bounds(B, world, rank) and out["local"].status.shape[0] == hi - lo
assert int(out["status_histogram"].sum()) == B               # every system of the WHOLE ensemble is accounted for
if rank == 0:
    full = br.solve_ivp_batched(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], device=dev,
                                **ens["options"])
    torch.cuda.synchronize()
    assert out["y"].shape == (B, 3, 8)
    assert torch.equal(out["y"], full.y) and torch.equal(out["status"], full.status)
    assert torch.equal(out["counters"], full.counters.t()) and torch.equal(out["y_final"], full.y_final)
    assert torch.equal(out["n_eval"], full.n_eval) and torch.equal(out["t_final"], full.t_final)
    assert torch.equal(out["counters_sum"], full.counters.sum(dim=1, dtype=torch.int64))
    assert out["status_histogram"].tolist() == [B, 0, 0, 0, 0]
    print(f"rank0 ok: {B} Robertson systems over {world} GPUs, gathered over NCCL, bit-identical to one GPU")
else:
    assert "y" not in out
dist.barrier()
dist.destroy_process_group()
