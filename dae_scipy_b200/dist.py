"""Multi-GPU sharding of an ensemble: one process per GPU, torch.distributed for the plumbing.

The systems of an ensemble are independent: there is no exchange step during
the integration, so the ensemble is partitioned across ranks and every rank runs the
single-GPU path on its shard.  Collectives are used only after the kernels: an `all_reduce`
of the summed counters / status histogram, and — on request — a `gather` of the per-system
results onto rank 0 (NCCL over NVLink when the backend is nccl; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(B: int, world_size: int, rank: int):
    """Contiguous, balanced partition of range(B): the first B % world_size ranks get one more."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_ensemble(ens: dict, world_size: int, rank: int) -> dict:
    """Slice the batch-sized entries (y0, params) of an ensemble description for this rank."""
    B = ens["y0"].shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    out = dict(ens)
    out["y0"] = ens["y0"][lo:hi]
    out["params"] = None if ens["params"] is None else ens["params"][lo:hi]
    out["shard"] = (lo, hi)
    return out


def reduce_statistics(counters_sum, status_hist, dist=None, device=None):
    """Sum the per-rank statistics vectors over all ranks (all_reduce SUM).  `counters_sum` is the
    int64 [9] sum of the per-system counters of the local shard, `status_hist` an int64 [5] vector
    (finished, too-small-step, lu-failed, nmax-step, ended-by-a-terminal-event)."""
    import torch
    vec = torch.cat([torch.as_tensor(counters_sum, dtype=torch.int64).reshape(-1),
                     torch.as_tensor(status_hist, dtype=torch.int64).reshape(-1)])
    if device is not None:
        vec = vec.to(device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec[:9].clone(), vec[9:].clone()


def gather_results(local: dict, B_total: int, dist, dst: int = 0):
    """Gather per-system result tensors of all shards onto rank `dst`, restoring ensemble order.

    `local` maps name -> tensor whose LAST axis is the local system axis (kernel SoA layout,
    e.g. y_eval [T, n, B_local], status [B_local]).  Shards may differ by one system, so each is
    padded to the largest shard before `dist.gather`.  Returns the dict of gathered tensors on
    `dst` and None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(B_total, world, r) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    out = {} if rank == dst else None
    for name, ten in local.items():
        pad = maxlen - ten.shape[-1]
        send = ten if pad == 0 else torch.nn.functional.pad(ten, (0, pad))
        send = send.contiguous()
        bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send, gather_list=bufs, dst=dst)
        if rank == dst:
            out[name] = torch.cat([b[..., :hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=-1)
    return out


def status_histogram(status):
    """[finished, too-small-step, lu-failed, nmax-step, terminal-event] counts of a status vector (BRDAE_STATUS_* 0, -1, -2,
    2, 1): the bins sum to the number of systems."""
    import torch
    s = status if isinstance(status, torch.Tensor) else torch.as_tensor(np.asarray(status))
    return torch.stack([(s == 0).sum(), (s == -1).sum(), (s == -2).sum(), (s == 2).sum(), (s == 1).sum()]).to(torch.int64)


def solve_ivp_sharded(problem, t_span, y0, params=None, *, dist=None, gather_dst=None, device=None, **kwargs):
    """The whole ensemble (`y0` [B, n], `params` [B, p]: host arrays every rank holds, or can slice) integrated by all
    ranks of the default process group, one GPU per process: this rank solves the contiguous shard
    `shard_bounds(B, world, rank)` with `solve_ivp_batched` (no data-path collective), the summed counters and the
    status histogram are all-reduced, and with `gather_dst=r` the per-system results are gathered onto rank r in
    ensemble order (NCCL over NVLink with the nccl backend).

    Returns a dict: `local` (this rank's BatchedOdeResult), `shard` (lo, hi), `counters_sum` [9] and
    `status_histogram` [5] over the WHOLE ensemble (on every rank; the bins sum to B), and on `gather_dst` also `y` [B, n, T],
    `status`, `n_eval`, `t_final`, `y_final` [B, n], `counters` [B, 9] as device tensors (None elsewhere)."""
    import torch
    from .api import solve_ivp_batched
    if dist is None:
        import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    world, rank = (dist.get_world_size(), dist.get_rank()) if multi else (1, 0)
    B = int(y0.shape[0])
    lo, hi = shard_bounds(B, world, rank)
    res = solve_ivp_batched(problem, t_span, y0[lo:hi], None if params is None else params[lo:hi], device=device, **kwargs)
    dev = res.status.device
    csum, hist = reduce_statistics(res.counters.sum(dim=1, dtype=torch.int64), status_histogram(res.status),
                                   dist if multi else None, dev)
    out = {"local": res, "shard": (lo, hi), "counters_sum": csum, "status_histogram": hist}
    if gather_dst is not None:
        local = {"y_eval": res._y_eval, "status": res.status, "n_eval": res.n_eval, "t_final": res.t_final,
                 "y_final": res.y_final.t().contiguous(), "counters": res.counters}
        g = gather_results(local, B, dist, dst=gather_dst) if multi else local
        if g is not None:
            out.update(y=g["y_eval"].permute(2, 1, 0), status=g["status"], n_eval=g["n_eval"], t_final=g["t_final"],
                       y_final=g["y_final"].t(), counters=g["counters"].t())
    return out
