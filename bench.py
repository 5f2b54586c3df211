#!/usr/bin/env python
"""bench.py — DAE trajectories/s to t_end (FP64) for the batched RadauDAE hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the UNMODIFIED reference (scipyDAE RadauDAE
                                                              # from oracle/_ref) on all host cores; its numpy port if
                                                              # the reference did not travel with the snapshot

A "step" is one pass of the hot path over one batch: the whole ensemble integrated from t0 to
t_end.  Default workload = BASELINE.json configs[1]: index-3 pendulum, 1,048,576 varied initial
conditions, thread-per-system, one B200 (rtol = atol = 1e-6, first_step = 1e-2, t_end = 2,
21 t_eval points; SURVEY 8d "C2 primary").  With N > 1 every rank integrates its own 1M-system
ensemble (weak scaling, different seeds); `--scaling strong` splits ONE ensemble of the workload's size over the ranks
(BASELINE configs[2]: `--workload robertson_4M --scaling strong`).  The collective inside the timed region is the
all_reduce of the summed counters; a second timed region adds the NCCL gather of every result onto rank 0
(`value_with_gather`).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DAE trajectories/sec to t_end (FP64)"
UNIT = "trajectories/s"

WORKLOADS = {
    # name: (ensemble generator key, kwargs, default B)
    "pendulum3_1M_tend2": ("pendulum", dict(t_end=2.0, n_t_eval=21), 1 << 20),
    "pendulum3_1M_tend10": ("pendulum", dict(t_end=10.0, n_t_eval=21), 1 << 20),
    "robertson_4M": ("robertson", dict(), 1 << 22),
    "transistor_256k": ("transistor", dict(t_end=0.2), 1 << 18),
    "npendulum_n20_16k": ("npendulum", dict(n_nodes=20), 1 << 14),
    "npendulum_n50_16k": ("npendulum", dict(n_nodes=50), 1 << 14),
    "npendulum_n100_16k": ("npendulum", dict(n_nodes=100), 1 << 14),
}


# DRAM traffic of one launch, per system: dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full` capture of the
# workload's kernel divided by the systems of the profiled launch; each entry names the committed summary it comes from.
NCU_DRAM_BYTES_PER_SYSTEM = {
    "pendulum3_1M_tend2": ((160.46e6 + 333.03e6) / 262144, "profiles/r02_k1_ncu_summary.md (v9, 262,144 systems)"),
    "pendulum3_1M_tend10": ((160.46e6 + 333.03e6) / 262144, "profiles/r02_k1_ncu_summary.md (same kernel and output grid as t_end = 2)"),
    "robertson_4M": ((78.3e6 + 448.8e6) / 1048576, "profiles/r02_k1_ncu_summary.md (Robertson, 1,048,576 systems)"),
    "transistor_256k": ((10.4e6 + 8.5e6) / 65536, "profiles/r02_k1_ncu_summary.md (transistor, 65,536 systems: part of that launch's output was still in L2 at kernel end)"),
    "npendulum_n20_16k": ((575.0e9 + 187.6e9) / 16384, "profiles/r02_k4_ncu_summary.md (K4, final capture of round 2, n = 20, 16,384 ropes)"),
}


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms
def _reference_kind():
    """"reference": the unmodified scipyDAE.radauDAE.RadauDAE is importable (oracle/_ref travels with the snapshot, /root/reference
    in the build container); "port": only its numpy restatement (oracle/radau_numpy.py, bit-identical where both run) is."""
    try:
        from oracle import reference_runner
        return "reference" if reference_runner.reference_location() is not None else "port"
    except Exception:
        return "port"


def _cpu_worker(args):
    """Integrate a slice of the ensemble on one core: with the reference itself (kind "reference") or its numpy port."""
    import contextlib
    import io
    import warnings
    warnings.filterwarnings("ignore")
    kind, ens = args
    # one BLAS thread per process: the pool already uses every core, and the LAPACK calls of the chain problems
    # (N = 100..500) would otherwise start a full OpenBLAS thread team in each forked worker (measured: 50x slower)
    # The limit only reaches libraries that are loaded when it is set: scipy brings its own OpenBLAS, so import it first
    # (a first call that loaded scipy under the limit ran 16 x 16 spinning threads: 240 s instead of 0.3 s per rope).
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    import scipy.linalg  # noqa: F401
    import scipy.sparse.linalg  # noqa: F401
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)
    except Exception:
        limit = contextlib.nullcontext()
    with limit, contextlib.redirect_stdout(io.StringIO()):
        if kind == "reference":
            from oracle import reference_runner
            out = reference_runner.run_reference(ens)
            return ens["y0"].shape[0], int((out["status"] == 0).sum())
        from oracle import radau_numpy as orc, problems_numpy as prob
        opts = dict(ens["options"])
        opts.pop("mass_matrix", None)
        var_index = opts.pop("var_index", None)
        done = 0
        for i in range(ens["y0"].shape[0]):
            p = ens["params"][i] if ens["params"] is not None else ()
            fun, jac, mass, vi = prob.build(ens["problem"], p, ens["y0"].shape[1])
            r = orc.solve(fun, ens["t_span"], ens["y0"][i], jac=jac, mass_matrix=mass,
                          var_index=None if var_index is None else np.asarray(var_index, dtype=float),
                          t_eval=ens["t_eval"], **opts)
            done += int(r.status == 0)
    return ens["y0"].shape[0], done


def cpu_reference_rate(ens, n_systems, cores, kind):
    """traj/s of the reference path (`kind`) over `cores` processes, on the first `n_systems` systems of the ensemble."""
    import multiprocessing as mp
    chunks = [c for c in np.array_split(np.arange(n_systems), cores) if c.size]
    sub = lambda c: dict(ens, y0=ens["y0"][c], params=None if ens["params"] is None else ens["params"][c])
    jobs = [(kind, sub(c)) for c in chunks]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
        pool.map(_cpu_worker, [(kind, sub(np.arange(1)))] * len(jobs))      # warm import
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    return n / dt, dt, n


# about what the reference needs for one system on one core (measured, 1 BLAS thread): sizes the bounded CPU samples
_REF_SECONDS_PER_SYSTEM = {"pendulum3_1M_tend2": 0.12, "robertson_4M": 0.1, "transistor_256k": 0.7, "pendulum3_1M_tend10": 0.7,
                           "npendulum_n20_16k": 0.8, "npendulum_n50_16k": 3.0, "npendulum_n100_16k": 20.0}


def _kind_text(kind):
    return ("the unmodified reference (scipyDAE.radauDAE.RadauDAE installed under oracle/_ref), stepped like scipy's solve_ivp"
            if kind == "reference" else "numpy restatement of the reference (oracle/radau_numpy.py; the reference itself is not importable here)")


def run_reference_arm(args, ens, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = _reference_kind()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # a bounded sample per step: at most 64 systems per core, fewer where one system costs the reference seconds (chains)
    per_step = max(cores, min(args.ref_sample, cores * min(64, max(1, int(25.0 / _REF_SECONDS_PER_SYSTEM.get(wl_name, 0.2))))))
    for _ in range(args.warmup):
        cpu_reference_rate(ens, min(per_step, 2 * cores), cores, kind)
    total_n, total_t = 0, 0.0
    for _ in range(args.steps):
        rate, dt, n = cpu_reference_rate(ens, per_step, cores, kind)
        total_n += n; total_t += dt
    value = total_n / total_t
    sample = f"first {per_step} systems of the {wl_name} ensemble per step, {_kind_text(kind)}, Pool({cores}), 1 BLAS thread per process"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl_name, "systems_per_step": per_step, "parallelism": f"cpu{cores}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="brdae", choices=["brdae", "reference"])
    ap.add_argument("--workload", default="pendulum3_1M_tend2", choices=sorted(WORKLOADS))
    ap.add_argument("--systems", type=int, default=0, help="override the ensemble size (per GPU with weak scaling, in total with strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns an ensemble of the workload's size; strong: ONE such ensemble split over the ranks")
    ap.add_argument("--ref-sample", type=int, default=4096, help="systems per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=0, help="systems of the in-run cpu_baseline (0 = 384 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the second timed region (solve + NCCL gather onto rank 0)")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the rank to the NUMA node of its GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "brdae" else args.warmup

    gen_key, gen_kw, B_default = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import dae_scipy_b200 as br
    if args.impl == "reference":
        ens = br.ensembles.ENSEMBLES[gen_key](max(args.ref_sample, 64 * (os.cpu_count() or 1)), **gen_kw)
        run_reference_arm(args, ens, args.workload)
        return

    import torch
    import torch.distributed as dist
    from dae_scipy_b200 import _abi, flops, numa
    from dae_scipy_b200.dist import gather_results, reduce_statistics, shard_bounds, status_histogram

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    # keep this rank's threads and its page-locked result buffers on the NUMA node of its GPU (multi-GPU end-to-end rate)
    placement = None if args.no_numa else numa.bind_to_device_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B_total = args.systems or B_default
    seed0 = {"pendulum": 0, "robertson": 1, "transistor": 2, "npendulum": 3}[gen_key]
    if args.scaling == "strong":
        # one ensemble, contiguous balanced shards (dist.shard_bounds); the first systems do not depend on the size drawn
        lo, hi = shard_bounds(B_total, world, rank)
        full = br.ensembles.ENSEMBLES[gen_key](B_total, seed=seed0, **gen_kw)
        ens = dict(full, y0=full["y0"][lo:hi], params=None if full["params"] is None else full["params"][lo:hi])
        B = hi - lo
        systems_all = B_total
    else:
        # weak scaling: every rank owns its own B-system ensemble (seed offset by rank)
        B = B_total
        ens = br.ensembles.ENSEMBLES[gen_key](B, seed=seed0 + 1000 * rank, **gen_kw)
        systems_all = B * world
    prob = br.get_problem(ens["problem"])
    n, T, npar = ens["y0"].shape[1], len(ens["t_eval"]), prob.n_params

    # inputs resident in HBM before the timed region
    y0_d = torch.as_tensor(ens["y0"]).to(dev)
    par_d = torch.as_tensor(ens["params"]).to(dev) if ens["params"] is not None else None
    stream = torch.cuda.current_stream(dev)

    def one_pass():
        return br.solve_ivp_batched(ens["problem"], ens["t_span"], y0_d, par_d, t_eval=ens["t_eval"],
                                    device=dev, **ens["options"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = None
    for _ in range(args.warmup):       # warm-up steps are whole steps: the statistics all_reduce (NCCL set-up) included
        res = one_pass()
        csum = res.counters.sum(dim=1, dtype=torch.int64)
        reduce_statistics(csum, status_histogram(res.status), dist if world > 1 else None, dev)
    barrier()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank)
    clk.__enter__()
    barrier()
    e_start.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        res = one_pass()
        ev[k][1].record(stream)
        csum = res.counters.sum(dim=1, dtype=torch.int64)
        csum, hist = reduce_statistics(csum, status_histogram(res.status), dist if world > 1 else None, dev)
    e_stop.record(stream)
    barrier()
    total_ms = max_over_ranks(e_start.elapsed_time(e_stop))
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    value = systems_all * args.steps / (total_ms * 1e-3)

    # second timed region: the same steps plus the gather of every per-system result (y_eval, y_final, t_final, status,
    # n_eval, counters) onto rank 0 over NCCL (dist.gather_results: NVLink / NVSwitch), which is what north_star uses NCCL for
    with_gather = None
    if not args.no_gather:
        def gathered_pass():
            r = one_pass()
            cs, _ = reduce_statistics(r.counters.sum(dim=1, dtype=torch.int64), status_histogram(r.status), dist if world > 1 else None, dev)
            if world > 1:
                local = {"y_eval": r._y_eval, "status": r.status, "n_eval": r.n_eval, "t_final": r.t_final,
                         "y_final": r.y_final.t().contiguous(), "counters": r.counters}
                if args.scaling == "strong":
                    return gather_results(local, systems_all, dist, dst=0)
                return gather_results(local, systems_all, dist, dst=0)       # equal shards under weak scaling: same call
            return r
        g = gathered_pass()                                                   # warm-up (gather buffers, NCCL channels)
        del g
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(args.steps):
            g = gathered_pass()
            del g
        g1.record(stream)
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1))
        gather_bytes = (world - 1) * (8 * B * n * T + 8 * B * n + 8 * B + 4 * B * 2 + 4 * B * 9) if world > 1 else 0
        with_gather = {"value": systems_all * args.steps / (gms * 1e-3), "unit": UNIT, "ms_per_step": gms / args.steps,
                       "gathered_bytes_per_step_into_rank0": gather_bytes,
                       "what": "solve + all_reduce of the statistics + dist.gather of y_eval, y_final, t_final, status, n_eval, counters onto rank 0 (NCCL)"}

    # measured FP64 roofline denominator: burst DFMA probe on this device, run AFTER the timed regions (60 ms at full FP64 power
    # right before them is a disturbance the measurement does not need)
    barrier()
    fp64_peak = br.fp64_peak_tflops(60.0)
    barrier()

    # roofline of the dominant (only) kernel: algorithmic FLOPs of one launch / its duration
    out_counts = res.counters.sum(dim=1, dtype=torch.int64).cpu().numpy()
    n_out = int(res.n_eval.sum().item())
    chain = gen_key == "npendulum" and os.environ.get("BRDAE_NPENDULUM_KERNEL", "") != "k2"
    fl = flops.algorithmic_flops(ens["problem"], n, out_counts, n_out, chain_block=chain)
    kms = float(np.mean(kern_ms))
    achieved = fl / (kms * 1e-3) / 1e12
    abytes = flops.algorithmic_bytes(n, npar, T, B)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback"
    hbm_gbs = abytes / (kms * 1e-3) / 1e9
    traffic, traffic_src = NCU_DRAM_BYTES_PER_SYSTEM.get(args.workload, (None, None))
    kernel_name = {"pendulum": "k1_kernel<Pendulum<3>,MassDiag01>", "npendulum": "k4_kernel (lane-per-rope block elimination)" if chain else "k2_kernel"}.get(gen_key, f"k1_kernel<{gen_key}>")
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                "traffic": None if traffic is None else traffic * B, "traffic_source": traffic_src,
                "peak_source": "measured in this run: dependent-chain DFMA probe (brdae_fp64_peak_probe)",
                "kernel": kernel_name, "kernel_ms": kms, "flops_per_launch": fl, "flops_per_trajectory": fl / B,
                "hbm": {"algorithmic_bytes": abytes, "achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak,
                        "frac": hbm_gbs / hbm_peak, "peak_source": hbm_src}}
    if chain:
        # The chain kernel is bound by HBM, not by the FP64 pipe: it streams the state of every rope (99 doubles per node) through
        # shared memory once per sweep (TMA bulk copies, csrc/chain_k4.cuh: Stager).  Algorithmic bytes = what the block-elimination
        # sweeps of this run stage and store per rope (flops.chain_state_bytes_per_node_sweep x the run's own counters), i.e. the
        # traffic if every lane of a warp were busy in every sweep; `traffic` is the DRAM traffic ncu measured for this workload.
        per = flops.chain_state_bytes_per_node_sweep()
        nodes = n // 5
        attempts = float(out_counts[4] + out_counts[5] + out_counts[6])         # accepted + rejected + failed: one set-up sweep each
        state_bytes = nodes * (float(out_counts[7]) * per["newton"] + float(out_counts[2]) * per["factor"] + float(out_counts[8]) * per["error_estimate"]
                               + attempts * per["setup"] + float(out_counts[4]) * per["accept"]) + abytes
        gbs = state_bytes / (kms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": None if traffic is None else traffic * B, "traffic_source": traffic_src, "peak_source": hbm_src,
                    "kernel": kernel_name, "kernel_ms": kms, "bytes_per_launch": state_bytes, "bytes_per_trajectory": state_bytes / B,
                    "fp64": {"achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                             "flops_per_launch": fl, "peak_source": "measured in this run: dependent-chain DFMA probe (brdae_fp64_peak_probe)"},
                    "note": "rope state staged into shared memory and stored by the sweeps, from the run's counters; see profiles/r02_k4_ncu_summary.md"}

    statuses = hist.cpu().numpy().tolist()

    # end-to-end through the C ABI with HOST buffers (pinned): H2D of inputs + D2H of every result
    e2e = None
    if not args.no_e2e:
        # the caller's arrays are ordinary numpy; staging and result buffers are page-locked and reused
        bufs = br.host_buffers(n, B, T, npar, pinned=True)
        for _ in range(2):
            br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"],
                                      buffers=bufs, **ens["options"])
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"],
                                            buffers=bufs, **ens["options"])
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = 8 * B * (n + npar) + 8 * T
        d2h = 8 * B * n * T + 8 * B + 8 * B * n + 4 * B * 2 + 4 * B * 9
        e2e = {"value": systems_all * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "host_write_gbs_all_ranks": d2h * world * args.steps / dt / 1e9,
               "path": "solve_ivp_batched_host -> brdae_solve_host_rows (C ABI, host buffers, row per system): argument "
                       "validation, stream-ordered device alloc, H2D of y0/params, device permutation to SoA, kernel, device "
                       "permutation of the results, D2H of y[B,n,T], y_final, t_final, status, n_eval, counters into reused "
                       "page-locked buffers and the final synchronisation are all inside the timed region",
               "all_finished": bool((out["status"] == 0).all())}

        # what bounds the end-to-end rate with several ranks on one host: the plain device-to-host copy rate of a buffer of the
        # result's size into page-locked memory, all ranks copying at the same time (no kernel, no host work)
        probe_d = probe_h = None
        try:
            probe_d = torch.empty(d2h // 8, dtype=torch.float64, device=dev)
            probe_h = torch.empty(d2h // 8, dtype=torch.float64, pin_memory=True)
            ok = 1.0
        except Exception as exc:
            ok = 0.0
            e2e["d2h_probe_error"] = repr(exc)
        if -max_over_ranks(-ok) > 0.5:           # every rank got its buffers (a rank that did not must not leave the others in a barrier)
            probe_h.copy_(probe_d, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                probe_h.copy_(probe_d, non_blocking=True)
            torch.cuda.synchronize(dev)
            dtp = max_over_ranks(time.perf_counter() - t0)
            e2e["d2h_copy_probe_gbs_all_ranks"] = 3 * d2h * world / dtp / 1e9
            e2e["d2h_bound_on_e2e"] = systems_all / (d2h * world / (e2e["d2h_copy_probe_gbs_all_ranks"] * 1e9))
        else:
            e2e["d2h_copy_probe_gbs_all_ranks"] = None
        del probe_d, probe_h

    clk.__exit__(None, None, None)

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = _reference_kind()
        # about 10-30 s of CPU work per workload (the reference integrates roughly 8 / 10 / 1.5 / 1.5 systems per second and core
        # on the pendulum t_end = 2 / Robertson / transistor / pendulum t_end = 10 configurations, 1 rope per second at n = 20)
        per_core = {"pendulum3_1M_tend2": 160, "robertson_4M": 160, "transistor_256k": 24, "pendulum3_1M_tend10": 24}.get(
            args.workload, max(1, int(15.0 / _REF_SECONDS_PER_SYSTEM.get(args.workload, 1.0))))
        ns = args.cpu_sample or min(B, per_core * cores)
        rate, dt, nn = cpu_reference_rate(ens, ns, cores, kind)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"first {nn} systems of the same ensemble, {_kind_text(kind)}, Pool({cores}), {dt:.1f} s"}

    # second, stronger CPU comparator: the scalar C restatement of the same arithmetic contract (oracle/radau_ref.c,
    # the parity checker of the tests), OpenMP with an explicit thread count.  N = 1 only: under torchrun OMP_NUM_THREADS is 1
    # and the cores are shared with the other ranks.
    cpu_native = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import c_oracle
            ns = int(min(B, 65536 * cores))      # about 10 s
            if gen_key == "npendulum":
                ns = int(min(B, cores * max(4, 4000 // gen_kw["n_nodes"])))
            o2 = dict(ens["options"]); mm = o2.pop("mass_matrix")
            par = ens["params"][:ns] if ens["params"] is not None else np.zeros((ns, 0))
            c_oracle.solve_batch(ens["problem"], ens["t_span"], ens["y0"][:cores], par[:cores], t_eval=ens["t_eval"], nthreads=cores, **o2)  # warm
            t0 = time.perf_counter()
            c_oracle.solve_batch(ens["problem"], ens["t_span"], ens["y0"][:ns], par, t_eval=ens["t_eval"], nthreads=cores, **o2)
            dt = time.perf_counter() - t0
            cpu_native = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
                          "sample": f"first {ns} systems of the same ensemble, scalar C restatement of the arithmetic contract "
                                    f"(oracle/radau_ref.c), omp_set_num_threads({cores}), {dt:.1f} s"}
        except Exception as exc:      # the checker library is test infrastructure; never fail the bench on it
            cpu_native = {"unavailable": repr(exc)}

    if rank == 0:
        geom = dict(zip(("grid", "block", "smem"), _launch_info(br, ens, B)))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "problem": ens["problem"], "systems_per_gpu": B,
                           "systems_total": systems_all, "t_span": list(ens["t_span"]), "n_t_eval": T,
                           "rtol": ens["options"]["rtol"], "atol": ens["options"]["atol"],
                           "first_step": ens["options"].get("first_step"),
                           "l2_policy": "inputs+outputs per step (%.0f MB) larger than the 126 MB L2" % (abytes / 1e6),
                           "parallelism": f"ensemble sharded over {world} GPU(s), no data-path collective",
                           "launch": geom, "host_placement": placement},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_baseline_native": cpu_native, "e2e": e2e,
                "value_with_gather": with_gather,
                "gpu_launches": args.steps, "clocks": clk.summary(),
                "status_histogram": dict(zip(("finished", "too_small_step", "lu_failed", "nmax_step", "terminal_event"), statuses)),
                "mean_accepted_steps": float(out_counts[4]) / B}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _launch_info(br, ens, B):
    import ctypes as C
    from dae_scipy_b200 import _abi
    hb = br.prepare(ens["problem"], ens["t_span"], (B, ens["y0"].shape[1]), t_eval=ens["t_eval"], **ens["options"])
    b = hb.struct({})
    g, bl, sm = C.c_int32(), C.c_int32(), C.c_int32()
    _abi.load().brdae_launch_info(C.byref(b), C.byref(g), C.byref(bl), C.byref(sm))
    return g.value, bl.value, sm.value


if __name__ == "__main__":
    main()
