#!/usr/bin/env python
"""bench.py — DAE trajectories/s to t_end (FP64) for the batched RadauDAE hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (numpy restatement
                                                              # of the reference, all host cores)

A "step" is one pass of the hot path over one batch: the whole ensemble integrated from t0 to
t_end.  Default workload = BASELINE.json configs[1]: index-3 pendulum, 1,048,576 varied initial
conditions, thread-per-system, one B200 (rtol = atol = 1e-6, first_step = 1e-2, t_end = 2,
21 t_eval points; SURVEY 8d "C2 primary").  With N > 1 every rank integrates its own 1M-system
ensemble (weak scaling, different seeds); the only collective in the timed region is the
all_reduce of the summed counters.  One JSON line is printed by rank 0.
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


# measured with ncu (dram__bytes_read.sum + dram__bytes_write.sum) / systems of the profiled launch
NCU_DRAM_BYTES_PER_SYSTEM = {"pendulum3_1M_tend2": (158.95e6 + 334.40e6) / 262144}     # profiles/r01_k1_fastpaths.md, v8


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
def _cpu_worker(args):
    """Integrate a slice of the ensemble with the numpy restatement of the reference (one core)."""
    import contextlib
    import io
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import radau_numpy as orc, problems_numpy as prob
    problem, y0s, params, t_span, t_eval, options = args
    opts = dict(options)
    opts.pop("mass_matrix", None)
    var_index = opts.pop("var_index", None)
    done = 0
    # one BLAS thread per process: the pool already uses every core, and the LAPACK calls of the chain problems
    # (N = 100..500) would otherwise start a full OpenBLAS thread team in each forked worker (measured: 50x slower)
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)
    except Exception:
        limit = contextlib.nullcontext()
    with limit, contextlib.redirect_stdout(io.StringIO()):
        for i in range(y0s.shape[0]):
            fun, jac, mass, vi = prob.build(problem, params[i], y0s.shape[1])
            r = orc.solve(fun, t_span, y0s[i], jac=jac, mass_matrix=mass,
                          var_index=None if var_index is None else np.asarray(var_index, dtype=float),
                          t_eval=t_eval, **opts)
            done += int(r.status == 0)
    return y0s.shape[0], done


def cpu_reference_rate(ens, n_systems, cores):
    """traj/s of the reference algorithm (numpy restatement, oracle/radau_numpy.py) over all host
    cores, on the first `n_systems` systems of the ensemble."""
    import multiprocessing as mp
    y0 = ens["y0"][:n_systems]
    params = ens["params"][:n_systems] if ens["params"] is not None else np.zeros((n_systems, 0))
    chunks = [c for c in np.array_split(np.arange(n_systems), cores) if c.size]
    jobs = [(ens["problem"], y0[c], params[c], ens["t_span"], ens["t_eval"], ens["options"]) for c in chunks]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
        pool.map(_cpu_worker, [(ens["problem"], y0[:1], params[:1], ens["t_span"], ens["t_eval"], ens["options"])] * len(jobs))  # warm import
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    return n / dt, dt, n


def run_reference_arm(args, ens, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, min(args.ref_sample, 64 * cores))
    for _ in range(args.warmup):
        cpu_reference_rate(ens, min(per_step, 2 * cores), cores)
    rates, total_n, total_t = [], 0, 0.0
    for _ in range(args.steps):
        rate, dt, n = cpu_reference_rate(ens, per_step, cores)
        rates.append(rate); total_n += n; total_t += dt
    value = total_n / total_t
    sample = f"first {per_step} systems of the {wl_name} ensemble per step, Pool({cores}), 1 BLAS thread per process"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl_name, "systems_per_step": per_step, "parallelism": f"cpu{cores}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    ap.add_argument("--systems", type=int, default=0, help="override the ensemble size per GPU")
    ap.add_argument("--ref-sample", type=int, default=4096, help="systems per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=0, help="systems of the in-run cpu_baseline (0 = 384 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "brdae" else args.warmup

    gen_key, gen_kw, B_default = WORKLOADS[args.workload]
    B = args.systems or B_default
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
    from dae_scipy_b200 import _abi, flops
    from dae_scipy_b200.dist import reduce_statistics, status_histogram

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: every rank owns its own B-system ensemble (seed offset by rank)
    seed = {"pendulum": 0, "robertson": 1, "transistor": 2, "npendulum": 3}[gen_key] + 1000 * rank
    ens = br.ensembles.ENSEMBLES[gen_key](B, seed=seed, **gen_kw)
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

    res = None
    for _ in range(args.warmup):       # warm-up steps are whole steps: the statistics all_reduce (NCCL set-up) included
        res = one_pass()
        csum = res.counters.sum(dim=1, dtype=torch.int64)
        reduce_statistics(csum, status_histogram(res.status), dist if world > 1 else None, dev)
    barrier()
    # measured FP64 roofline denominator (burst DFMA probe on this device)
    fp64_peak = br.fp64_peak_tflops(60.0)
    barrier()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank)
    clk.__enter__()
    if True:
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
    total_ms = e_start.elapsed_time(e_stop)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    t_ms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    value = B * world * args.steps / (total_ms * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic FLOPs of one launch / its duration
    out_counts = res.counters.sum(dim=1, dtype=torch.int64).cpu().numpy()
    n_out = int(res.n_eval.sum().item())
    fl = flops.algorithmic_flops(ens["problem"], n, out_counts, n_out)
    kms = float(np.mean(kern_ms))
    achieved = fl / (kms * 1e-3) / 1e12
    abytes = flops.algorithmic_bytes(n, npar, T, B)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback"
    hbm_gbs = abytes / (kms * 1e-3) / 1e9
    # DRAM traffic of one launch: per-system dram__bytes_read + dram__bytes_write of the ncu --set full capture
    # committed under profiles/ (r01_k1_fastpaths.md: 493 MB for 262,144 pendulum systems), scaled to this launch
    traffic = NCU_DRAM_BYTES_PER_SYSTEM.get(args.workload)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                "traffic": None if traffic is None else traffic * B,
                "traffic_source": None if traffic is None else "ncu --set full, profiles/r01_k1_fastpaths.md (per system x systems of this launch)",
                "peak_source": "measured in this run: dependent-chain DFMA probe (brdae_fp64_peak_probe)",
                "kernel": {"pendulum": "k1_kernel<Pendulum<3>,MassDiag01>", "npendulum": "k2_kernel"}.get(gen_key, f"k1_kernel<{gen_key}>"),
                "kernel_ms": kms, "flops_per_launch": fl, "flops_per_trajectory": fl / B,
                "hbm": {"algorithmic_bytes": abytes, "achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak,
                        "frac": hbm_gbs / hbm_peak, "peak_source": hbm_src}}

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
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        h2d = 8 * B * (n + npar) + 8 * T
        d2h = 8 * B * n * T + 8 * B + 8 * B * n + 4 * B * 2 + 4 * B * 9
        e2e = {"value": B * world * args.steps / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "path": "solve_ivp_batched_host -> brdae_solve_host_rows (C ABI, host buffers, row per system): argument "
                       "validation, stream-ordered device alloc, H2D of y0/params, device permutation to SoA, kernel, device "
                       "permutation of the results, D2H of y[B,n,T], y_final, t_final, status, n_eval, counters into reused "
                       "page-locked buffers and the final synchronisation are all inside the timed region",
               "all_finished": bool((out["status"] == 0).all())}

    clk.__exit__(None, None, None)

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = args.cpu_sample or min(B, 384 * cores)     # about 15-20 s of CPU work on the pendulum workload
        rate, dt, nn = cpu_reference_rate(ens, ns, cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {nn} systems of the same ensemble, numpy restatement of the reference "
                                  f"(oracle/radau_numpy.py, bit-identical to RadauDAE where the reference runs), "
                                  f"Pool({cores}), {dt:.1f} s"}

    # second, stronger CPU comparator: the scalar C restatement of the same arithmetic
    # contract (oracle/radau_ref.c, the parity checker of the tests), OpenMP over all host cores
    cpu_native = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import c_oracle
            cores = os.cpu_count() or 1
            ns = int(min(B, 65536 * cores))      # about 10 s
            if gen_key == "npendulum":           # chains: a few ropes per core (the restatement stores the band in dense arrays)
                ns = int(min(B, cores * max(1, 400 // gen_kw["n_nodes"])))
            o2 = dict(ens["options"]); mm = o2.pop("mass_matrix")
            par = ens["params"][:ns] if ens["params"] is not None else np.zeros((ns, 0))
            c_oracle.solve_batch(ens["problem"], ens["t_span"], ens["y0"][:cores], par[:cores], t_eval=ens["t_eval"], **o2)  # warm
            t0 = time.perf_counter()
            c_oracle.solve_batch(ens["problem"], ens["t_span"], ens["y0"][:ns], par, t_eval=ens["t_eval"], **o2)
            dt = time.perf_counter() - t0
            cpu_native = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
                          "sample": f"first {ns} systems of the same ensemble, scalar C restatement of the arithmetic contract "
                                    f"(oracle/radau_ref.c), OpenMP over {cores} threads, {dt:.1f} s"}
        except Exception as exc:      # the checker library is test infrastructure; never fail the bench on it
            cpu_native = {"unavailable": repr(exc)}

    if rank == 0:
        geom = dict(zip(("grid", "block", "smem"), _launch_info(br, ens, B)))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "problem": ens["problem"], "systems_per_gpu": B,
                           "systems_total": B * world, "t_span": list(ens["t_span"]), "n_t_eval": T,
                           "rtol": ens["options"]["rtol"], "atol": ens["options"]["atol"],
                           "first_step": ens["options"].get("first_step"),
                           "l2_policy": "inputs+outputs per step (%.0f MB) larger than the 126 MB L2" % (abytes / 1e6),
                           "parallelism": f"ensemble sharded over {world} GPU(s), no data-path collective",
                           "launch": geom},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_baseline_native": cpu_native, "e2e": e2e,
                "gpu_launches": args.steps, "clocks": clk.summary(),
                "status_histogram": dict(zip(("finished", "too_small_step", "lu_failed", "nmax_step"), statuses)),
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
