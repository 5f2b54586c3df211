"""Shared helpers of the test-suite: golden fixtures, the three CPU checkers and the parity rule."""
from __future__ import annotations

import ctypes as C
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
HAVE_REFERENCE = os.path.isdir("/root/reference/scipyDAE")

# parity tolerance of BASELINE.json north_star: every trajectory within 10*rtol (scaled by atol)
# of the reference at each t_eval point
PARITY_FACTOR = 10.0


def golden_names():
    """Trajectory fixtures of the stock problems (the events and user-problem fixtures have their own tests)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("events_", "user_")))       # user_*: tests/test_user_problem.py


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    opts = dict(meta["options"])
    opts["var_index"] = z["var_index"] if bool(z["has_var_index"]) else None
    opts["mass_matrix"] = z["mass"] if "mass" in z.files else meta["mass_matrix"]
    ens = dict(problem=meta["problem"], t_span=tuple(meta["t_span"]), y0=z["y0"],
               params=z["params"] if "params" in z.files else None, t_eval=z["t_eval"], options=opts)
    ref = {k: z[k] for k in ("y", "status", "n_eval", "counters", "y_final", "t_final")}
    return ens, ref


def scaled_error(y, y_ref, rtol, atol):
    """max over everything of |y - y_ref| / (atol + rtol |y_ref|); NaN pattern must agree."""
    assert y.shape == y_ref.shape
    nan_a, nan_b = np.isnan(y), np.isnan(y_ref)
    assert np.array_equal(nan_a, nan_b), "different sets of reached t_eval points"
    atol = np.asarray(atol, dtype=float)
    if atol.ndim == 1:
        atol = atol[None, :, None]
    err = np.abs(y - y_ref) / (atol + rtol * np.abs(y_ref))
    return float(np.nanmax(err)) if np.isfinite(err).any() else 0.0


def count_agreement(c, c_ref):
    """fraction of systems whose (naccpt, nrejct, nfailed) are identical"""
    return float((c[:, 4:7] == c_ref[:, 4:7]).all(axis=1).mean())


# ------------------------------------------------------------------ checker 1: C oracle
def run_c_oracle(ens, nthreads=0, **over):
    from oracle import c_oracle
    opts = dict(ens["options"])
    opts.update(over)
    mm = opts.pop("mass_matrix")
    B = ens["y0"].shape[0]
    params = ens["params"] if ens["params"] is not None else np.zeros((B, 0))
    r = c_oracle.solve_batch(ens["problem"], ens["t_span"], ens["y0"], params, t_eval=ens["t_eval"],
                             mass_matrix=None if isinstance(mm, str) else (np.eye(ens["y0"].shape[1]) if mm is None else mm),
                             nthreads=nthreads, **opts)
    out = dict(y=r.y, status=r.status, n_eval=r.n_eval, counters=r.counters, y_final=r.y_final, t_final=r.t_final)
    if r.reports is not None:
        out["reports"] = r.reports
    return out


def _report_buffers(B, cap):
    """Host buffers of the per-attempt report in the kernel layout, and the conversion to the reference's keys."""
    bufs = {"report_count": np.zeros(B, np.int32), "report_code": np.zeros((cap, B), np.int32),
            "report_it": np.full((cap, 2, B), -1, np.int32), "report_val": np.full((cap, 4, B), np.nan)}

    def as_dict():
        return dict(n=bufs["report_count"], code=bufs["report_code"].T.copy(), newton_iterations=bufs["report_it"][:, 0].T.copy(),
                    bad_iterations=bufs["report_it"][:, 1].T.copy(), t=bufs["report_val"][:, 0].T.copy(),
                    dt=bufs["report_val"][:, 1].T.copy(), err1=bufs["report_val"][:, 2].T.copy(), err2=bufs["report_val"][:, 3].T.copy())
    return bufs, as_dict


def assert_reports_identical(a, b, what=""):
    """Two per-attempt reports (dicts with the reference's keys, one row per system) agree bit for bit on the recorded part."""
    assert np.array_equal(a["n"], b["n"]), f"{what}: report counts differ"
    cap = a["code"].shape[1]
    for i, k in enumerate(np.minimum(a["n"], cap)):
        for key in ("code", "newton_iterations", "bad_iterations"):
            assert np.array_equal(a[key][i, :k], b[key][i, :k]), f"{what}: system {i} {key}"
        for key in ("t", "dt", "err1", "err2"):
            assert np.array_equal(a[key][i, :k], b[key][i, :k], equal_nan=True), f"{what}: system {i} {key}"


# ------------------------------------------------------------------ checker 2: numpy oracle
def run_numpy_oracle(ens, **over):
    from oracle import radau_numpy as orc, problems_numpy as prob
    opts = dict(ens["options"])
    opts.update(over)
    mm = opts.pop("mass_matrix")
    vi = opts.pop("var_index")
    fd = "jac" in opts and opts.pop("jac") is None          # jac=None: finite differences, as in the reference
    B, n = ens["y0"].shape
    T = len(ens["t_eval"])
    out = dict(y=np.full((B, n, T), np.nan), status=np.zeros(B, np.int32), n_eval=np.zeros(B, np.int32),
               counters=np.zeros((B, 9), np.int32), y_final=np.zeros((B, n)), t_final=np.zeros(B))
    for i in range(B):
        p = ens["params"][i] if ens["params"] is not None else ()
        fun, jac, mass, _ = prob.build(ens["problem"], p, n)
        if not isinstance(mm, str):
            mass = mm
        if fd:
            jac = None
        r = orc.solve(fun, ens["t_span"], ens["y0"][i], jac=jac, mass_matrix=mass,
                      var_index=None if vi is None else np.asarray(vi, dtype=float), t_eval=ens["t_eval"], **opts)
        k = r.y.shape[1]
        out["y"][i, :, :k] = r.y
        out["n_eval"][i] = k
        out["status"][i] = r.status
        out["counters"][i] = r.counters.as_array()
        out["y_final"][i] = r.y_final
        out["t_final"][i] = r.t_final
    return out


# ------------------------------------------------------------------ kernel source on the CPU
_sims = {}


def _host_sim(problem):
    """The harness library: the stock one, or -- for a user problem of dae_scipy_b200/plugin.py -- the same harness
    compiled (g++, next to the plugin library) with that problem's generated functor header."""
    import subprocess
    from dae_scipy_b200 import _abi
    src_dir = os.path.join(ROOT, "tests", "host_sim")
    path = os.path.join(src_dir, "libk1_host_sim.so")
    if problem.header_path:
        path = os.path.join(os.path.dirname(problem.header_path), "libk1_host_sim_user.so")
        if not os.path.exists(path):
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                            "-mfma", "-x", "c++", f'-DBRDAE_USER_HEADER="{problem.header_path}"',
                            os.path.join(src_dir, "k1_host_sim.cpp"), "-o", path], check=True)
    if path not in _sims:
        sim = C.CDLL(path)
        sim.sim_solve.restype = C.c_int
        sim.sim_solve.argtypes = [C.POINTER(_abi.BrdaeBatch), C.POINTER(_abi.BrdaeOptions)]
        _sims[path] = sim
    return _sims[path]


def run_kernel_source_on_cpu(ens, dense_cap=0, events=None, max_events=8, report_cap=0, **over):
    """tests/host_sim: the K1 kernel SOURCE compiled for the host, one lane.  With dense_cap > 0 the per-step
    record is requested too and returned as out["sol"] (a BatchedOdeSolution); with `events` (LinearEvents) the
    device-side event location runs and out["event_count" | "event_t" | "event_y"] are returned ([B, E, ...])."""
    import dae_scipy_b200 as br
    _sim = _host_sim(br.get_problem(ens["problem"]))
    opts = dict(ens["options"])
    opts.update(over)
    y0 = np.ascontiguousarray(ens["y0"], dtype=np.float64)
    B, n = y0.shape
    hb = br.prepare(ens["problem"], ens["t_span"], y0.shape, t_eval=ens["t_eval"], **opts)
    T = hb.t_eval.size
    y0s = np.ascontiguousarray(y0.T)
    ps = np.ascontiguousarray(ens["params"].T) if ens["params"] is not None else None
    out = {"y_eval": np.empty((max(T, 1), n, B)), "n_eval": np.empty(B, np.int32), "t_final": np.empty(B),
           "y_final": np.empty((n, B)), "status": np.empty(B, np.int32), "counters": np.empty((9, B), np.int32)}
    ptr = {k: v.ctypes.data for k, v in out.items()}
    ptr.update(y0=y0s.ctypes.data, params=ps.ctypes.data if ps is not None else None,
               t_eval=hb.t_eval.ctypes.data if T else None)
    if dense_cap:
        dense = {"dense_t": np.full((dense_cap, B), np.nan), "dense_y": np.full((dense_cap, n, B), np.nan),
                 "dense_q": np.full((dense_cap, n, 3, B), np.nan)}
        ptr.update({k: v.ctypes.data for k, v in dense.items()}, dense_cap=dense_cap)
    if report_cap:
        rbufs, rdict = _report_buffers(B, report_cap)
        ptr.update({k: v.ctypes.data for k, v in rbufs.items()}, report_cap=report_cap)
    if events:
        from dae_scipy_b200.api import pack_linear_events
        hb.events = pack_linear_events(events, n, hb.problem.n_event_functions)
        E_ = len(events)
        evb = {"event_count": np.zeros((E_, B), np.int32), "event_t": np.full((E_, max_events, B), np.nan),
               "event_y": np.full((E_, max_events, n, B), np.nan)}
        ptr.update({k: v.ctypes.data for k, v in evb.items()}, event_cap=max_events)
    b = hb.struct(ptr)
    rc = _sim.sim_solve(C.byref(b), C.byref(hb.options))
    assert rc == 0, f"sim_solve returned {rc}"
    extra = {}
    if report_cap:
        extra["reports"] = rdict()
    if events:
        extra.update(event_count=evb["event_count"].T.copy(), event_t=evb["event_t"].transpose(2, 0, 1).copy(),
                     event_y=evb["event_y"].transpose(3, 0, 1, 2).copy())
    if dense_cap:
        from dae_scipy_b200.api import make_ode_solution
        extra["sol"] = make_ode_solution(hb, y0, dense["dense_t"], dense["dense_y"], dense["dense_q"], out["status"], out["counters"])
    return dict(extra, y=np.ascontiguousarray(out["y_eval"][:T].transpose(2, 1, 0)), status=out["status"],
                n_eval=out["n_eval"], counters=np.ascontiguousarray(out["counters"].T),
                y_final=np.ascontiguousarray(out["y_final"].T), t_final=out["t_final"])


def run_chain_kernel_source_on_cpu(ens, dense_cap=0, report_cap=0, **over):
    """tests/host_sim/k4_host_sim.cpp: the lane-per-rope chain kernel SOURCE (csrc/chain_k4.cuh) compiled for the host, one
    lane with slot stride 1, pulling the ropes of the batch one after the other from the work counter."""
    import dae_scipy_b200 as br
    from dae_scipy_b200 import _abi
    path = os.path.join(ROOT, "tests", "host_sim", "libk4_host_sim.so")
    if path not in _sims:
        sim = C.CDLL(path)
        sim.sim_solve_chain.restype = C.c_int
        sim.sim_solve_chain.argtypes = [C.POINTER(_abi.BrdaeBatch), C.POINTER(_abi.BrdaeOptions)]
        _sims[path] = sim
    opts = dict(ens["options"])
    opts.update(over)
    y0 = np.ascontiguousarray(ens["y0"], dtype=np.float64)
    B, n = y0.shape
    hb = br.prepare(ens["problem"], ens["t_span"], y0.shape, t_eval=ens["t_eval"], **opts)
    T = hb.t_eval.size
    y0s = np.ascontiguousarray(y0.T)
    out = {"y_eval": np.empty((max(T, 1), n, B)), "n_eval": np.empty(B, np.int32), "t_final": np.empty(B),
           "y_final": np.empty((n, B)), "status": np.empty(B, np.int32), "counters": np.empty((9, B), np.int32)}
    ptr = {k: v.ctypes.data for k, v in out.items()}
    ptr.update(y0=y0s.ctypes.data, params=None, t_eval=hb.t_eval.ctypes.data if T else None)
    if dense_cap:
        dense = {"dense_t": np.full((dense_cap, B), np.nan), "dense_y": np.full((dense_cap, n, B), np.nan),
                 "dense_q": np.full((dense_cap, n, 3, B), np.nan)}
        ptr.update({k: v.ctypes.data for k, v in dense.items()}, dense_cap=dense_cap)
    if report_cap:
        rbufs, rdict = _report_buffers(B, report_cap)
        ptr.update({k: v.ctypes.data for k, v in rbufs.items()}, report_cap=report_cap)
    b = hb.struct(ptr)
    rc = _sims[path].sim_solve_chain(C.byref(b), C.byref(hb.options))
    assert rc == 0, f"sim_solve_chain returned {rc}"
    extra = {}
    if report_cap:
        extra["reports"] = rdict()
    if dense_cap:
        from dae_scipy_b200.api import make_ode_solution
        extra["sol"] = make_ode_solution(hb, y0, dense["dense_t"], dense["dense_y"], dense["dense_q"], out["status"], out["counters"])
    return dict(extra, y=np.ascontiguousarray(out["y_eval"][:T].transpose(2, 1, 0)), status=out["status"],
                n_eval=out["n_eval"], counters=np.ascontiguousarray(out["counters"].T),
                y_final=np.ascontiguousarray(out["y_final"].T), t_final=out["t_final"])


# ------------------------------------------------------------------ the product (GPU)
def run_gpu(ens, host_api=False, report_cap=0, **over):
    import dae_scipy_b200 as br
    opts = dict(ens["options"])
    opts.update(over)
    if host_api:
        r = br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], **opts)
        return dict(y=np.ascontiguousarray(r["y"]), status=r["status"].copy(), n_eval=r["n_eval"].copy(),
                    counters=np.ascontiguousarray(r["counters"]), y_final=np.ascontiguousarray(r["y_final"]),
                    t_final=r["t_final"].copy())
    import torch
    res = br.solve_ivp_batched(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"],
                               report=bool(report_cap), max_reports=max(int(report_cap), 1), **opts)
    torch.cuda.synchronize()
    o = res.numpy()
    out = dict(y=o["y"], status=o["status"], n_eval=o["n_eval"], counters=o["counters"], y_final=o["y_final"],
               t_final=o["t_final"])
    if report_cap:
        out["reports"] = o["reports"]
    return out


def assert_bit_identical(a, b, what=""):
    for k in ("status", "n_eval", "counters"):
        assert np.array_equal(a[k], b[k]), f"{what}: {k} differ"
    for k in ("y", "y_final", "t_final"):
        assert np.array_equal(a[k], b[k], equal_nan=True), f"{what}: {k} not bit-identical"


def well_conditioned(name):
    """Fixtures on which the reference itself is reproducible under rounding noise:
    step counts must then be identical, not just statistically close."""
    return name in ("g1_pendulum3_single", "g2_robertson_single", "ens_pendulum3_tend2", "ens_pendulum2_tend2",
                    "ens_pendulum1_tend2", "opt_pendulum3_max_step", "opt_pendulum3_nmax_step",
                    "opt_pendulum3_dense_mass", "opt_pendulum3_backward", "opt_pendulum3_controller_off",
                    "ens_npendulum_n4", "ens_npendulum_n10", "ens_npendulum_n20", "ens_npendulum_n50", "ens_npendulum_n100",
                    "fd_npendulum_n4", "fd_npendulum_n20",
                    "script_robertson",
                    "kat_jay", "kat_polydae2", "kat_linear7_with_mass", "kat_linear7_without_mass", "kat_robertson_ode")


def _numpy_oracle_worker(ens):
    import warnings
    warnings.filterwarnings("ignore")
    return run_numpy_oracle(ens)


def run_numpy_oracle_parallel(ens, nproc=None):
    """The reference-equivalent numpy restatement over all host cores (one slice per process)."""
    import multiprocessing as mp
    nproc = nproc or min(os.cpu_count() or 1, 32)
    B = ens["y0"].shape[0]
    chunks = [c for c in np.array_split(np.arange(B), nproc) if c.size]
    subs = [dict(ens, y0=ens["y0"][c], params=None if ens["params"] is None else ens["params"][c]) for c in chunks]
    with mp.get_context("fork").Pool(len(subs)) as pool:
        outs = pool.map(_numpy_oracle_worker, subs)
    return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}


def run_reference_or_port_parallel(ens):
    """(result, kind): the UNMODIFIED reference over all host cores when it is importable (oracle/_ref travels with the
    snapshot; /root/reference exists in the build container), else the numpy port that is bit-identical to it."""
    from oracle import reference_runner
    if reference_runner.reference_location() is not None:
        return reference_runner.run_reference_parallel(ens), "reference"
    return run_numpy_oracle_parallel(ens), "port"


def parity_statistics(out, ref, rtol, atol, n_diff):
    """north_star acceptance numbers: fraction of systems with identical (naccpt, nrejct, nfailed); per-system max
    of |dy| / (atol + rtol |y_ref|) over the differential unknowns (first n_diff) and over the algebraic ones."""
    err = np.abs(out["y"] - ref["y"]) / (atol + rtol * np.abs(ref["y"]))
    return dict(identical_counts=count_agreement(out["counters"], ref["counters"]),
                status_equal=bool(np.array_equal(out["status"], ref["status"])),
                diff_err=np.nanmax(err[:, :n_diff, :], axis=(1, 2)),
                alg_err=np.nanmax(err[:, n_diff:, :], axis=(1, 2)) if err.shape[1] > n_diff else np.zeros(err.shape[0]))
