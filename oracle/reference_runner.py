"""ORACLE (test infrastructure, NOT product code) -- drives THE REFERENCE ITSELF, `scipyDAE.radauDAE.RadauDAE`.

The reference is pure Python.  It is imported from `/root/reference` in the build container, or from
`oracle/_ref/` -- a `pip install --target oracle/_ref` of the unmodified reference made by
`__graft_entry__.build()` (git-ignored, so no reference source enters the history; it travels to the GPU box with the
snapshot, where `/root/reference` does not exist).  Used by `oracle/gen_golden.py` (fixtures), by
`bench.py --impl reference` / `cpu_baseline` (the CPU arm, kind "reference") and by the self-consistency measurement
`profiles/tools/ref_selfconsistency.py`.  The solver is stepped exactly like scipy's `solve_ivp` loop (ivp.py:659-732)
so that the RadauDAE counters can be read from the solver object.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_INSTALL = os.path.join(_HERE, "_ref")


def reference_location():
    """Directory to put on sys.path for `import scipyDAE`, or None when the reference is not available."""
    for cand in ("/root/reference", REF_INSTALL):
        if os.path.isfile(os.path.join(cand, "scipyDAE", "radauDAE.py")):
            return cand
    return None


def install_reference(src="/root/reference"):
    """`pip install --target oracle/_ref` of the unmodified reference (from a scratch copy: the source tree is read-only).
    Returns the install directory or None when the source is not there."""
    import shutil
    import subprocess
    import tempfile
    if not os.path.isfile(os.path.join(src, "setup.py")):
        return REF_INSTALL if os.path.isdir(os.path.join(REF_INSTALL, "scipyDAE")) else None
    marker = os.path.join(REF_INSTALL, "scipyDAE", "radauDAE.py")
    if os.path.exists(marker) and os.path.getmtime(marker) >= os.path.getmtime(os.path.join(src, "scipyDAE", "radauDAE.py")):
        return REF_INSTALL
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(src, copy)
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
                        "--find-links", "/opt/wheelhouse", "--target", REF_INSTALL, copy], check=True, capture_output=True)
    return REF_INSTALL


_RadauDAE = None


def radau_class():
    global _RadauDAE
    if _RadauDAE is None:
        loc = reference_location()
        if loc is None:
            raise ImportError("the reference (scipyDAE) is neither at /root/reference nor installed under oracle/_ref")
        if loc not in sys.path:
            sys.path.insert(0, loc)
        from scipyDAE.radauDAE import RadauDAE
        _RadauDAE = RadauDAE
    return _RadauDAE


def run_reference(ens, over=None, build_system=None):
    """Integrate every system of the ensemble dict with the reference solver; returns the same fields as the product's
    result (y [B, n, T] with NaN beyond the reached points, status, n_eval, the nine counters, y_final, t_final)."""
    from oracle import problems_numpy as prob
    RadauDAE = radau_class()
    build_system = build_system or prob.build
    opts = dict(ens["options"])
    opts.update(over or {})
    mm = opts.pop("mass_matrix")
    vi = opts.pop("var_index")
    nmax_step = opts.pop("nmax_step", np.inf)
    fd = "jac" in opts and opts.pop("jac") is None     # jac=None: the reference's finite-difference Jacobian
    B, n = ens["y0"].shape
    te = np.asarray(ens["t_eval"], dtype=float)
    T = te.size
    t0, tf = ens["t_span"]
    y = np.full((B, n, T), np.nan)
    status = np.zeros(B, dtype=np.int32)
    n_eval = np.zeros(B, dtype=np.int32)
    counters = np.zeros((B, 9), dtype=np.int32)
    y_final = np.zeros((B, n))
    t_final = np.zeros(B)
    for i in range(B):
        p = ens["params"][i] if ens["params"] is not None else ()
        fun, jac, mass, _ = build_system(ens["problem"], p, n)
        if not isinstance(mm, str):
            mass = mm
        if fd:
            jac = None
            if ens["problem"] == "rod" and "jac_sparsity" not in opts:      # the band example: tridiagonal
                import scipy.sparse
                opts["jac_sparsity"] = scipy.sparse.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1], format="csc")
            if ens["problem"] == "npendulum" and "jac_sparsity" not in opts:
                # the reference's own chain run: banded sparsity for num_jac's column groups (recursive_pendulum.py:165-166)
                import scipy.sparse
                opts["jac_sparsity"] = scipy.sparse.diags(diagonals=[np.ones(n - abs(k)) for k in range(-10, 10)],
                                                          offsets=list(range(-10, 10)), format="csc")
        with contextlib.redirect_stdout(io.StringIO()):
            solver = RadauDAE(fun, t0, np.array(ens["y0"][i], dtype=float), tf, jac=jac, mass_matrix=mass,
                              var_index=None if vi is None else np.asarray(vi, dtype=float), **opts)
            ei = 0
            st = None
            nsteps = 0
            while st is None:
                nsteps += 1
                if nsteps > nmax_step:
                    st = 2
                    break
                solver.step()
                if solver.status == 'finished':
                    st = 0
                elif solver.status == 'failed':
                    st = -1
                    break
                t = solver.t
                if tf > t0:
                    ei_new = int(np.searchsorted(te, t, side='right'))
                    pts = te[ei:ei_new]
                else:
                    # user order is decreasing: points >= t
                    ei_new = ei
                    while ei_new < T and te[ei_new] >= t:
                        ei_new += 1
                    pts = te[ei:ei_new]
                if pts.size:
                    y[i, :, ei:ei_new] = solver.dense_output()(pts)
                    ei = ei_new
        n_eval[i] = ei
        status[i] = st
        counters[i] = [solver.nfev, solver.njev, solver.nlu, solver.nstep, solver.naccpt, solver.nrejct,
                       solver.nfailed, solver._nlusolve, solver._nlusolve_errorest]
        y_final[i] = solver.y
        t_final[i] = solver.t
    return dict(y=y, status=status, n_eval=n_eval, counters=counters, y_final=y_final, t_final=t_final)


def _worker(args):
    import warnings
    warnings.filterwarnings("ignore")
    import scipy.linalg  # noqa: F401  (scipy's own OpenBLAS must be loaded before the limit is set, or it is not limited)
    import scipy.sparse.linalg  # noqa: F401
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)      # one BLAS thread per process: the pool already uses every core
    except Exception:
        limit = contextlib.nullcontext()
    with limit:
        return run_reference(args)


def run_reference_parallel(ens, nproc=None):
    """The reference over all host cores (`multiprocessing.Pool`, one slice of the ensemble per process)."""
    import multiprocessing as mp
    nproc = nproc or min(os.cpu_count() or 1, 64)
    B = ens["y0"].shape[0]
    chunks = [c for c in np.array_split(np.arange(B), nproc) if c.size]
    subs = [dict(ens, y0=ens["y0"][c], params=None if ens["params"] is None else ens["params"][c]) for c in chunks]
    with mp.get_context("fork").Pool(len(subs)) as pool:
        outs = pool.map(_worker, subs)
    return {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
