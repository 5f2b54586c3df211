"""ORACLE (test infrastructure, NOT product code) — ctypes front end of oracle/radau_ref.c.

`solve_batch` integrates a batch of systems on the host cores with the scalar C restatement
and returns arrays in the same layout as the product's result object, so tests can compare
field by field.  Built by `oracle/Makefile` (also from `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libradau_oracle.so")

PROBLEM_IDS = {"pendulum3": 0, "pendulum2": 1, "pendulum1": 2, "robertson": 3, "transistor": 4,
               "npendulum": 5, "robertson_ode": 6, "jay": 7, "polydae2": 8, "linear7": 9}
COUNTER_NAMES = ("nfev", "njev", "nlu", "nstep", "naccpt", "nrejct", "nfailed", "nlusolve",
                 "nlusolve_err")


class OrcOptions(C.Structure):
    _fields_ = [("first_step", C.c_double), ("max_step", C.c_double), ("newton_tol", C.c_double),
                ("factor_on_non_convergence", C.c_double), ("safety_factor", C.c_double),
                ("min_factor", C.c_double), ("max_factor", C.c_double),
                ("jac_recompute_factor", C.c_double),
                ("max_newton_ite", C.c_int32), ("max_bad_ite", C.c_int32),
                ("scale_residuals", C.c_int32), ("scale_newton_norm", C.c_int32),
                ("scale_error", C.c_int32), ("zero_algebraic_error", C.c_int32),
                ("always_2nd_estimate", C.c_int32), ("predictive_controller", C.c_int32),
                ("predictive_newton_stop", C.c_int32), ("extrapolated_guess", C.c_int32),
                ("all_newton_iterations", C.c_int32), ("constant_dt", C.c_int32),
                ("jac_finite_differences", C.c_int32), ("reserved", C.c_int32),
                ("nmax_step", C.c_int64)]


def build(force=False):
    src = os.path.join(_HERE, "radau_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_solve.restype = C.c_int
        _lib.orc_eval_rhs_jac.restype = C.c_int
    return _lib


def make_options(first_step=None, max_step=np.inf, newton_tol=None, factor_on_non_convergence=0.5,
                 safety_factor=0.9, min_factor=0.2, max_factor=10, jacobianRecomputeFactor=1e-3,
                 max_newton_ite=6, max_bad_ite=None, scale_residuals=True, scale_newton_norm=True,
                 scale_error=True, zero_algebraic_error=False, bAlwaysApply2ndEstimate=True,
                 bUsePredictiveController=True, bUsePredictiveNewtonStoppingCriterion=True,
                 bUseExtrapolatedGuess=True, bPerformAllNewtonIterations=False, constant_dt=False,
                 nmax_step=None, jac="analytic", chain_linear_algebra="block"):
    """`chain_linear_algebra`: "block" = the chain spec of the n-pendulum (block elimination, what the lane-per-rope kernel
    implements), "band" = the banded LU of the general contract (the CTA-per-system band kernel)."""
    o = OrcOptions()
    o.reserved = {"block": 0, "band": 1}[chain_linear_algebra]
    o.jac_finite_differences = int(jac is None)     # jac=None: finite differences, as in the reference
    o.first_step = -1.0 if first_step is None else float(first_step)
    o.max_step = float(max_step)
    o.newton_tol = -1.0 if newton_tol is None else float(newton_tol)
    o.factor_on_non_convergence = factor_on_non_convergence
    o.safety_factor = safety_factor
    o.min_factor = min_factor
    o.max_factor = max_factor
    o.jac_recompute_factor = jacobianRecomputeFactor
    o.max_newton_ite = max_newton_ite
    o.max_bad_ite = -1 if max_bad_ite is None else int(max_bad_ite)
    o.scale_residuals = int(scale_residuals)
    o.scale_newton_norm = int(scale_newton_norm)
    o.scale_error = int(scale_error)
    o.zero_algebraic_error = int(zero_algebraic_error)
    o.always_2nd_estimate = int(bAlwaysApply2ndEstimate)
    o.predictive_controller = int(bUsePredictiveController)
    o.predictive_newton_stop = int(bUsePredictiveNewtonStoppingCriterion)
    o.extrapolated_guess = int(bUseExtrapolatedGuess)
    o.all_newton_iterations = int(bPerformAllNewtonIterations)
    o.constant_dt = int(constant_dt)
    o.nmax_step = 0 if nmax_step is None or not np.isfinite(nmax_step) else int(nmax_step)
    return o


@dataclass
class BatchResult:
    t: np.ndarray          # [T]
    y: np.ndarray          # [B, n, T], NaN where a system did not reach the point
    n_eval: np.ndarray     # [B]
    status: np.ndarray     # [B]
    t_final: np.ndarray    # [B]
    y_final: np.ndarray    # [B, n]
    counters: np.ndarray   # [B, 9]
    reports: dict | None = None   # report_cap > 0: n [B], code [B, K], newton_iterations / bad_iterations [B, K], t / dt / err1 / err2 [B, K]

    def counter(self, name):
        return self.counters[:, COUNTER_NAMES.index(name)]


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def solve_batch(problem, t_span, y0, params, *, t_eval=None, rtol=1e-3, atol=1e-6, var_index=None,
                mass_matrix=None, jac_const=None, nthreads=0, report_cap=0, **options):
    """y0: [B, n]; params: [B, p] (may be [B, 0]).  Returns BatchResult.  `report_cap` > 0 also records the reference's
    per-attempt report (`_report`, radauDAE.py:424-443) for the first `report_cap` events of every system."""
    L = lib()
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64))
    B, n = y0.shape
    params = np.asarray(params, dtype=np.float64).reshape(B, -1)
    p = params.shape[1]
    y0_soa = np.ascontiguousarray(y0.T)
    par_soa = np.ascontiguousarray(params.T) if p else np.zeros((1, max(B, 1)))
    t0, tb = float(t_span[0]), float(t_span[1])
    te = np.ascontiguousarray(np.asarray(t_eval if t_eval is not None else [], dtype=np.float64))
    T = te.size
    vi = np.zeros(n, dtype=np.int32) if var_index is None else np.ascontiguousarray(var_index, dtype=np.int32)
    at = np.ascontiguousarray(np.broadcast_to(np.asarray(atol, dtype=np.float64), (n,)))
    M = None if mass_matrix is None else np.ascontiguousarray(mass_matrix, dtype=np.float64)
    Jc = None if jac_const is None else np.ascontiguousarray(jac_const, dtype=np.float64)
    y_eval = np.empty((max(T, 1), n, B))
    n_eval = np.zeros(B, dtype=np.int32)
    t_final = np.empty(B)
    y_final = np.empty((n, B))
    status = np.empty(B, dtype=np.int32)
    counters = np.zeros((9, B), dtype=np.int32)
    o = make_options(**options)
    K = int(report_cap)
    rep_count = np.zeros(B, dtype=np.int32)
    rep_code = np.zeros((max(K, 1), B), dtype=np.int32)
    rep_it = np.full((max(K, 1), 2, B), -1, dtype=np.int32)
    rep_val = np.full((max(K, 1), 4, B), np.nan)
    L.orc_solve_report.restype = C.c_int
    rc = L.orc_solve_report(C.c_int(PROBLEM_IDS[problem]), C.c_int(n), C.c_int(p), C.c_int64(B),
                     _dptr(y0_soa), _dptr(par_soa), _dptr(M), _dptr(Jc),
                     vi.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(at), C.c_double(rtol),
                     C.c_double(t0), C.c_double(tb), _dptr(te), C.c_int(T), C.byref(o),
                     _dptr(y_eval), n_eval.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(t_final),
                     _dptr(y_final), status.ctypes.data_as(C.POINTER(C.c_int32)),
                     counters.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(nthreads), C.c_int(K),
                     rep_count.ctypes.data_as(C.POINTER(C.c_int32)), rep_code.ctypes.data_as(C.POINTER(C.c_int32)),
                     rep_it.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(rep_val))
    if rc != 0:
        raise RuntimeError(f"orc_solve failed with code {rc}")
    reports = None
    if K > 0:
        reports = dict(n=rep_count, code=rep_code.T.copy(), newton_iterations=rep_it[:, 0].T.copy(), bad_iterations=rep_it[:, 1].T.copy(),
                       t=rep_val[:, 0].T.copy(), dt=rep_val[:, 1].T.copy(), err1=rep_val[:, 2].T.copy(), err2=rep_val[:, 3].T.copy())
    return BatchResult(reports=reports, t=te, y=np.ascontiguousarray(y_eval[:T].transpose(2, 1, 0)), n_eval=n_eval,
                       status=status, t_final=t_final, y_final=np.ascontiguousarray(y_final.T),
                       counters=np.ascontiguousarray(counters.T))


def eval_rhs_jac(problem, params, t, y):
    L = lib()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.size
    p = np.ascontiguousarray(params, dtype=np.float64)
    f = np.empty(n)
    J = np.empty((n, n))
    L.orc_eval_rhs_jac(C.c_int(PROBLEM_IDS[problem]), C.c_int(n), _dptr(p), C.c_double(t), _dptr(y),
                       _dptr(f), _dptr(J))
    return f, J
