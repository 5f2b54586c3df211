"""ctypes binding of libbrdae.so (the C ABI declared in include/brdae.h).

Nothing here computes: it loads the CUDA library, mirrors the two POD structs and turns the
RadauDAE keyword surface (reference radauDAE.py:263-278) into a `brdae_options`.  If the
library is missing the import of the product path fails loudly — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRDAE_LIB") or os.path.join(_HERE, "libbrdae.so")   # BRDAE_LIB: tuning builds

COUNTER_NAMES = ("nfev", "njev", "nlu", "nstep", "naccpt", "nrejct", "nfailed", "nlusolve",
                 "nlusolve_errorest")
NCOUNTERS = len(COUNTER_NAMES)

ABI_VERSION = 4
MAX_EVENTS = 4
STATUS_FINISHED = 0
STATUS_EVENT = 1
STATUS_TOO_SMALL_STEP = -1
STATUS_LU_FAILED = -2
STATUS_NMAX_STEP = 2

# OdeResult.message texts (scipy ivp.py:13-14, base.py:130; radauDAE.py:986-989)
MESSAGES = {
    STATUS_FINISHED: "The solver successfully reached the end of the integration interval.",
    STATUS_EVENT: "A termination event occurred.",
    STATUS_TOO_SMALL_STEP: "Required step size is less than spacing between numbers.",
    STATUS_LU_FAILED: "LU decomposition failed",
    STATUS_NMAX_STEP: "Too many time steps have been performed",
}


class BrdaeOptions(C.Structure):
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


class BrdaeBatch(C.Structure):
    _fields_ = [("problem", C.c_int32), ("n", C.c_int32), ("n_params", C.c_int32),
                ("n_t_eval", C.c_int32), ("n_systems", C.c_int64),
                ("y0", C.c_void_p), ("params", C.c_void_p), ("t_eval", C.c_void_p),
                ("mass", C.c_void_p), ("jac_const", C.c_void_p), ("var_index", C.c_void_p),
                ("atol", C.c_void_p), ("rtol", C.c_double), ("t0", C.c_double),
                ("t_bound", C.c_double),
                ("y_eval", C.c_void_p), ("n_eval", C.c_void_p), ("t_final", C.c_void_p),
                ("y_final", C.c_void_p), ("status", C.c_void_p), ("counters", C.c_void_p),
                ("dense_cap", C.c_int32), ("reserved0", C.c_int32),
                ("dense_t", C.c_void_p), ("dense_y", C.c_void_p), ("dense_q", C.c_void_p),
                ("n_events", C.c_int32), ("event_cap", C.c_int32),
                ("event_coef", C.c_void_p), ("event_level", C.c_void_p), ("event_direction", C.c_void_p),
                ("event_terminal", C.c_void_p), ("event_count", C.c_void_p), ("event_t", C.c_void_p),
                ("event_y", C.c_void_p), ("event_functor", C.c_void_p),
                ("report_cap", C.c_int32), ("reserved1", C.c_int32), ("report_count", C.c_void_p),
                ("report_code", C.c_void_p), ("report_it", C.c_void_p), ("report_val", C.c_void_p)]


# keyword -> (options field, default) : the RadauDAE constructor surface, radauDAE.py:263-278
_FLOAT_OPTS = {
    "factor_on_non_convergence": ("factor_on_non_convergence", 0.5),
    "safety_factor": ("safety_factor", 0.9),
    "min_factor": ("min_factor", 0.2),
    "max_factor": ("max_factor", 10.0),
    "jacobianRecomputeFactor": ("jac_recompute_factor", 1e-3),
}
_BOOL_OPTS = {
    "scale_residuals": ("scale_residuals", True),
    "scale_newton_norm": ("scale_newton_norm", True),
    "scale_error": ("scale_error", True),
    "zero_algebraic_error": ("zero_algebraic_error", False),
    "bAlwaysApply2ndEstimate": ("always_2nd_estimate", True),
    "bUsePredictiveController": ("predictive_controller", True),
    "bUsePredictiveNewtonStoppingCriterion": ("predictive_newton_stop", True),
    "bUseExtrapolatedGuess": ("extrapolated_guess", True),
    "bPerformAllNewtonIterations": ("all_newton_iterations", False),
    "constant_dt": ("constant_dt", False),
}
# accepted for signature compatibility, no effect on the batched path (printing / debugging aids
# of the reference: bPrint :267, bPrintProgress :268, bDebug :271, bReport :276; `vectorized` and
# `jac_sparsity` only matter for the finite-difference Jacobian)
IGNORED_OPTS = ("bPrint", "bPrintProgress", "bDebug", "bReport", "vectorized", "jac_sparsity")

SOLVER_OPTION_NAMES = (tuple(_FLOAT_OPTS) + tuple(_BOOL_OPTS) +
                       ("first_step", "max_step", "newton_tol", "max_newton_ite", "max_bad_ite",
                        "nmax_step"))


def make_options(*, first_step=None, max_step=math.inf, newton_tol=None, max_newton_ite=6,
                 max_bad_ite=None, nmax_step=math.inf, jac_finite_differences=False, **kw) -> BrdaeOptions:
    """Keyword surface of RadauDAE.__init__ -> brdae_options (validation as in scipy
    common.py:10-24: first_step/max_step must be positive)."""
    o = BrdaeOptions()
    if first_step is not None and not first_step > 0:
        raise ValueError("`first_step` must be positive.")
    if not max_step > 0:
        raise ValueError("`max_step` must be positive.")
    o.first_step = -1.0 if first_step is None else float(first_step)
    o.max_step = float(max_step)
    o.newton_tol = -1.0 if newton_tol is None else float(newton_tol)
    if int(max_newton_ite) < 1:
        raise ValueError("`max_newton_ite` must be at least 1.")
    o.max_newton_ite = int(max_newton_ite)
    o.max_bad_ite = -1 if max_bad_ite is None else int(max_bad_ite)
    o.nmax_step = 0 if nmax_step is None or math.isinf(nmax_step) else int(nmax_step)
    o.jac_finite_differences = int(bool(jac_finite_differences))
    for key, (field, default) in _FLOAT_OPTS.items():
        setattr(o, field, float(kw.pop(key, default)))
    for key, (field, default) in _BOOL_OPTS.items():
        setattr(o, field, int(bool(kw.pop(key, default))))
    if kw:
        raise TypeError(f"unknown solver options: {sorted(kw)}")
    return o


_lib = None
_plugin_libs = {}      # path -> CDLL of a plugin build (dae_scipy_b200/plugin.py)


def load(path: str | None = None):
    """Load libbrdae.so (or the plugin build at `path`) and declare its prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is not None and path in _plugin_libs:
        return _plugin_libs[path]
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build the CUDA extension first (python -m dae_scipy_b200.build or "
            "__graft_entry__.build()).  The batched solver has no CPU fallback.")
    lib = C.CDLL(p)
    lib.brdae_abi_version.restype = C.c_int
    have = int(lib.brdae_abi_version())
    if have != ABI_VERSION:        # a stale build would misread the larger brdae_batch of this version: refuse it
        raise RuntimeError(f"{p} implements ABI version {have}, this package needs {ABI_VERSION}: rebuild it "
                           "(python -m dae_scipy_b200.build --force; plugin libraries: delete dae_scipy_b200/user_libs).")
    lib.brdae_strerror.restype = C.c_char_p
    lib.brdae_strerror.argtypes = [C.c_int]
    lib.brdae_problem_lookup.restype = C.c_int
    lib.brdae_problem_lookup.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32)]
    lib.brdae_default_options.restype = None
    lib.brdae_default_options.argtypes = [C.POINTER(BrdaeOptions)]
    lib.brdae_workspace_bytes.restype = C.c_size_t
    lib.brdae_workspace_bytes.argtypes = [C.POINTER(BrdaeBatch), C.POINTER(BrdaeOptions)]
    lib.brdae_solve.restype = C.c_int
    lib.brdae_solve.argtypes = [C.POINTER(BrdaeBatch), C.POINTER(BrdaeOptions), C.c_void_p,
                                C.c_size_t, C.c_void_p]
    lib.brdae_solve_host.restype = C.c_int
    lib.brdae_solve_host.argtypes = [C.POINTER(BrdaeBatch), C.POINTER(BrdaeOptions)]
    lib.brdae_launch_info.restype = C.c_int
    lib.brdae_launch_info.argtypes = [C.POINTER(BrdaeBatch), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.brdae_solve_host_rows.restype = C.c_int
    lib.brdae_solve_host_rows.argtypes = [C.POINTER(BrdaeBatch), C.POINTER(BrdaeOptions)]
    lib.brdae_alloc_pinned.restype = C.c_int
    lib.brdae_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    lib.brdae_free_pinned.restype = C.c_int
    lib.brdae_free_pinned.argtypes = [C.c_void_p]
    lib.brdae_fp64_peak_probe.restype = C.c_int
    lib.brdae_fp64_peak_probe.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_void_p]
    if path is None:
        _lib = lib
    else:
        _plugin_libs[path] = lib
    return lib


EXPORTED_SYMBOLS = ("brdae_abi_version", "brdae_strerror", "brdae_problem_lookup",
                    "brdae_default_options", "brdae_workspace_bytes", "brdae_solve",
                    "brdae_solve_host", "brdae_solve_host_rows", "brdae_launch_info", "brdae_fp64_peak_probe",
                    "brdae_alloc_pinned", "brdae_free_pinned")


ERR_NONFINITE_Y0 = 6          # include/brdae.h


def check(rc: int, what: str = "brdae call"):
    if rc != 0:
        msg = load().brdae_strerror(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")
