"""dae_scipy_b200 — B200-native batched RadauDAE (drop-in for solve_ivp(method=RadauDAE) on
ensembles of small DAE systems).  Hot path: hand-written CUDA for sm_100a in csrc/, reached
through the C ABI of include/brdae.h; this package is the thin Python host side."""
from .api import (BatchedOdeResult, DeviceFun, FunctorEvent, LinearEvent, RadauDAE, fp64_peak_tflops, host_buffers, prepare, solve_ivp_batched,
                  solve_ivp_batched_host)
from .plugin import compile_band_problem, compile_problem
from .problems import PROBLEMS, Problem, get_problem
from . import ensembles, problems

__all__ = ["solve_ivp_batched", "solve_ivp_batched_host", "RadauDAE", "DeviceFun", "BatchedOdeResult", "LinearEvent", "FunctorEvent", "prepare",
           "fp64_peak_tflops", "host_buffers", "compile_problem", "compile_band_problem", "PROBLEMS", "Problem", "get_problem", "ensembles", "problems"]
