"""`RadauDAE` as a scipy `OdeSolver`: `scipy.integrate.solve_ivp(fun, t_span, y0, method=RadauDAE, **options)`.

The reference plugs into scipy through the `OdeSolver` protocol (class RadauDAE(OdeSolver), scipyDAE/radauDAE.py:126;
README.md:19-21).  This class does the same for ONE system of a device problem, so that a script written against the
reference runs unchanged except for `fun`: the right-hand side is not a Python closure but a `DeviceFun` -- the handle of a
CUDA device functor (a stock problem or one made by `compile_problem`) plus this system's parameters.

    sol = scipy.integrate.solve_ivp(br.DeviceFun("pendulum3", [m, r0, g]), (0, 2), y0, method=br.RadauDAE,
                                    rtol=1e-6, atol=1e-6, mass_matrix=M, var_index=[0, 0, 0, 0, 3], dense_output=True)

scipy drives a solver step by step; the kernels integrate a whole trajectory per launch.  So the first `step()` runs the
device integration to `t_bound` with the per-step record switched on (`solve_ivp_batched(..., dense_output=True)`, B = 1), and
every `step()` then hands scipy the next recorded step: `t`, `y`, `step_size`, the step's cubic as `dense_output()`.  scipy's own
loop does the rest (`t_eval`, `dense_output`, events on the interpolants -- ivp.py:659-732), exactly as it does for the
reference.  `nfev`, `njev`, `nlu` and the other RadauDAE counters are those of the whole integration.

For ensembles call `solve_ivp_batched` directly: one launch for B systems instead of B solver objects.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import DenseOutput, OdeSolver

from . import _abi


class DeviceFun:
    """`fun` of `solve_ivp(fun, ..., method=RadauDAE)`: a device problem (name or `Problem`) and the parameters of one system."""

    def __init__(self, problem, params=None):
        self.problem = problem
        self.params = None if params is None else np.atleast_1d(np.asarray(params, dtype=np.float64))

    def __call__(self, t, y):
        raise RuntimeError("a DeviceFun is evaluated on the GPU by the RadauDAE kernels; it cannot be called on the host "
                           "(use it with solve_ivp(..., method=dae_scipy_b200.RadauDAE) or solve_ivp_batched).")


class StepCubic(DenseOutput):
    """The local interpolant of one step (RadauDenseOutput, radauDAE.py:959-982): y_old + Q [x, x^2, x^3], x = (t - t_old) / h."""

    def __init__(self, t_old, t, y_old, Q):
        super().__init__(t_old, t)
        self.h = t - t_old
        self.Q = Q
        self.y_old = y_old

    def _call_impl(self, t):
        x = (np.asarray(t) - self.t_old) / self.h
        p = np.stack([x, x * x, x * x * x])                   # [3] or [3, m]
        y = self.Q @ p
        return y + (self.y_old if y.ndim == 1 else self.y_old[:, None])


def _integrate_on_device(problem, t_span, y0, params, max_dense_steps, options):
    """One system through the batched entry point with the step record on; returns (BatchedOdeSolution, host result dict)."""
    from .api import solve_ivp_batched
    res = solve_ivp_batched(problem, t_span, y0[None, :], None if params is None else params[None, :], dense_output=True,
                            max_dense_steps=max_dense_steps, **options)
    return res.sol, res.numpy()


class RadauDAE(OdeSolver):
    """Radau IIA(5) for DAEs M y' = f(t, y), integrated on the GPU (reference class scipyDAE/radauDAE.py:126).

    Also the `method=` marker of `solve_ivp_batched`.  Keyword options are the reference's (radauDAE.py:263-278) and are
    passed to the kernels unchanged; `max_dense_steps` (default 1 << 16, doubled until the trajectory fits) bounds the step
    record."""
    name = "RadauDAE"
    _integrate = staticmethod(_integrate_on_device)      # tests replace it with the CPU harness

    def __init__(self, fun, t0, y0, t_bound, vectorized=False, max_dense_steps=1 << 16, **options):
        if not isinstance(fun, DeviceFun):
            raise TypeError("RadauDAE integrates device problems: pass fun=dae_scipy_b200.DeviceFun(problem, params) "
                            "(the right-hand side is a CUDA functor, not a Python callable).")
        self.device_fun = fun
        super().__init__(lambda t, y: np.zeros_like(y), t0, y0, t_bound, vectorized, support_complex=False)
        self.options = dict(options)
        self.max_dense_steps = int(max_dense_steps)
        self._k = 0
        self._record = None
        self.y_old = None
        self.h_abs = None
        self.sol = None
        self.njev = self.nlu = 0
        self.counters = {}

    def _run(self):
        cap = self.max_dense_steps
        while True:
            sol, out = type(self)._integrate(self.device_fun.problem, (self.t, self.t_bound), np.asarray(self.y, dtype=np.float64),
                                             self.device_fun.params, cap, self.options)
            if not sol.truncated[0] or cap >= (1 << 26):
                break
            cap *= 4
        self._record = (sol, out)
        for name in _abi.COUNTER_NAMES:
            self.counters[name] = int(out[name][0])
        self.nfev, self.njev, self.nlu = self.counters["nfev"], self.counters["njev"], self.counters["nlu"]

    def _step_impl(self):
        if self._record is None:
            self._run()
        sol, out = self._record
        ns = int(sol.n_steps[0])
        if self._k >= ns:                                    # the record ends before t_bound: the integration failed there
            return False, _abi.MESSAGES.get(int(out["status"][0]), "the integration failed")
        k = self._k
        self.y_old = sol.ys[0, :, k].copy()
        t_new = float(sol.ts[0, k + 1])
        self.h_abs = abs(t_new - self.t)
        self.sol = StepCubic(float(sol.ts[0, k]), t_new, self.y_old, sol.Q[0, k])
        self.t = t_new
        self.y = sol.ys[0, :, k + 1].copy()
        self._k += 1
        return True, None

    def _dense_output_impl(self):
        return self.sol
