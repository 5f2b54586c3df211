"""`scipy.integrate.solve_ivp(DeviceFun(...), ..., method=RadauDAE)`: the OdeSolver-protocol entry (reference:
class RadauDAE(OdeSolver), scipyDAE/radauDAE.py:126; README.md:19-21).  The device integration is replaced by the CPU
harness of the kernel source here; the GPU test runs the real thing."""
import contextlib
import io

import numpy as np
import pytest
from scipy.integrate import solve_ivp

import dae_scipy_b200 as br
from dae_scipy_b200 import ensembles as E
from _util import run_kernel_source_on_cpu


def _harness_integrate(problem, t_span, y0, params, max_dense_steps, options):
    ens = dict(problem=problem, t_span=t_span, y0=y0[None, :], params=None if params is None else params[None, :], t_eval=None,
               options=options)
    out = run_kernel_source_on_cpu(ens, dense_cap=max_dense_steps)
    host = {name: out["counters"][:, i] for i, name in enumerate(br.api._abi.COUNTER_NAMES)}
    host["status"] = out["status"]
    return out["sol"], host


def _reference_solve(ens, i, **kw):
    from oracle import problems_numpy as prob, reference_runner
    if reference_runner.reference_location() is None:
        pytest.skip("the reference is not importable here")
    RadauDAE = reference_runner.radau_class()
    opts = dict(ens["options"])
    opts.pop("mass_matrix")
    fun, jac, mass, _ = prob.build(ens["problem"], ens["params"][i], ens["y0"].shape[1])
    opts["var_index"] = np.asarray(opts["var_index"], float)
    with contextlib.redirect_stdout(io.StringIO()):
        return solve_ivp(fun, ens["t_span"], ens["y0"][i].copy(), method=RadauDAE, jac=jac, mass_matrix=mass, **opts, **kw)


def _ours(ens, i, **kw):
    return solve_ivp(br.DeviceFun(ens["problem"], ens["params"][i]), ens["t_span"], ens["y0"][i].copy(), method=br.RadauDAE,
                     **ens["options"], **kw)


def test_solve_ivp_with_the_device_solver_matches_the_reference_call(monkeypatch):
    monkeypatch.setattr(br.RadauDAE, "_integrate", staticmethod(_harness_integrate))
    ens = E.pendulum_ensemble(2, t_end=1.0)
    te = np.linspace(0, 1, 11)
    for i in range(2):
        ref = _reference_solve(ens, i, t_eval=te, dense_output=True)
        sol = _ours(ens, i, t_eval=te, dense_output=True)
        assert sol.status == ref.status == 0 and sol.success
        assert sol.t.shape == ref.t.shape and np.array_equal(sol.t, ref.t)
        tol = 10 * (1e-6 + 1e-6 * np.abs(ref.y))
        assert (np.abs(sol.y - ref.y) <= tol).all()
        tt = np.linspace(0, 1, 57)
        assert (np.abs(sol.sol(tt) - ref.sol(tt)) <= 10 * (1e-6 + 1e-6 * np.abs(ref.sol(tt)))).all()
        assert (sol.nfev, sol.njev, sol.nlu) == (ref.nfev, ref.njev, ref.nlu)
        # without t_eval scipy returns the steps themselves: same number of steps as the reference took
        steps, rsteps = _ours(ens, i), _reference_solve(ens, i)
        assert steps.t.size == rsteps.t.size and np.allclose(steps.t, rsteps.t, rtol=0, atol=1e-6)


def test_events_are_located_by_scipy_on_the_step_interpolants(monkeypatch):
    monkeypatch.setattr(br.RadauDAE, "_integrate", staticmethod(_harness_integrate))
    ens = E.pendulum_ensemble(1, t_end=2.0)
    ev = lambda t, y: y[0]                      # x crosses zero
    ev.terminal, ev.direction = False, 0.0
    ref = _reference_solve(ens, 0, events=ev)
    sol = _ours(ens, 0, events=ev)
    assert len(sol.t_events[0]) == len(ref.t_events[0]) > 0
    assert np.allclose(sol.t_events[0], ref.t_events[0], rtol=0, atol=1e-5)
    ev.terminal = True
    sol, ref = _ours(ens, 0, events=ev), _reference_solve(ens, 0, events=ev)
    assert sol.status == ref.status == 1 and abs(sol.t[-1] - ref.t[-1]) < 1e-5


def test_failure_and_argument_errors(monkeypatch):
    monkeypatch.setattr(br.RadauDAE, "_integrate", staticmethod(_harness_integrate))
    ens = E.pendulum_ensemble(1)
    with pytest.raises(TypeError, match="DeviceFun"):
        solve_ivp(lambda t, y: y, (0, 1), ens["y0"][0], method=br.RadauDAE)
    with pytest.raises(RuntimeError, match="GPU"):
        br.DeviceFun("pendulum3", ens["params"][0])(0.0, ens["y0"][0])
    opts = dict(ens["options"], nmax_step=5)     # the integration stops after 5 steps: scipy reports the failure
    sol = solve_ivp(br.DeviceFun("pendulum3", ens["params"][0]), ens["t_span"], ens["y0"][0], method=br.RadauDAE, **opts)
    assert sol.status == -1 and not sol.success and sol.t.size == 6


@pytest.mark.gpu
def test_gpu_solve_ivp_with_the_device_solver():
    ens = E.pendulum_ensemble(3)
    te = np.linspace(0, 2, 21)
    batched = br.solve_ivp_batched(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=te, **ens["options"]).numpy()
    for i in range(3):
        sol = _ours(ens, i, t_eval=te, dense_output=True)
        assert sol.status == 0 and np.array_equal(sol.t, te)
        # the step record is bit-identical to the batched run's own interpolation only at the step ends; t_eval goes through
        # scipy's evaluation of the same cubics (one rounding of x): compare to rounding level
        assert np.allclose(sol.y, batched["y"][i], rtol=1e-12, atol=1e-13)
        assert (sol.nfev, sol.njev, sol.nlu) == (batched["nfev"][i], batched["njev"][i], batched["nlu"][i])
