"""The per-step record behind `dense_output=True` (OdeSolution of the reference: radauDAE.py:951-982, scipy
base.py OdeSolution) -- kernel source on the CPU against the numpy restatement of the reference."""
import numpy as np
import pytest

import dae_scipy_b200 as br
from _util import PARITY_FACTOR, run_kernel_source_on_cpu

E = br.ensembles


def _reference_sols(ens):
    from oracle import radau_numpy as orc, problems_numpy as prob
    opts = dict(ens["options"])
    mm, vi = opts.pop("mass_matrix"), opts.pop("var_index")
    res = []
    for i in range(ens["y0"].shape[0]):
        fun, jac, mass, _ = prob.build(ens["problem"], ens["params"][i] if ens["params"] is not None else (), ens["y0"].shape[1])
        res.append(orc.solve(fun, ens["t_span"], ens["y0"][i], jac=jac, mass_matrix=mass if isinstance(mm, str) else mm,
                             var_index=None if vi is None else np.asarray(vi, dtype=float), dense_output=True, **opts))
    return res


@pytest.mark.parametrize("make,rtol,atol,backward", [
    (lambda: E.pendulum_ensemble(6, t_end=1.0), 1e-6, 1e-6, False),
    (lambda: E.robertson_ensemble(4, t_end=40.0), 1e-6, 1e-8, False),
    (lambda: E.pendulum_ensemble(3, t_end=1.0), 1e-6, 1e-6, True),
])
def test_step_record_reproduces_the_reference_ode_solution(make, rtol, atol, backward):
    ens = make()
    if backward:
        ens["t_span"] = (ens["t_span"][1], ens["t_span"][0])
        ens["t_eval"] = ens["t_eval"][::-1].copy()
    out = run_kernel_source_on_cpu(ens, dense_cap=600)
    sol, refs = out["sol"], _reference_sols(ens)
    assert not sol.truncated.any()
    rng = np.random.default_rng(5)
    lo, hi = min(ens["t_span"]), max(ens["t_span"])
    tq = np.sort(rng.uniform(lo, hi, 64))
    yq = sol(tq)
    for i, r in enumerate(refs):
        k = len(r.interpolants)
        assert sol.n_steps[i] == out["counters"][i, 4]                           # every accepted step, nothing else
        if sol.n_steps[i] == k:      # same step sequence as the reference (not guaranteed under rounding noise)
            np.testing.assert_allclose(sol.ts[i, 1:k + 1], r.steps_t, rtol=1e-5, atol=1e-6)    # OdeResult.t of t_eval=None
        k = int(sol.n_steps[i])
        ref = r.sol(tq)
        n_diff = 4 if ens["problem"].startswith("pendulum") else 2
        err = np.abs(yq[i] - ref) / (atol + rtol * np.abs(ref))
        assert err[:n_diff].max() <= PARITY_FACTOR                               # differential unknowns
        # the cubic of a step ends in the step's end state, and reproduces the t_eval output of the same run
        ends = sol(sol.ts[i, 1:k + 1])[i]
        np.testing.assert_allclose(ends, sol.ys[i, :, 1:k + 1], rtol=1e-9, atol=1e-12)
    on_grid = sol(ens["t_eval"])
    np.testing.assert_allclose(on_grid, out["y"], rtol=1e-9, atol=1e-11)


def test_step_record_is_capped_and_flags_truncation():
    ens = E.pendulum_ensemble(3, t_end=1.0)
    full = run_kernel_source_on_cpu(ens, dense_cap=600)
    cut = run_kernel_source_on_cpu(ens, dense_cap=7)
    assert cut["sol"].truncated.all() and (cut["sol"].n_steps == 7).all()
    assert np.array_equal(cut["sol"].ts[:, :8], full["sol"].ts[:, :8])
    assert np.array_equal(cut["y"], full["y"]) and np.array_equal(cut["counters"], full["counters"])   # recording changes nothing else
    tq = np.linspace(0.0, float(cut["sol"].ts[:, 7].min()), 9)
    assert np.array_equal(cut["sol"](tq), full["sol"](tq))


def test_dense_output_needs_a_cap_and_host_entry_refuses_it():
    ens = E.pendulum_ensemble(2)
    with pytest.raises(ValueError, match="max_dense_steps"):
        br.solve_ivp_batched(ens["problem"], ens["t_span"], ens["y0"], ens["params"], dense_output=True, **ens["options"])


def test_substeps_reproduce_the_reference_stage_points():
    """return_substeps (radauDAE.py:661-663, 1175-1191): the collocation points recovered from the step record against
    `solver.t_substeps` / `solver.y_substeps` of the reference itself, stepped like solve_ivp (reference importable only)."""
    import contextlib
    import io
    import dae_scipy_b200 as br
    from _util import run_kernel_source_on_cpu
    from oracle import problems_numpy as prob, reference_runner
    if reference_runner.reference_location() is None:
        pytest.skip("the reference is not importable here")
    RadauDAE = reference_runner.radau_class()
    ens = br.ensembles.pendulum_ensemble(3, t_end=1.0)
    sol = run_kernel_source_on_cpu(ens, dense_cap=400)["sol"]
    tsub, ysub = sol.substeps()
    assert tsub.shape == (3, 1200) and ysub.shape == (3, 5, 1200)
    opts = dict(ens["options"])
    opts.pop("mass_matrix")
    vi = opts.pop("var_index")
    for i in range(3):
        fun, jac, mass, _ = prob.build("pendulum3", ens["params"][i], 5)
        with contextlib.redirect_stdout(io.StringIO()):
            s = RadauDAE(fun, 0.0, ens["y0"][i].copy(), 1.0, jac=jac, mass_matrix=mass, var_index=np.asarray(vi, float), **opts)
            ts, ys = [], []
            while s.status == "running":
                s.step()
                ts.extend(s.t_substeps)
                ys.extend(s.y_substeps)
        ts, ys = np.array(ts), np.array(ys).T
        k = ts.size
        assert k == 3 * sol.n_steps[i] and np.isnan(tsub[i, k:]).all() and np.isnan(ysub[i, :, k:]).all()
        np.testing.assert_allclose(tsub[i, :k], ts, rtol=0, atol=1e-6)
        assert np.max(np.abs(ysub[i, :, :k] - ys) / (1e-6 + 1e-6 * np.abs(ys))) <= 10.0        # the north_star tolerance
        assert np.array_equal(ysub[i, :, 2:k:3], sol.ys[i, :, 1:sol.n_steps[i] + 1])            # third stage = end of the step
