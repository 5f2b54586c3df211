"""ORACLE (test infrastructure) — generate tests/golden/*.npz by running THE REFERENCE ITSELF.

Run in the build container, where /root/reference is mounted:
    python oracle/gen_golden.py [name-prefix ...]
It imports `scipyDAE.radauDAE.RadauDAE` from /root/reference and drives it with
`scipy.integrate.solve_ivp(method=RadauDAE)` exactly like the reference scripts do
(tests/pendulum.py:266-280, tests/robertson.py:105-116, tests/transistor_amplifier.py:78-88,
tests/recursive_pendulum.py:233-246), on the seeded ensembles of dae_scipy_b200.ensembles.  The
right-hand sides are the parametrised restatements of oracle/problems_numpy.py (checked in
tests/test_oracle_numpy.py to return the same bits as the reference closures) and the Jacobians
are the analytic ones, so that every implementation is handed the same Jacobian.

Each fixture stores inputs (y0, params, t_eval, options) and the reference outputs
(y [B, n, T], status, the nine counters).  Nothing is copied from the reference sources.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from scipy.integrate import solve_ivp  # noqa: E402
from oracle import reference_runner  # noqa: E402
RadauDAE = reference_runner.radau_class()  # the reference

from dae_scipy_b200 import ensembles  # noqa: E402  (numpy-only input generators)
from oracle import problems_numpy as prob  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_system(problem, p, n=None):
    return prob.build(problem, p, n)


def run_reference_with_counters(ens, over=None):
    """The reference solver stepped like scipy's solve_ivp loop, counters read from the solver object
    (oracle/reference_runner.py)."""
    return reference_runner.run_reference(ens, over, build_system)


ONLY = tuple(sys.argv[1:])       # optional name prefixes: regenerate just those fixtures


def save(name, ens, over=None):
    if ONLY and not name.startswith(ONLY):
        return
    out = run_reference_with_counters(ens, over)
    opts = dict(ens["options"])
    opts.update(over or {})
    vi = opts.pop("var_index")
    mm = opts.pop("mass_matrix")
    meta = dict(problem=ens["problem"], t_span=list(ens["t_span"]),
                options={k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in opts.items()},
                mass_matrix=mm if isinstance(mm, str) else "array")
    arrays = dict(y0=ens["y0"], t_eval=np.asarray(ens["t_eval"], dtype=float),
                  var_index=np.zeros(0, dtype=np.int32) if vi is None else np.asarray(vi, dtype=np.int32),
                  has_var_index=np.array(vi is not None), meta=np.array(json.dumps(meta)), **out)
    if ens["params"] is not None:
        arrays["params"] = ens["params"]
    if not isinstance(mm, str):
        arrays["mass"] = np.asarray(mm, dtype=float)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
    c = out["counters"]
    print(f"{name}: B={ens['y0'].shape[0]} status {np.bincount(out['status'] + 1).tolist()} "
          f"mean acc/rej/fail {c[:, 4].mean():.1f}/{c[:, 5].mean():.1f}/{c[:, 6].mean():.1f}")


def pendulum_events():
    """Event functions of the events fixture / tests: x crosses zero (both directions, not terminal) and the
    vertical velocity crosses zero upwards, terminal at its second occurrence."""
    def x_zero(t, y):
        return y[0]

    def vy_up(t, y):
        return y[3]
    vy_up.direction = 1.0
    vy_up.terminal = 2
    return [x_zero, vy_up]


def save_events(name, ens):
    """scipy.integrate.solve_ivp(method=RadauDAE, events=...) of the reference, system by system."""
    if ONLY and not name.startswith(ONLY):
        return
    from scipy.integrate import solve_ivp
    opts = dict(ens["options"])
    mm, vi = opts.pop("mass_matrix"), opts.pop("var_index")
    B, n = ens["y0"].shape
    events = pendulum_events()
    tev = np.full((B, len(events), 16), np.nan)
    status = np.zeros(B, dtype=np.int32)
    t_final = np.zeros(B)
    y_final = np.zeros((B, n))
    n_eval = np.zeros(B, dtype=np.int32)
    for i in range(B):
        fun, jac, mass, _ = build_system(ens["problem"], ens["params"][i], n)
        with contextlib.redirect_stdout(io.StringIO()):
            r = solve_ivp(fun, ens["t_span"], ens["y0"][i], method=RadauDAE, t_eval=ens["t_eval"], events=events, jac=jac,
                          mass_matrix=mass, var_index=np.asarray(vi, dtype=float), dense_output=True, **opts)
        for e, te in enumerate(r.t_events):
            tev[i, e, :len(te)] = te
        status[i] = r.status
        n_eval[i] = r.t.size
        t_final[i] = r.sol.ts[-1] if r.status != 1 else r.t_events[1][-1]
        y_final[i] = r.sol(t_final[i])
    meta = dict(problem=ens["problem"], t_span=list(ens["t_span"]),
                options={k: v for k, v in opts.items()}, mass_matrix="problem")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), y0=ens["y0"], params=ens["params"], t_eval=ens["t_eval"],
                        var_index=np.asarray(vi, dtype=np.int32), t_events=tev, status=status, t_final=t_final, y_final=y_final,
                        n_eval=n_eval, meta=np.array(json.dumps(meta)))
    print(f"{name}: status {status.tolist()} events/system {np.isfinite(tev).sum(axis=2).tolist()}")


def single(problem, y0, params, t_span, t_eval, options):
    return dict(problem=problem, t_span=tuple(t_span), y0=np.atleast_2d(np.asarray(y0, dtype=float)),
                params=None if params is None else np.atleast_2d(np.asarray(params, dtype=float)),
                t_eval=np.asarray(t_eval, dtype=float), options=options)


def main():
    os.makedirs(GOLD, exist_ok=True)
    E = ensembles
    from dae_scipy_b200 import problems as P
    # --- well-conditioned known answers of SURVEY 4.2 (G1, G2, G3)
    pend = P.PROBLEMS["pendulum3"]
    save("g1_pendulum3_single", single("pendulum3", prob.pendulum_ic(np.pi / 4, 0.), pend.param_defaults, (0, 2),
                                       np.linspace(0, 2, 21),
                                       dict(rtol=1e-6, atol=1e-6, first_step=1e-2, mass_matrix="problem",
                                            var_index=pend.var_index_array())))
    rob = P.PROBLEMS["robertson"]
    save("g2_robertson_single", single("robertson", prob.ROBERTSON_Y0, rob.param_defaults, (0, 40), [0.4, 4., 40.],
                                       dict(rtol=1e-6, atol=1e-8, first_step=1e-8, max_bad_ite=0,
                                            scale_newton_norm=False, mass_matrix="problem",
                                            var_index=rob.var_index_array())))
    tra = E.transistor_ensemble(1, t_end=0.2)
    tra["params"] = P.PROBLEMS["transistor"].default_params(1)
    tra["y0"] = P.transistor_initial_state(tra["params"])
    save("g3_transistor_single", tra)
    # --- reference script configurations
    save("script_pendulum3", single("pendulum3", prob.pendulum_ic(np.pi / 2, 0., r0=3.), (1.0, 9.81, 3.0), (0, 10),
                                    np.linspace(0, 10, 51),
                                    dict(rtol=1e-8, atol=1e-8, newton_tol=1e-2, mass_matrix="problem",
                                         var_index=pend.var_index_array())))
    save("script_robertson", single("robertson", prob.ROBERTSON_Y0, rob.param_defaults, (0, 5e6),
                                    np.logspace(-5, np.log10(5e6), 16).clip(max=5e6),
                                    dict(rtol=1e-3, atol=1e-4, first_step=1e-8, max_newton_ite=6, max_bad_ite=0,
                                         scale_residuals=True, scale_newton_norm=False, scale_error=True,
                                         mass_matrix="problem", var_index=rob.var_index_array())))
    # --- seeded ensembles (first systems of the BASELINE configurations)
    save("ens_pendulum3_tend2", E.pendulum_ensemble(64, t_end=2.0))
    save("ens_pendulum3_tend10", E.pendulum_ensemble(16, t_end=10.0))
    save("ens_pendulum2_tend2", E.pendulum_ensemble(16, index=2))
    save("ens_pendulum1_tend2", E.pendulum_ensemble(16, index=1))
    save("ens_robertson", E.robertson_ensemble(32))
    save("ens_transistor_t0p05", E.transistor_ensemble(8, t_end=0.05))
    # --- option surface / edge cases on the pendulum
    e = E.pendulum_ensemble(8, t_end=1.0)
    save("opt_pendulum3_controller_off", e, dict(bUsePredictiveController=False, bAlwaysApply2ndEstimate=False,
                                                 bUsePredictiveNewtonStoppingCriterion=False, zero_algebraic_error=True,
                                                 max_newton_ite=8, rtol=1e-4))
    save("opt_pendulum3_no_scaling", e, dict(scale_residuals=False, scale_newton_norm=False, scale_error=False,
                                             bUseExtrapolatedGuess=False, rtol=1e-4, first_step=1e-3))
    save("opt_pendulum3_max_step", e, dict(max_step=0.02))
    save("opt_pendulum3_default_first_step", E.pendulum_ensemble(8, t_end=0.25), dict(first_step=None))
    save("opt_pendulum3_nmax_step", e, dict(nmax_step=25))
    save("opt_pendulum3_dense_mass", e, dict(mass_matrix=np.diag([1., 1., 1., 1., 0.]) * 2.0))
    back = E.pendulum_ensemble(8, t_end=1.0)
    back["t_span"] = (1.0, 0.0)
    back["t_eval"] = np.linspace(1.0, 0.0, 11)
    save("opt_pendulum3_backward", back)
    # --- the reference's own known-answer problems, analytic Jacobians
    jay_opts = dict(rtol=1e-3, atol=1e-4, first_step=1e-5, max_newton_ite=6, min_factor=0.2, max_factor=10,
                    zero_algebraic_error=True, scale_residuals=False, scale_newton_norm=False, scale_error=True,
                    max_bad_ite=1, bUsePredictiveNewtonStoppingCriterion=False, mass_matrix="problem",
                    var_index=np.array([0, 0, 0, 0, 3]))                       # jay_dae.py:76-101
    save("kat_jay", dict(problem="jay", t_span=(0.0, 0.1), y0=np.ones((2, 5)), params=np.array([[1.0], [0.0]]),
                         t_eval=np.linspace(0, 0.1, 11), options=jay_opts))
    poly_opts = dict(rtol=1e-4, atol=1e-5, first_step=0.25, max_step=0.25, min_factor=0.2, max_factor=10, max_newton_ite=6,
                     zero_algebraic_error=True, scale_residuals=True, scale_newton_norm=False, scale_error=True,
                     max_bad_ite=2, bUsePredictiveNewtonStoppingCriterion=False, bPerformAllNewtonIterations=True,
                     mass_matrix="problem", var_index=np.array([0, 2]))        # validation_dense_output.py:77-105
    a_full = np.array([0.01, 0.02, .05, -0.1, -0.4, 0.5, 1., 2.])
    P_rows, y0_rows = [], []
    for deg in range(0, 5):
        a = np.flip(np.flip(a_full)[:deg + 1])
        a5 = np.concatenate([np.zeros(5 - a.size), a])
        P_rows.append(a5); y0_rows.append([a[-1], a[-1]])
    save("kat_polydae2", dict(problem="polydae2", t_span=(0.0, 2.0), y0=np.array(y0_rows), params=np.array(P_rows),
                              t_eval=np.linspace(0, 2, 41), options=poly_opts))
    lin_opts = dict(rtol=1e-10, atol=1e-10, var_index=None)                    # linear_ODE_with_mass_matrix.py:26-48
    eig = prob.LINEAR7_EIG
    save("kat_linear7_with_mass", dict(problem="linear7", t_span=(0.0, 1.0), y0=np.ones((1, 7)), params=np.ones((1, 7)),
                                       t_eval=np.array([0.5, 1.0]), options=dict(lin_opts, mass_matrix=np.diag(1 / eig))))
    save("kat_linear7_without_mass", dict(problem="linear7", t_span=(0.0, 1.0), y0=np.ones((1, 7)), params=eig[None, :],
                                          t_eval=np.array([0.5, 1.0]), options=dict(lin_opts, mass_matrix=np.eye(7))))
    rob_script = dict(rtol=1e-3, atol=1e-4, first_step=1e-8, max_newton_ite=6, max_bad_ite=0, scale_residuals=True,
                      scale_newton_norm=False, scale_error=True)                # robertson.py:105-132
    save("kat_robertson_ode", dict(problem="robertson_ode", t_span=(0.0, 5e6), y0=prob.ROBERTSON_Y0[None, :],
                                   params=np.array([[0.04, 1e4, 3e7]]), t_eval=np.array([1.0, 1e3, 5e6]),
                                   options=dict(rob_script, mass_matrix="problem", var_index=np.array([0, 0, 0]))))
    # --- the reference's default Jacobian mode: jac=None, scipy num_jac finite differences (radauDAE.py:468-481)
    save("fd_pendulum3", E.pendulum_ensemble(8, t_end=1.0), dict(jac=None))
    save("fd_pendulum2", E.pendulum_ensemble(8, t_end=1.0, index=2), dict(jac=None))
    save("fd_robertson", E.robertson_ensemble(8, t_end=1e3), dict(jac=None))
    save("fd_transistor", E.transistor_ensemble(4, t_end=0.05), dict(jac=None))
    # the reference's OWN chain configuration: jac=None with the banded jac_sparsity (recursive_pendulum.py:128, 165-166)
    save("fd_npendulum_n4", E.npendulum_ensemble(3, n_nodes=4), dict(jac=None))
    save("fd_npendulum_n20", E.npendulum_ensemble(2, n_nodes=20), dict(jac=None))
    # --- a user problem (plugin path): the batch reactor of the reference's tests/reactor.py with the script's
    #     options, analytic Jacobian and the script's own jac=None
    save("user_reactor", E.reactor_ensemble(8))
    save("user_reactor_fd", E.reactor_ensemble(8), dict(jac=None))
    # --- a user BAND problem (48 unknowns, tridiagonal Jacobian by finite differences with jac_sparsity)
    save("user_rod", E.rod_ensemble(6))
    # --- events (scipy's solve_ivp driving the reference solver)
    save_events("events_pendulum3", E.pendulum_ensemble(6, t_end=3.0, n_t_eval=31))
    # --- n-pendulum chains (dense analytic Jacobian handed to the reference)
    save("ens_npendulum_n4", E.npendulum_ensemble(3, n_nodes=4))
    save("ens_npendulum_n10", E.npendulum_ensemble(2, n_nodes=10))
    save("ens_npendulum_n20", E.npendulum_ensemble(2, n_nodes=20))
    # the chain sizes of the bench configurations (BASELINE configs[4]: n = 20..100), first ropes of the bench ensembles,
    # full t_end (the reference factorises dense 250 x 250 / 500 x 500 matrices here: minutes per rope)
    save("ens_npendulum_n50", E.npendulum_ensemble(4, n_nodes=50))
    save("ens_npendulum_n100", E.npendulum_ensemble(4, n_nodes=100))


if __name__ == "__main__":
    main()
