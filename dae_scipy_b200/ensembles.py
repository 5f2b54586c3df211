"""Seeded synthetic ensembles for the BASELINE.json configurations.

Every generator returns a dict with the arguments of `solve_ivp_batched`:
`problem, t_span, y0 [B,n], params [B,p] | None, t_eval, options` — numpy only, so that the
GPU path, the oracles and the bench all integrate exactly the same inputs.
The first `k` systems of an ensemble do not depend on B (the draws are made for the whole
requested size with `numpy.random.default_rng(seed)` and are sliced), which is what lets a
parity subsample of a 1M-system run be re-integrated on the CPU.
"""
from __future__ import annotations

import numpy as np

from . import problems as P


def pendulum_ensemble(B, t_end=2.0, n_t_eval=21, seed=0, index=3, rtol=1e-6, atol=1e-6, first_step=1e-2):
    """C2: index-3 pendulum, theta0 ~ U(0.1, 3.0), theta_dot0 ~ U(-1, 1), r0 = m = 1, g = 9.81."""
    rng = np.random.default_rng(seed)
    draws = rng.random((B, 2))
    theta0 = 0.1 + 2.9 * draws[:, 0]
    thetad0 = -1.0 + 2.0 * draws[:, 1]
    name = f"pendulum{index}"
    prob = P.PROBLEMS[name]
    return dict(problem=name, t_span=(0.0, float(t_end)), y0=P.pendulum_initial_state(theta0, thetad0),
                params=prob.default_params(B), t_eval=np.linspace(0.0, t_end, n_t_eval),
                options=dict(rtol=rtol, atol=atol, first_step=first_step, mass_matrix="problem",
                             var_index=prob.var_index_array()))


def robertson_ensemble(B, t_end=5e6, n_t_eval=16, seed=1, rtol=1e-6, atol=1e-8):
    """C3: Robertson DAE, k_i = k_i0 * 10**U(-0.5, 0.5)."""
    rng = np.random.default_rng(seed)
    k0 = np.array([0.04, 1e4, 3e7])
    params = k0[None, :] * 10.0 ** (rng.random((B, 3)) - 0.5)
    prob = P.PROBLEMS["robertson"]
    t_eval = np.logspace(-5, np.log10(t_end), n_t_eval)
    t_eval[-1] = min(t_eval[-1], t_end)
    return dict(problem="robertson", t_span=(0.0, float(t_end)), y0=P.robertson_initial_state(B),
                params=params, t_eval=t_eval,
                options=dict(rtol=rtol, atol=atol, first_step=1e-8, max_bad_ite=0, scale_newton_norm=False,
                             mass_matrix="problem", var_index=prob.var_index_array()))


def transistor_ensemble(B, t_end=0.2, n_t_eval=21, seed=2):
    """C4: transistor amplifier, C1..C3 and R0..R5 each scaled by 10**U(-0.1, 0.1); script options
    of transistor_amplifier.py:78-88."""
    rng = np.random.default_rng(seed)
    prob = P.PROBLEMS["transistor"]
    params = prob.default_params(B)
    params[:, :9] *= 10.0 ** (0.2 * rng.random((B, 9)) - 0.1)
    return dict(problem="transistor", t_span=(0.0, float(t_end)), y0=P.transistor_initial_state(params),
                params=params, t_eval=np.linspace(0.0, t_end, n_t_eval),
                options=dict(rtol=1e-5, atol=1e-5, first_step=1e-8, scale_residuals=False,
                             scale_newton_norm=False, scale_error=True, zero_algebraic_error=False,
                             max_bad_ite=1, mass_matrix="problem", var_index=None))


def npendulum_ensemble(B, n_nodes=20, n_t_eval=21, seed=3, t_end=None):
    """C5: n-pendulum chain, initial_angle ~ U(pi/8, 3pi/8); script options of
    recursive_pendulum.py:233-246 with the analytic Jacobian."""
    rng = np.random.default_rng(seed)
    ang = np.pi / 8 + (np.pi / 4) * rng.random(B)
    prob = P.PROBLEMS["npendulum"]
    tf = prob.t_span[1] if t_end is None else float(t_end)
    n = 5 * n_nodes
    return dict(problem="npendulum", t_span=(0.0, tf), y0=P.npendulum_initial_state(n_nodes, ang), params=None,
                t_eval=np.linspace(0.0, tf, n_t_eval),
                options=dict(rtol=1e-5, atol=1e-5, first_step=1e-3, max_step=prob.t_span[1] / 10, max_newton_ite=7,
                             max_bad_ite=2, scale_residuals=True, scale_newton_norm=True, scale_error=True,
                             mass_matrix="problem", var_index=prob.var_index_array(n)))


# ---------------------------------------------------------------- a USER problem (dae_scipy_b200/plugin.py)
# The batch-reactor DAE of the reference's tests/reactor.py:74-133 is not compiled into the stock library: it is
# the worked example of the plugin path -- three lines of CUDA C++ instead of the script's Python closure.
REACTOR_RHS = """f[0] = -p[0] * y[0] * y[1];
                 f[1] = y[1] - (p[2] - (p[1] - y[0]));
                 f[2] = y[2] - (p[1] - y[0]);"""
REACTOR_JAC = """J[0][0] = -p[0] * y[1]; J[0][1] = -p[0] * y[0];
                 J[1][0] = -1.0; J[1][1] = 1.0;
                 J[2][0] = 1.0; J[2][2] = 1.0;"""


def reactor_problem(analytic_jacobian=True):
    """Compile (once, cached) and register the reactor functor; returns its `Problem`.  Without the Jacobian body
    the plugin has finite differences only, like the reference script (jac=None)."""
    from .plugin import compile_problem
    return compile_problem("reactor" if analytic_jacobian else "reactor_fd", 3, REACTOR_RHS,
                           REACTOR_JAC if analytic_jacobian else None, param_names=("k", "Ca0", "Cb0"),
                           param_defaults=(1.47, 0.0209, 0.0209 / 3), mass_diag=(1, 0, 0), var_index=(0, 1, 1),
                           t_span=(0.0, 20.0))


def reactor_ensemble(B, t_end=20.0, n_t_eval=21, seed=4):
    """k = 1.47 * 10**U(-0.5, 0.5), Ca0 = 0.0209 * 10**U(-0.3, 0.3), Cb0 = Ca0 / 3 * 10**U(-0.2, 0.2); script
    options of reactor.py:108-133.  `problem` is the NAME "reactor": call `reactor_problem()` first on the paths
    that need the device functor."""
    rng = np.random.default_rng(seed)
    u = rng.random((B, 3))
    k = 1.47 * 10.0 ** (u[:, 0] - 0.5)
    Ca0 = 0.0209 * 10.0 ** (0.6 * u[:, 1] - 0.3)
    Cb0 = Ca0 / 3 * 10.0 ** (0.4 * u[:, 2] - 0.2)
    y0 = np.stack([Ca0, Cb0, np.zeros(B)], axis=1)
    return dict(problem="reactor", t_span=(0.0, float(t_end)), y0=y0, params=np.stack([k, Ca0, Cb0], axis=1),
                t_eval=np.linspace(0.0, t_end, n_t_eval),
                options=dict(rtol=1e-6, atol=1e-7, first_step=1e-8, max_newton_ite=20, max_bad_ite=3, min_factor=0.2,
                             max_factor=10, scale_residuals=False, scale_newton_norm=False, scale_error=False,
                             mass_matrix="problem", var_index=np.array([0, 1, 1], dtype=np.int32)))


# ---------------------------------------------------------------- a USER BAND problem (more than 8 unknowns, CTA-per-system kernel)
# A reaction-diffusion rod by the method of lines with algebraic boundary rows (an index-1 DAE, tridiagonal Jacobian):
#   0 = u_0 - (1 + a sin(w t));   u_i' = D (n-1)^2 (u_{i-1} - 2 u_i + u_{i+1}) - k u_i^3;   0 = u_{n-1} - u_{n-2}
ROD_N = 48
ROD_RHS_ROW = """if (i == 0) return y[0] - (1.0 + p[2] * sin(p[3] * t));
                 if (i == n - 1) return y[n - 1] - y[n - 2];
                 const double s = (double)(n - 1) * (double)(n - 1);
                 return (p[0] * s) * ((y[i - 1] - 2.0 * y[i]) + y[i + 1]) - p[1] * ((y[i] * y[i]) * y[i]);"""


def rod_problem():
    """Compile (once, cached) and register the rod as a band problem (kl = ku = 1, jac=None); returns its `Problem`."""
    from .plugin import compile_band_problem
    diag = (0,) + (1,) * (ROD_N - 2) + (0,)
    return compile_band_problem("rod", ROD_N, 1, 1, ROD_RHS_ROW, mass_row="i != 0 && i != n - 1", mass_diag=diag,
                                param_names=("D", "k", "a", "w"), param_defaults=(0.05, 2.0, 0.5, 6.0), t_span=(0.0, 2.0))


def rod_ensemble(B, t_end=2.0, n_t_eval=11, seed=5):
    """D = 0.05 * 10**U(-0.5, 0.5), k = 2 * 10**U(-0.5, 0.5), a = U(0.2, 0.8), w = U(3, 9); u(0) = 1 (consistent)."""
    rng = np.random.default_rng(seed)
    u = rng.random((B, 4))
    params = np.stack([0.05 * 10.0 ** (u[:, 0] - 0.5), 2.0 * 10.0 ** (u[:, 1] - 0.5), 0.2 + 0.6 * u[:, 2], 3.0 + 6.0 * u[:, 3]], axis=1)
    vi = np.zeros(ROD_N, dtype=np.int32)
    vi[0] = vi[-1] = 1
    return dict(problem="rod", t_span=(0.0, float(t_end)), y0=np.ones((B, ROD_N)), params=params,
                t_eval=np.linspace(0.0, t_end, n_t_eval),
                options=dict(rtol=1e-6, atol=1e-8, first_step=1e-6, jac=None, mass_matrix="problem", var_index=vi))


ENSEMBLES = {"pendulum": pendulum_ensemble, "robertson": robertson_ensemble,
             "transistor": transistor_ensemble, "npendulum": npendulum_ensemble}
