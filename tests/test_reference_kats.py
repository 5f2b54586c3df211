"""The reference's OWN assertions for this path, restated for every implementation:

  linear_ODE_with_mass_matrix.py:71-75  final state == y0 exp(eig t) to 10*max(atol, rtol) = 1e-9 relative, with
                                        and without the mass matrix, and step counts within 5 % of each other
  robertson.py:249-251                  DAE and ODE formulations both succeed and agree at t = 5e6 to
                                        5*sqrt(max(atol, rtol)) relative
  validation_dense_output.py:111,157-166 success, and the dense output of the algebraic z equals that of the
                                        differential y to 1e-12 relative on 1000 points for polynomial degree <= 3
  jay_dae.py:131, transistor_amplifier.py:91  success

Checkers: the numpy restatement, the C restatement, the K1 kernel source compiled for the host (no GPU),
and the CUDA kernels themselves (-m gpu)."""
import numpy as np
import pytest

import dae_scipy_b200 as br
from _util import run_c_oracle, run_gpu, run_kernel_source_on_cpu, run_numpy_oracle

CPU_RUNNERS = {"numpy_oracle": run_numpy_oracle, "c_oracle": run_c_oracle, "kernel_source_cpu": run_kernel_source_on_cpu}
EIG = np.array([float(i) * (-1) ** i for i in range(1, 8)])
A_FULL = np.array([0.01, 0.02, .05, -0.1, -0.4, 0.5, 1., 2.])


def linear_case(with_mass):
    opts = dict(rtol=1e-10, atol=1e-10, var_index=None)
    if with_mass:
        return dict(problem="linear7", t_span=(0.0, 1.0), y0=np.ones((1, 7)), params=np.ones((1, 7)),
                    t_eval=np.array([1.0]), options=dict(opts, mass_matrix=np.diag(1 / EIG)))
    return dict(problem="linear7", t_span=(0.0, 1.0), y0=np.ones((1, 7)), params=EIG[None, :],
                t_eval=np.array([1.0]), options=dict(opts, mass_matrix=np.eye(7)))


def check_linear(run):
    true_solution = np.exp(EIG * 1.0)
    a, b = run(linear_case(True)), run(linear_case(False))
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    np.testing.assert_allclose(b["y"][0, :, -1], true_solution, rtol=1e-9)
    np.testing.assert_allclose(a["y"][0, :, -1], true_solution, rtol=1e-9)
    np.testing.assert_allclose(a["counters"][0, 4], b["counters"][0, 4], rtol=0.05)


def robertson_cases():
    script = dict(rtol=1e-3, atol=1e-4, first_step=1e-8, max_newton_ite=6, max_bad_ite=0, scale_residuals=True,
                  scale_newton_norm=False, scale_error=True, mass_matrix="problem")
    base = dict(t_span=(0.0, 5e6), y0=np.array([[1., 0., 0.]]), params=np.array([[0.04, 1e4, 3e7]]), t_eval=np.array([5e6]))
    dae = dict(base, problem="robertson", options=dict(script, var_index=np.array([0, 0, 1])))
    ode = dict(base, problem="robertson_ode", options=dict(script, var_index=np.array([0, 0, 0])))
    return dae, ode


def check_robertson(run):
    dae, ode = robertson_cases()
    a, b = run(dae), run(ode)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    np.testing.assert_allclose(b["y"][0, :, -1], a["y"][0, :, -1], rtol=5 * 1e-3 ** 0.5)


def poly_case(n_points=1000):
    opts = dict(rtol=1e-4, atol=1e-5, first_step=0.25, max_step=0.25, min_factor=0.2, max_factor=10, max_newton_ite=6,
                zero_algebraic_error=True, scale_residuals=True, scale_newton_norm=False, scale_error=True,
                max_bad_ite=2, bUsePredictiveNewtonStoppingCriterion=False, bPerformAllNewtonIterations=True,
                mass_matrix="problem", var_index=np.array([0, 2]))
    P, y0 = [], []
    for deg in range(0, 5):
        a = np.flip(np.flip(A_FULL)[:deg + 1])
        P.append(np.concatenate([np.zeros(5 - a.size), a]))
        y0.append([a[-1], a[-1]])
    return dict(problem="polydae2", t_span=(0.0, 2.0), y0=np.array(y0), params=np.array(P),
                t_eval=np.linspace(0.0, 2.0, n_points), options=opts)


def check_poly(run):
    ens = poly_case()
    out = run(ens)
    assert (out["status"] == 0).all()
    y, z = out["y"][:, 0, :], out["y"][:, 1, :]
    rel = np.abs(z - y) / (1e-10 + np.abs(y))
    for deg in range(0, 4):
        assert rel[deg].max() < 1e-12, f"dense output not exact for degree {deg}"
        # and for these degrees the cubic interpolant reproduces the polynomial itself
        exact = np.polyval(ens["params"][deg], ens["t_eval"])
        assert np.abs(y[deg] - exact).max() < 1e-10


def check_smoke(run):
    jay = dict(problem="jay", t_span=(0.0, 0.1), y0=np.ones((2, 5)), params=np.array([[1.0], [0.0]]), t_eval=np.array([0.1]),
               options=dict(rtol=1e-3, atol=1e-4, first_step=1e-5, zero_algebraic_error=True, scale_residuals=False,
                            scale_newton_norm=False, scale_error=True, max_bad_ite=1,
                            bUsePredictiveNewtonStoppingCriterion=False, mass_matrix="problem",
                            var_index=np.array([0, 0, 0, 0, 3])))
    out = run(jay)
    assert (out["status"] == 0).all()
    assert np.allclose(out["y"][:, 0, -1] * out["y"][:, 1, -1] ** 2, 1.0, atol=1e-3)     # the constraint y1 y2^2 = 1
    tra = br.ensembles.transistor_ensemble(2, t_end=0.01)
    assert (run(tra)["status"] == 0).all()


CHECKS = {"linear_ode_with_mass": check_linear, "robertson_dae_vs_ode": check_robertson,
          "dense_output_polynomial": check_poly, "jay_and_transistor_succeed": check_smoke}


@pytest.mark.parametrize("runner", sorted(CPU_RUNNERS))
@pytest.mark.parametrize("check", sorted(CHECKS))
def test_reference_assertions_on_cpu(check, runner):
    CHECKS[check](CPU_RUNNERS[runner])


@pytest.mark.gpu
@pytest.mark.parametrize("check", sorted(CHECKS))
def test_reference_assertions_on_gpu(check):
    CHECKS[check](run_gpu)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["linear_with", "linear_without", "robertson_ode", "poly", "jay"])
def test_gpu_kat_problems_bit_identical_to_c_oracle(case):
    from _util import assert_bit_identical
    if case == "linear_with":
        ens = linear_case(True)
    elif case == "linear_without":
        ens = linear_case(False)
    elif case == "robertson_ode":
        ens = robertson_cases()[1]
        ens = dict(ens, y0=np.repeat(ens["y0"], 64, 0), params=ens["params"] * (1 + 0.01 * np.arange(64))[:, None])
    elif case == "poly":
        ens = poly_case(50)
    else:
        ens = dict(problem="jay", t_span=(0.0, 0.1), y0=np.ones((2, 5)), params=np.array([[1.0], [0.0]]),
                   t_eval=np.linspace(0, 0.1, 5), options=dict(rtol=1e-3, atol=1e-4, first_step=1e-5, mass_matrix="problem",
                                                               var_index=np.array([0, 0, 0, 0, 3]), max_bad_ite=1))
    assert_bit_identical(run_gpu(ens), run_c_oracle(ens), case)
