"""The K1 kernel SOURCE (dae_scipy_b200/csrc/radau_k1.cuh), compiled for the host by
tests/host_sim and run as a warp of one lane, against the C oracle: the two are independent
implementations of the same arithmetic contract and must agree bit for bit.  This is the
no-GPU stand-in for tests/test_gpu_parity.py (which runs the real kernels)."""
import numpy as np
import pytest

import dae_scipy_b200 as br
from _util import assert_bit_identical, golden_names, load_golden, run_c_oracle, run_kernel_source_on_cpu

E = br.ensembles

CASES = {
    "pendulum3_t2": (lambda: E.pendulum_ensemble(96), {}),
    "pendulum3_t10": (lambda: E.pendulum_ensemble(16, t_end=10.0), {}),
    "pendulum2": (lambda: E.pendulum_ensemble(32, index=2), {}),
    "pendulum1": (lambda: E.pendulum_ensemble(32, index=1), {}),
    "pendulum3_default_first_step": (lambda: E.pendulum_ensemble(16, t_end=0.5), dict(first_step=None)),
    "pendulum3_dense_mass": (lambda: E.pendulum_ensemble(16), dict(mass_matrix=np.diag([1., 1, 1, 1, 0]) * 1.5)),
    "pendulum3_identity_mass": (lambda: E.pendulum_ensemble(8, t_end=0.2), dict(mass_matrix=None, var_index=None)),
    "pendulum3_controller_off": (lambda: E.pendulum_ensemble(16), dict(
        zero_algebraic_error=True, bAlwaysApply2ndEstimate=False, bUsePredictiveController=False,
        bUsePredictiveNewtonStoppingCriterion=False, max_newton_ite=8, rtol=1e-4)),
    "pendulum3_failing": (lambda: E.pendulum_ensemble(16), dict(
        bUseExtrapolatedGuess=False, scale_residuals=False, scale_newton_norm=False, scale_error=False,
        bPerformAllNewtonIterations=True, newton_tol=1e-3, rtol=1e-4)),
    "pendulum3_nmax_step": (lambda: E.pendulum_ensemble(16), dict(nmax_step=40)),
    "pendulum3_max_step": (lambda: E.pendulum_ensemble(16), dict(max_step=0.02)),
    "pendulum3_constant_dt": (lambda: E.pendulum_ensemble(8, t_end=0.5), dict(constant_dt=True, max_step=0.01, first_step=0.01, newton_tol=1e-8)),
    "pendulum3_constant_jac": (lambda: E.pendulum_ensemble(4, t_end=0.3), dict(
        jac=np.array([[0, 0, 1., 0, 0], [0, 0, 0, 1., 0], [-5., 0, 0, 0, -0.5], [0, -5., 0, 0, 0.8], [1., -1.6, 0, 0, 0]]))),
    "robertson": (lambda: E.robertson_ensemble(96), {}),
    "robertson_script_tol": (lambda: E.robertson_ensemble(48), dict(rtol=1e-3, atol=1e-4)),
    "robertson_vector_atol": (lambda: E.robertson_ensemble(16), dict(atol=np.array([1e-8, 1e-12, 1e-8]))),
    "transistor": (lambda: E.transistor_ensemble(6, t_end=0.02), {}),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_kernel_source_bit_identical_to_c_oracle(case):
    make, over = CASES[case]
    ens = make()
    jac = over.pop("jac", None) if isinstance(over, dict) else None
    over = dict(over)
    if case == "pendulum3_constant_jac":
        jac = CASES[case][1].get("jac", jac)
    sim = run_kernel_source_on_cpu(ens, **({"jac": jac} if jac is not None else {}), **{k: v for k, v in over.items() if k != "jac"})
    orc = run_c_oracle(ens, **({"jac_const": jac} if jac is not None else {}), **{k: v for k, v in over.items() if k != "jac"})
    assert_bit_identical(sim, orc, case)


def test_backward_integration():
    ens = E.pendulum_ensemble(8)
    ens["t_span"] = (2.0, 0.0)
    ens["t_eval"] = np.linspace(2.0, 0.0, 9)
    assert_bit_identical(run_kernel_source_on_cpu(ens), run_c_oracle(ens), "backward")


def test_t0_equals_t_bound_returns_initial_state():
    ens = E.pendulum_ensemble(4)
    ens["t_span"] = (1.0, 1.0)
    ens["t_eval"] = np.array([1.0])
    ens["options"] = dict(ens["options"], first_step=None)
    sim = run_kernel_source_on_cpu(ens)
    assert (sim["status"] == 0).all() and np.array_equal(sim["y"][:, :, 0], ens["y0"])
    assert_bit_identical(sim, run_c_oracle(ens), "t0 == t_bound")


@pytest.mark.parametrize("name", [n for n in golden_names() if "npendulum" not in n and n not in ("script_pendulum3", "g3_transistor_single")])
def test_kernel_source_on_reference_fixture_inputs(name):
    ens, _ = load_golden(name)
    assert_bit_identical(run_kernel_source_on_cpu(ens), run_c_oracle(ens), name)
