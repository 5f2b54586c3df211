"""The numpy restatement (oracle/radau_numpy.py) against the reference: committed fixtures that
oracle/gen_golden.py produced by running the reference itself, and — where /root/reference is
mounted — the live reference, bit for bit."""
import contextlib
import io
import sys
import warnings

import numpy as np
import pytest

from _util import (HAVE_REFERENCE, PARITY_FACTOR, count_agreement, golden_names, load_golden,
                   run_numpy_oracle, scaled_error, well_conditioned)

# the long chains (dense 250 x 250 / 500 x 500 LAPACK factorisations in the port) are pinned through the C oracle and the kernel
# source instead (tests/test_oracle_c.py, tests/test_kernel_source_cpu.py)
FAST = [n for n in golden_names() if n not in ("ens_npendulum_n20", "ens_npendulum_n50", "ens_npendulum_n100",
                                               "g3_transistor_single", "script_pendulum3")]


@pytest.mark.parametrize("name", FAST)
def test_numpy_oracle_matches_reference_fixture(name):
    """Same machine class as the fixture generator -> expected bit-identical; on a host whose BLAS
    picks another kernel family the well-conditioned fixtures must still meet the parity bar."""
    ens, ref = load_golden(name)
    out = run_numpy_oracle(ens)
    assert np.array_equal(out["status"], ref["status"])
    if np.array_equal(out["y"], ref["y"], equal_nan=True):
        assert np.array_equal(out["counters"], ref["counters"])
        return
    if well_conditioned(name):
        assert np.array_equal(out["n_eval"], ref["n_eval"])
        assert scaled_error(out["y"], ref["y"], ens["options"]["rtol"], ens["options"]["atol"]) <= PARITY_FACTOR
        assert count_agreement(out["counters"], ref["counters"]) >= 0.99


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is only mounted in the build container")
@pytest.mark.parametrize("case", ["pendulum3", "pendulum3_default_h0", "robertson", "transistor", "failing"])
def test_numpy_oracle_bit_identical_to_live_reference(case):
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    from scipy.integrate import solve_ivp
    from scipyDAE.radauDAE import RadauDAE
    from oracle import radau_numpy as orc, problems_numpy as prob
    if case.startswith("pendulum3"):
        fun, jac, mass, vi = prob.pendulum(3)
        y0 = prob.pendulum_ic(1.1, -0.3)
        span, te = (0, 1.5), np.linspace(0, 1.5, 16)
        opts = dict(rtol=1e-6, atol=1e-6) if case.endswith("h0") else dict(rtol=1e-6, atol=1e-6, first_step=1e-2)
    elif case == "robertson":
        fun, jac, mass, vi = prob.robertson(0.05, 2e4, 1e7)
        y0, span, te = prob.ROBERTSON_Y0, (0, 1e4), np.logspace(-4, 4, 9)
        opts = dict(rtol=1e-6, atol=1e-8, first_step=1e-8, max_bad_ite=0, scale_newton_norm=False)
    elif case == "transistor":
        fun, jac, mass, vi = prob.transistor()
        y0, span, te = prob.transistor_y0(), (0, 0.02), np.linspace(0, 0.02, 5)
        opts = dict(rtol=1e-5, atol=1e-5, first_step=1e-8, scale_residuals=False, scale_newton_norm=False, max_bad_ite=1)
    else:
        fun, jac, mass, vi = prob.pendulum(3)
        y0, span, te = prob.pendulum_ic(0.8, 0.), (0, 1), np.linspace(0, 1, 5)
        opts = dict(rtol=1e-4, atol=1e-6, first_step=1e-3, scale_residuals=False, scale_newton_norm=False,
                    scale_error=False, bUseExtrapolatedGuess=False)
    vif = None if vi is None else vi.astype(float)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = solve_ivp(fun, span, y0, method=RadauDAE, jac=jac, mass_matrix=mass, var_index=vif, t_eval=te, **opts)
    out = orc.solve(fun, span, y0, jac=jac, mass_matrix=mass, var_index=vif, t_eval=te, **opts)
    assert out.status == ref.status
    assert (out.counters.nfev, out.counters.njev, out.counters.nlu) == (ref.nfev, ref.njev, ref.nlu)
    ref_y = ref.y if np.ndim(ref.y) == 2 else np.empty((y0.size, 0))
    assert out.y.shape == ref_y.shape and np.array_equal(out.y, ref_y)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is only mounted in the build container")
def test_problem_restatements_return_reference_bits():
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    from scipyDAE.tests import pendulum as rp, robertson as rr, transistor_amplifier as rt, recursive_pendulum as rn
    from oracle import problems_numpy as prob
    rng = np.random.default_rng(5)
    for idx in (3, 2, 1):
        _, f_ref, j_ref, m_ref, x0, vi_ref = rp.generateSystem(idx, theta_0=0.7, theta_dot0=0.2)
        f, j, m, vi = prob.pendulum(idx)
        x = x0 + 0.1 * rng.standard_normal(5)
        assert np.array_equal(f(0.3, x), f_ref(0.3, x)) and np.array_equal(j(0.3, x), j_ref(0.3, x))
        assert np.array_equal(m, m_ref) and np.array_equal(vi, vi_ref)
        assert np.array_equal(prob.pendulum_ic(0.7, 0.2), x0)
    _, y0, f_ref, j_ref, m_ref, vi_ref = rr.generateSystem(True)
    f, j, m, vi = prob.robertson()
    y = np.array([0.9, 3e-5, 0.1])
    assert np.array_equal(f(0., y), f_ref(0., y)) and np.array_equal(j(0., y), j_ref(0., y)) and np.array_equal(m, m_ref)
    _, y0, f_ref, j_ref, m_ref, _, _ = rt.generateSystem()
    f, j, m, _ = prob.transistor()
    y = y0 + 0.01 * rng.standard_normal(5)
    assert np.array_equal(f(0.0013, y), f_ref(0.0013, y)) and np.array_equal(m, m_ref)
    assert np.allclose(j(0.0013, y), j_ref(0.0013, y), rtol=1e-12, atol=0)     # complex-step == analytic
    assert np.array_equal(prob.transistor_y0(), y0)
    for n in (3, 7):
        _, f_ref, _, _, m_ref, x0, vi_ref, _, _ = rn.generateSystem(n, np.pi / 5, 3)
        f, j, m, vi = prob.npendulum(n)
        x = x0 + 0.01 * rng.standard_normal(5 * n)
        assert np.array_equal(f(0., x), f_ref(0., x))
        assert np.array_equal(m, m_ref.toarray()) and np.array_equal(vi, vi_ref)
        assert np.array_equal(prob.npendulum_ic(n, np.pi / 5), x0)


@pytest.mark.parametrize("n", [2, 5])
def test_npendulum_analytic_jacobian_matches_complex_step(n):
    from oracle import problems_numpy as prob
    f, j, _, _ = prob.npendulum(n)
    rng = np.random.default_rng(n)
    x = prob.npendulum_ic(n, 0.6) + 0.05 * rng.standard_normal(5 * n)
    J = j(0., x)
    Jcs = np.empty_like(J)
    for k in range(5 * n):
        xp = x.astype(complex)
        xp[k] += 1e-30j
        Jcs[:, k] = f(0., xp).imag / 1e-30
    assert np.allclose(J, Jcs, rtol=1e-12, atol=1e-12)
