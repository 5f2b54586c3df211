"""The kernel SOURCES (thread-per-system K1, dae_scipy_b200/csrc/radau_k1.cuh; lane-per-rope chain kernel K4,
csrc/chain_k4.cuh), compiled for the host by tests/host_sim and run as a warp of one lane, against the C oracle: the
two are independent implementations of the same arithmetic contract and must agree bit for bit.  This is the
no-GPU stand-in for tests/test_gpu_parity.py (which runs the real kernels)."""
import numpy as np
import pytest

import dae_scipy_b200 as br
from _util import (PARITY_FACTOR, assert_bit_identical, assert_reports_identical, count_agreement, golden_names, load_golden,
                   run_c_oracle, run_chain_kernel_source_on_cpu, run_kernel_source_on_cpu, scaled_error)

E = br.ensembles

CASES = {
    "pendulum3_t2": (lambda: E.pendulum_ensemble(96), {}),
    "pendulum3_t10": (lambda: E.pendulum_ensemble(16, t_end=10.0), {}),
    "pendulum2": (lambda: E.pendulum_ensemble(32, index=2), {}),
    "pendulum1": (lambda: E.pendulum_ensemble(32, index=1), {}),
    "pendulum3_long_t_eval": (lambda: E.pendulum_ensemble(8, n_t_eval=401), {}),      # more points than the kernel stages on chip
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


def _sim_counts(ens, **over):
    """Run the kernel source and return how often its fast paths and their fall-backs ran (tests/host_sim counters):
    [optimistic LU attempts, abandoned ones, general LUs, Newton passes on static addresses, on pivoted ones,
    first-iteration convergence tests]."""
    import ctypes as C
    from _util import _host_sim
    sim = _host_sim(br.get_problem(ens["problem"]))
    sim.sim_counts.restype = C.POINTER(C.c_long)
    c = sim.sim_counts()
    for i in range(8):
        c[i] = 0
    out = run_kernel_source_on_cpu(ens, **over)
    return out, [c[i] for i in range(6)]


def test_fast_paths_and_their_fallbacks_are_exercised_and_bit_identical():
    """K1 fast paths (profiles/r01_k1_fastpaths.md) never change a bit: the index-3 pendulum and Robertson never pivot
    (optimistic LU and static substitutions throughout), the index-1 / index-2 pendulums and a mass matrix that shifts
    the pivots abandon the optimistic LU and run the general code, the transistor (traits off) never tries it."""
    out, c = _sim_counts(E.pendulum_ensemble(48))
    assert c[0] > 0 and c[1] == 0 and c[2] == 0 and c[3] > 0 and c[4] == 0 and c[5] > 0
    out, c = _sim_counts(E.robertson_ensemble(48))
    assert c[0] > 0 and c[1] == 0 and c[2] == 0 and c[4] == 0 and c[5] > 0
    ens = E.pendulum_ensemble(32, index=2)
    out, c = _sim_counts(ens)
    assert c[1] > 0 and c[2] > 0 and c[3] > 0 and c[4] > 0           # both kinds of factorisation and of substitution
    assert_bit_identical(out, run_c_oracle(ens), "pendulum2: mixed static / general")
    ens = E.pendulum_ensemble(48)
    M = np.diag([0.05, 0.05, 1.0, 1.0, 0.0])                          # pivots move off the diagonal, systems still finish
    out, c = _sim_counts(ens, mass_matrix=M)
    assert c[1] == 48 and c[0] == 48 and c[2] > 0 and c[4] > 0 and (out["status"] == 0).all()
    assert_bit_identical(out, run_c_oracle(ens, mass_matrix=M), "pendulum3 with shifted pivots")
    out, c = _sim_counts(E.transistor_ensemble(4, t_end=0.02))
    assert c[0] == 0 and c[5] == 0 and c[2] > 0 and c[4] > 0


def test_kernel_source_random_option_combinations():
    """Seeded differential run over the option surface (RadauDAE keywords of radauDAE.py:263-278 in random combinations,
    on all three problem families, analytic and finite-difference Jacobians, perturbed mass matrices): the kernel source and the
    C oracle must stay bit-identical whatever is switched on.  Step counts are capped so that pathological combinations
    (a method configured against itself takes millions of failed attempts) end quickly -- with the same status."""
    rng = np.random.default_rng(20261020)
    flags = ["scale_residuals", "scale_newton_norm", "scale_error", "zero_algebraic_error", "bAlwaysApply2ndEstimate",
             "bUsePredictiveController", "bUsePredictiveNewtonStoppingCriterion", "bUseExtrapolatedGuess", "bPerformAllNewtonIterations"]
    ranges = {"rtol": (-8, -3), "atol": (-10, -4), "max_step": (-3, 0), "newton_tol": (-8, -2), "jacobianRecomputeFactor": (-5, 0)}
    for case in range(24):
        kind = rng.choice(["pendulum3", "pendulum2", "pendulum1", "robertson", "transistor"])
        seed = int(rng.integers(1 << 30))
        if kind.startswith("pendulum"):
            ens = E.pendulum_ensemble(4, index=int(kind[-1]), t_end=float(rng.choice([0.3, 1.0, 2.0])), seed=seed)
        elif kind == "robertson":
            ens = E.robertson_ensemble(4, t_end=float(rng.choice([1e1, 1e3, 1e5])), seed=seed)
        else:
            ens = E.transistor_ensemble(2, t_end=0.01, seed=seed)
        over = {"nmax_step": int(rng.integers(100, 400))}
        for f in flags:
            if rng.random() < 0.3:
                over[f] = bool(rng.integers(2))
        for k, (lo, hi) in ranges.items():
            if rng.random() < 0.25:
                over[k] = float(10 ** rng.uniform(lo, hi))
        if rng.random() < 0.2: over["max_newton_ite"] = int(rng.integers(2, 12))
        if rng.random() < 0.2: over["max_bad_ite"] = int(rng.integers(0, 4))
        if rng.random() < 0.2: over["min_factor"] = float(rng.uniform(0.05, 0.5))
        if rng.random() < 0.2: over["max_factor"] = float(rng.uniform(2, 20))
        if rng.random() < 0.2: over["safety_factor"] = float(rng.uniform(0.5, 0.99))
        if rng.random() < 0.15: over["factor_on_non_convergence"] = float(rng.uniform(0.1, 0.8))
        if rng.random() < 0.15: over["jac"] = None
        if rng.random() < 0.15 and kind != "transistor":
            n = ens["y0"].shape[1]
            M = br.get_problem(ens["problem"]).mass().copy()
            M[:n - 1, :n - 1] += 0.1 * rng.standard_normal((n - 1, n - 1))
            over["mass_matrix"] = M
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert_bit_identical(run_kernel_source_on_cpu(ens, **over), run_c_oracle(ens, **over), f"case {case}: {kind} {over}")


# ---------------------------------------------------------------- the chain kernel (K4, block elimination)
CHAIN_CASES = [
    (2, 5, {}), (4, 7, {}), (10, 6, {}), (20, 4, {}), (50, 2, {}), (100, 2, {}),
    (4, 3, dict(nmax_step=30)), (6, 3, dict(rtol=1e-3, atol=1e-3, max_step=np.inf)),
    (3, 4, dict(zero_algebraic_error=True, bAlwaysApply2ndEstimate=False, bUsePredictiveController=False)),
    (5, 3, dict(scale_residuals=False, scale_newton_norm=False, scale_error=False, bUseExtrapolatedGuess=False)),
    (7, 3, dict(first_step=0.05, max_newton_ite=4, max_bad_ite=0)),                  # failed Newton attempts, stale-Jacobian retries
    (8, 2, dict(constant_dt=True, max_step=0.002, first_step=0.002, newton_tol=1e-8, nmax_step=150)),
    (5, 2, dict(rtol=1e-8, atol=1e-8, bPerformAllNewtonIterations=True, nmax_step=200)),
    (6, 2, dict(atol=np.tile([1e-5, 1e-6, 1e-4, 1e-5, 1e-3], 6), var_index=np.tile([0, 0, 1, 2, 3], 6))),   # per-unknown atol / indices
]


@pytest.mark.parametrize("nodes,B,over", CHAIN_CASES)
def test_chain_kernel_source_bit_identical_to_c_oracle(nodes, B, over):
    ens = E.npendulum_ensemble(B, n_nodes=nodes)
    assert_bit_identical(run_chain_kernel_source_on_cpu(ens, **over), run_c_oracle(ens, **over), f"chain n={nodes} {sorted(over)}")


def test_chain_kernel_source_backward_and_degenerate_spans():
    ens = E.npendulum_ensemble(3, n_nodes=5, t_end=1.0)
    ens["t_span"], ens["t_eval"] = (1.0, 0.0), np.linspace(1.0, 0.0, 7)
    assert_bit_identical(run_chain_kernel_source_on_cpu(ens), run_c_oracle(ens), "chain backward")
    ens = E.npendulum_ensemble(2, n_nodes=3)
    ens["t_span"], ens["t_eval"] = (0.5, 0.5), np.array([0.5])
    ens["options"] = dict(ens["options"], first_step=None)
    out = run_chain_kernel_source_on_cpu(ens)
    assert (out["status"] == 0).all() and np.array_equal(out["y"][:, :, 0], ens["y0"])
    assert_bit_identical(out, run_c_oracle(ens), "chain t0 == t_bound")


@pytest.mark.parametrize("name", [n for n in golden_names() if "npendulum" in n and not n.startswith("fd_")])   # K4: analytic Jacobian
def test_chain_kernel_source_against_reference_fixture(name):
    """The chain kernel's source reproduces the reference's own runs (fixtures generated by oracle/gen_golden.py from
    scipyDAE.radauDAE.RadauDAE, n = 4 ... 100 nodes): every value within tolerance and identical step counts."""
    ens, ref = load_golden(name)
    out = run_chain_kernel_source_on_cpu(ens)
    o = ens["options"]
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["n_eval"], ref["n_eval"])
    assert scaled_error(out["y"], ref["y"], o["rtol"], o["atol"]) <= PARITY_FACTOR
    assert count_agreement(out["counters"], ref["counters"]) == 1.0
    assert_bit_identical(out, run_c_oracle(ens), name)


def test_chain_dense_output_record_of_the_kernel_source():
    ens = E.npendulum_ensemble(3, n_nodes=6)
    out = run_chain_kernel_source_on_cpu(ens, dense_cap=600)
    sol = out["sol"]
    assert (sol.n_steps == out["counters"][:, 4]).all() and not sol.truncated.any()
    np.testing.assert_allclose(sol(ens["t_eval"]), out["y"], rtol=1e-8, atol=1e-10)
    last = sol.ys[np.arange(3), :, sol.n_steps]
    assert np.array_equal(last, out["y_final"])


# ---------------------------------------------------------------- per-attempt report (the reference's bReport=True)
@pytest.mark.parametrize("make,over,chain", [
    (lambda: E.pendulum_ensemble(12), {}, False),
    (lambda: E.pendulum_ensemble(6, index=2), dict(first_step=0.2, max_newton_ite=4, max_bad_ite=0), False),   # failed attempts, Jacobian updates
    (lambda: E.robertson_ensemble(12), {}, False),
    (lambda: E.transistor_ensemble(3, t_end=0.02), {}, False),
    (lambda: E.npendulum_ensemble(3, n_nodes=6), {}, True),
    (lambda: E.npendulum_ensemble(2, n_nodes=7), dict(first_step=0.05, max_newton_ite=4, max_bad_ite=0), True),
])
def test_per_attempt_report_of_the_kernel_sources_equals_the_c_oracle(make, over, chain):
    ens = make()
    run = run_chain_kernel_source_on_cpu if chain else run_kernel_source_on_cpu
    sim, orc = run(ens, report_cap=3000, **over), run_c_oracle(ens, report_cap=3000, **over)
    assert_bit_identical(sim, orc, "with the report switched on")
    assert_reports_identical(sim["reports"], orc["reports"], "kernel source vs C oracle")
    codes = set(np.unique(np.concatenate([orc["reports"]["code"][i, :k] for i, k in enumerate(orc["reports"]["n"])])).tolist())
    assert {0, -10} <= codes <= {0, 2, 5, -10, -11}
    if over:
        assert 5 in codes and -11 in codes                 # non-converged attempts and Jacobian updates were reported
    # a capped report keeps the first records and still counts all of them
    few = run(ens, report_cap=7, **over)["reports"]
    assert np.array_equal(few["n"], sim["reports"]["n"]) and np.array_equal(few["code"], sim["reports"]["code"][:, :7])


def test_per_attempt_report_reproduces_the_reference_report():
    """`RadauDAE(bReport=True).reports` of the unmodified reference against the restatement's record on a well-conditioned
    ensemble: the same events in the same order with the same Newton iteration counts; times, steps and error norms to
    rounding.  (Runs where the reference is importable: /root/reference or oracle/_ref.)"""
    import contextlib
    import io
    from oracle import problems_numpy as prob, reference_runner
    if reference_runner.reference_location() is None:
        pytest.skip("the reference is not importable here")
    RadauDAE = reference_runner.radau_class()
    ens = E.pendulum_ensemble(6)
    sim = run_kernel_source_on_cpu(ens, report_cap=3000)["reports"]
    opts = dict(ens["options"])
    opts.pop("mass_matrix")
    vi = opts.pop("var_index")
    for i in range(6):
        fun, jac, mass, _ = prob.build("pendulum3", ens["params"][i], 5)
        with contextlib.redirect_stdout(io.StringIO()):
            s = RadauDAE(fun, 0.0, ens["y0"][i].copy(), 2.0, jac=jac, mass_matrix=mass, var_index=np.asarray(vi, float), bReport=True, **opts)
            while s.status == "running":
                s.step()
        r, k = s.reports, int(sim["n"][i])
        assert k == len(r["code"])
        assert np.array_equal(np.array(r["code"]), sim["code"][i, :k])
        assert np.array_equal(np.array(r["newton_iterations"]), sim["newton_iterations"][i, :k])
        assert np.array_equal(np.array(r["bad_iterations"]), sim["bad_iterations"][i, :k])
        np.testing.assert_allclose(sim["t"][i, :k], np.array(r["t"], float), rtol=0, atol=1e-6)
        np.testing.assert_allclose(sim["dt"][i, :k], np.array(r["dt"], float), rtol=1e-4, atol=1e-9)
        acc = np.array(r["code"]) == 0
        np.testing.assert_allclose(sim["err1"][i, :k][acc], np.array(r["err1"], float)[acc], rtol=2e-3, atol=1e-12)


@pytest.mark.parametrize("make,p", [(lambda: E.pendulum_ensemble(96, index=2), 64), (lambda: E.pendulum_ensemble(96, index=2), 200),
                                    (lambda: E.pendulum_ensemble(48, index=1), 64), (lambda: E.pendulum_ensemble(48), 64),
                                    (lambda: E.robertson_ensemble(48), 64), (lambda: E.robertson_ensemble(64, rtol=1e-3, atol=1e-4), 100)])
def test_fast_paths_vetoed_by_a_neighbouring_lane_stay_bit_identical(make, p):
    """On the device the fast paths (optimistic LU, pattern-restricted substitutions) are warp-wide votes: a neighbouring lane
    that pivoted sends every lane of the warp through the general code for that block.  The harness (one lane per warp)
    emulates that neighbour at random; whatever mix of paths a system sees, its bits must not change.  (Regression: a system
    that pivoted inside a general factorisation forced by a neighbour went back to the pattern-restricted optimistic
    factorisation, which left the earlier fill outside the pattern in place -- GPU run of pendulum2_1024, round 2.)"""
    import ctypes as C
    from _util import _host_sim
    ens = make()
    sim = _host_sim(br.get_problem(ens["problem"]))
    ref = run_c_oracle(ens)
    try:
        for seed in (1, 2, 3):
            sim.sim_neighbour(C.c_uint(p), C.c_uint(seed))
            assert_bit_identical(run_kernel_source_on_cpu(ens), ref, f"neighbour veto p={p} seed={seed}")
    finally:
        sim.sim_neighbour(C.c_uint(0), C.c_uint(0))
