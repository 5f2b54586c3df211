"""User problems (dae_scipy_b200/plugin.py): the `fun` / `jac` closures of the reference as a device functor
compiled into a plugin build of the library.

Checked here without a GPU: the generated functor header, that the plugin library builds for sm_100a and
reports its problem through the C ABI, and the kernel SOURCE instantiated with the functor and run on the
CPU (tests/host_sim) against (a) fixtures produced by the reference itself for its batch-reactor script
(tests/reactor.py:74-133; tests/golden/user_reactor*.npz from oracle/gen_golden.py) and (b) the
reference-equivalent numpy oracle driven, live, by the same right-hand side written as a Python closure.
The GPU tests of the plugin path are in tests/test_gpu_parity.py.
"""
import ctypes as C

import numpy as np
import pytest

import _util as U
import dae_scipy_b200 as br
from dae_scipy_b200 import _abi, plugin
from dae_scipy_b200 import ensembles as E

VDP_RHS = """f[0] = y[1];
             f[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0] + p[1] * sin(p[2] * t);"""
VDP_JAC = """J[0][1] = 1.0;
             J[1][0] = -2.0 * p[0] * y[0] * y[1] - 1.0;
             J[1][1] = p[0] * (1.0 - y[0] * y[0]);"""


def vdp_problem():
    """Forced Van der Pol oscillator (stiff ODE, time-dependent right-hand side, libm call), parameters (mu, A, w)."""
    return br.compile_problem("vdp_forced", 2, VDP_RHS, VDP_JAC, param_names=("mu", "A", "w"),
                              param_defaults=(100.0, 0.0, 1.0), t_span=(0.0, 100.0))


def vdp_ensemble(B, t_end=100.0, seed=7):
    rng = np.random.default_rng(seed)
    u = rng.random((B, 3))
    params = np.stack([10.0 ** (1 + 1.5 * u[:, 0]), 0.5 * u[:, 1], 0.5 + u[:, 2]], axis=1)
    y0 = np.stack([np.full(B, 2.0), np.zeros(B)], axis=1)
    return dict(problem="vdp_forced", t_span=(0.0, t_end), y0=y0, params=params, t_eval=np.linspace(0.0, t_end, 21),
                options=dict(rtol=1e-6, atol=1e-8, first_step=1e-4, mass_matrix="problem", var_index=None))


def vdp_numpy_oracle(ens, jac=True):
    """The numpy restatement of the reference (bit-identical to RadauDAE here) with the SAME problem as closures."""
    from oracle import radau_numpy as orc
    B, n = ens["y0"].shape
    T = len(ens["t_eval"])
    o = dict(ens["options"])
    o.pop("mass_matrix"); o.pop("var_index")
    out = dict(y=np.full((B, n, T), np.nan), counters=np.zeros((B, 9), np.int32), status=np.zeros(B, np.int32))
    for i in range(B):
        mu, A, w = ens["params"][i]
        fun = lambda t, y: np.array([y[1], mu * (1.0 - y[0] * y[0]) * y[1] - y[0] + A * np.sin(w * t)])   # noqa: E731
        dfun = lambda t, y: np.array([[0.0, 1.0], [-2.0 * mu * y[0] * y[1] - 1.0, mu * (1.0 - y[0] * y[0])]])   # noqa: E731
        r = orc.solve(fun, ens["t_span"], ens["y0"][i], jac=dfun if jac else None, mass_matrix=np.eye(n), var_index=None,
                      t_eval=ens["t_eval"], **o)
        out["y"][i, :, :r.y.shape[1]] = r.y
        out["counters"][i] = r.counters.as_array()
        out["status"][i] = r.status
    return out


def test_generated_header_and_argument_checks():
    text = plugin.generate_header("reactor", 3, 3, E.REACTOR_RHS, E.REACTOR_JAC, mass_diag=(1, 0, 0))
    assert "static constexpr int N = 3, NP = 3, BLOCK = 384;" in text and "MassDiag01<N, 0x1u>" in text
    assert "f[0] = -p[0] * y[0] * y[1];" in text and "J[2][2] = 1.0;" in text
    assert plugin._block_for(5) == 256 and plugin._block_for(7) == 128 and plugin._block_for(8) == 96
    with pytest.raises(ValueError):
        plugin.generate_header("bad name", 3, 0, "f[0] = 0;")
    with pytest.raises(ValueError):
        plugin.generate_header("big", 9, 0, "f[0] = 0;")           # thread-per-system kernels: n <= 8
    with pytest.raises(ValueError):
        plugin.generate_header("nodiag", 3, 0, "f[0] = 0;", mass_diag=(1, 0))
    with pytest.raises(ValueError):
        plugin.generate_header("norhs", 3, 0, "  ")
    with pytest.raises(ValueError):
        br.compile_problem("pendulum3", 5, "f[0] = 0;")            # names of the stock library are taken


def test_plugin_library_builds_and_reports_its_problem():
    prob = E.reactor_problem()
    assert prob.problem_id == plugin.BRDAE_USER and prob.n == 3 and prob.n_params == 3 and prob.has_jac
    assert br.get_problem("reactor") is prob                       # selectable by name, like the stock functors
    lib = _abi.load(prob.lib_path)
    for sym in _abi.EXPORTED_SYMBOLS:                              # a complete libbrdae: the same C ABI
        getattr(lib, sym)
    assert lib.brdae_abi_version() == 4
    pid, n, npar = C.c_int32(-1), C.c_int32(-1), C.c_int32(-1)
    for name in (b"reactor", b"user"):
        assert lib.brdae_problem_lookup(name, C.byref(pid), C.byref(n), C.byref(npar)) == 0
        assert (pid.value, n.value, npar.value) == (100, 3, 3)
    assert lib.brdae_problem_lookup(b"pendulum3", C.byref(pid), C.byref(n), C.byref(npar)) == 0 and pid.value == 0
    stock = _abi.load()
    assert stock.brdae_problem_lookup(b"reactor", C.byref(pid), C.byref(n), C.byref(npar)) == 2     # unknown problem
    # the batch of a user problem is validated like any other (n / n_params of the functor)
    hb = br.prepare(prob, (0.0, 20.0), (4, 3), rtol=1e-6, atol=1e-7, first_step=1e-8)
    assert hb.mass is None and hb.var_index.tolist() == [0, 0, 0]      # var_index=None: all differential, as in the reference
    with pytest.raises(ValueError):
        br.prepare(prob, (0.0, 20.0), (4, 5))


@pytest.mark.parametrize("name", ["user_reactor", "user_reactor_fd"])
def test_user_kernel_source_on_cpu_against_the_reference_fixture(name):
    """The K1 kernel source instantiated with the reactor functor reproduces the reference run of the same script:
    trajectories within the north_star tolerance, identical accepted/rejected/failed counts."""
    E.reactor_problem()
    ens, ref = U.load_golden(name)
    out = U.run_kernel_source_on_cpu(ens)
    o = ens["options"]
    assert np.array_equal(out["status"], ref["status"]) and (out["status"] == 0).all()
    assert U.scaled_error(out["y"], ref["y"], o["rtol"], o["atol"]) <= U.PARITY_FACTOR
    assert U.count_agreement(out["counters"], ref["counters"]) == 1.0
    # and the numpy oracle, handed the closure form of the problem, is the reference bit for bit
    npo = U.run_numpy_oracle(ens)
    assert np.array_equal(npo["y"], ref["y"], equal_nan=True) and np.array_equal(npo["counters"], ref["counters"])


def test_user_problem_without_jacobian_body_uses_finite_differences():
    prob = E.reactor_problem(analytic_jacobian=False)
    assert not prob.has_jac
    hb = br.prepare(prob, (0.0, 20.0), (2, 3), rtol=1e-6, atol=1e-7, first_step=1e-8)
    assert hb.options.jac_finite_differences == 1                  # jac='analytic' default falls back to num_jac
    ens, ref = U.load_golden("user_reactor_fd")
    ens = dict(ens, problem=prob)
    ens["options"].pop("jac")
    out = U.run_kernel_source_on_cpu(ens)
    assert U.scaled_error(out["y"], ref["y"], 1e-6, 1e-7) <= U.PARITY_FACTOR
    assert U.count_agreement(out["counters"], ref["counters"]) == 1.0


def test_user_vdp_kernel_source_on_cpu_against_numpy_oracle():
    """A stiff, forced oscillator (time-dependent RHS with a libm call, rejected steps, Jacobian refreshes): the kernel
    source with the CUDA functor against the reference-equivalent numpy oracle with the Python closure."""
    vdp_problem()
    ens = vdp_ensemble(8)
    out = U.run_kernel_source_on_cpu(ens)
    ref = vdp_numpy_oracle(ens)
    assert (out["status"] == 0).all() and (ref["status"] == 0).all()
    assert U.scaled_error(out["y"], ref["y"], 1e-6, 1e-8) <= U.PARITY_FACTOR
    assert U.count_agreement(out["counters"], ref["counters"]) >= 0.75     # relaxation oscillations amplify rounding noise
    tot, tot_ref = out["counters"].sum(axis=0), ref["counters"].sum(axis=0)
    assert np.all(np.abs(tot - tot_ref) <= 0.02 * np.maximum(tot_ref, 50))
    assert out["counters"][:, 5].sum() > 0 and out["counters"][:, 1].min() > 1    # the hard paths were exercised


def reactor_with_events():
    """The reactor functor with two nonlinear event functions compiled in: the reaction rate k Ca Cb, and the squared
    distance of (Ca, Cb) from the origin."""
    return br.compile_problem("reactor_ev", 3, E.REACTOR_RHS, E.REACTOR_JAC, param_names=("k", "Ca0", "Cb0"),
                              param_defaults=(1.47, 0.0209, 0.0209 / 3), mass_diag=(1, 0, 0), var_index=(0, 1, 1),
                              t_span=(0.0, 20.0), events=["p[0] * y[0] * y[1] - 2.0e-4", "y[0] * y[0] + y[1] * y[1] - 4.0e-4"],
                              static_lu=True, first_iter_skip=True)


def test_user_event_functions_are_located_by_the_kernel():
    """FunctorEvent: nonlinear g_k(t, y) compiled into the plugin, located inside the step by the kernel source (run on the
    CPU) with scipy's semantics -- checked against scipy's own event helpers applied, on the host, to the step record of the
    uninterrupted run with the Python restatements of the same functions."""
    from dae_scipy_b200.api import FunctorEvent, find_events
    prob = reactor_with_events()
    assert prob.n_event_functions == 2
    ens = dict(E.reactor_ensemble(24), problem=prob)
    k = ens["params"][:, 0]

    def make(i):        # per-system closures for the host-side checker (the rate event depends on the system's k)
        rate = lambda t, y: k[i] * y[0] * y[1] - 2.0e-4              # noqa: E731
        rate.direction, rate.terminal = -1.0, True
        dist = lambda t, y: y[0] * y[0] + y[1] * y[1] - 4.0e-4       # noqa: E731
        return [rate, dist]
    evs = [FunctorEvent(0, direction=-1.0, terminal=True), FunctorEvent(1)]
    out = U.run_kernel_source_on_cpu(ens, events=evs, max_events=4)
    full = U.run_kernel_source_on_cpu(ens, dense_cap=200)
    n_term = 0
    for i in range(ens["y0"].shape[0]):
        one = full["sol"]
        sub = type(one).__new__(type(one))                # the record of system i alone
        sub.n_steps, sub.truncated, sub.ascending = one.n_steps[i:i + 1], one.truncated[i:i + 1], one.ascending
        sub.ts, sub.ys, sub.Q = one.ts[i:i + 1], one.ys[i:i + 1], one.Q[i:i + 1]
        t_host, y_host, term, t_stop, y_stop = find_events(sub, make(i), ens["t_span"][0], ens["y0"][i:i + 1])
        assert bool(term[0]) == (out["status"][i] == 1)
        for e in range(2):
            kk = out["event_count"][i, e]
            assert kk == t_host[0][e].size
            np.testing.assert_allclose(out["event_t"][i, e, :kk], t_host[0][e], rtol=0, atol=1e-9)
        if term[0]:
            n_term += 1
            np.testing.assert_allclose(out["t_final"][i], t_stop[0], rtol=0, atol=1e-9)
            np.testing.assert_allclose(out["y_final"][i], y_stop[0], rtol=0, atol=1e-12)
            assert abs(k[i] * out["y_final"][i, 0] * out["y_final"][i, 1] - 2.0e-4) < 1e-15
            assert out["counters"][i, 3] <= full["counters"][i, 3]
    assert 0 < n_term < ens["y0"].shape[0]                # some systems stop at the rate event, some never reach it
    assert out["event_count"][:, 1].sum() > 0             # the observer fired too
    with pytest.raises(ValueError):
        U.run_kernel_source_on_cpu(ens, events=[FunctorEvent(2)])          # only two event functions were compiled


# ---------------------------------------------------------------- band problems with more than 8 unknowns (CTA-per-system kernel)
def rod_numpy_oracle(ens):
    """The numpy restatement of the reference on the rod ensemble (finite-difference Jacobian, like the fixture)."""
    from oracle import radau_numpy as orc, problems_numpy as prob
    B, n = ens["y0"].shape
    T = len(ens["t_eval"])
    o = dict(ens["options"])
    o.pop("mass_matrix"); vi = o.pop("var_index"); o.pop("jac")
    out = dict(y=np.full((B, n, T), np.nan), counters=np.zeros((B, 9), np.int32), status=np.zeros(B, np.int32), n_eval=np.zeros(B, np.int32))
    for i in range(B):
        fun, _, mass, _ = prob.build("rod", ens["params"][i], n)
        r = orc.solve(fun, ens["t_span"], ens["y0"][i], jac=None, mass_matrix=mass, var_index=np.asarray(vi, float),
                      t_eval=ens["t_eval"], **o)
        out["y"][i, :, :r.y.shape[1]] = r.y
        out["counters"][i] = r.counters.as_array()
        out["status"][i] = r.status
        out["n_eval"][i] = r.y.shape[1]
    return out


def test_band_problem_header_build_and_abi():
    text = plugin.generate_band_header("rod", 48, 4, 1, 1, E.ROD_RHS_ROW, "i != 0 && i != n - 1")
    assert "struct UserBand" in text and "KL = 1, KU = 1" in text and "N = 48, NP = 4" in text
    with pytest.raises(ValueError):
        plugin.generate_band_header("rod", 48, 0, 20, 12, "return 0.0;")        # kl + ku > 31
    with pytest.raises(ValueError):
        plugin.generate_band_header("rod", 48, 0, 1, 1, "y[i];")                # nothing returned
    prob = E.rod_problem()
    assert prob.problem_id == plugin.BRDAE_USER and prob.n == 48 and prob.band == (1, 1) and not prob.has_jac
    lib = _abi.load(prob.lib_path)
    for sym in _abi.EXPORTED_SYMBOLS:                                           # a complete libbrdae: the same C ABI
        getattr(lib, sym)
    pid, n, npar = C.c_int32(-1), C.c_int32(-1), C.c_int32(-1)
    for name in (b"rod", b"user"):
        assert lib.brdae_problem_lookup(name, C.byref(pid), C.byref(n), C.byref(npar)) == 0
        assert (pid.value, n.value, npar.value) == (100, 48, 4)
    assert lib.brdae_problem_lookup(b"npendulum", C.byref(pid), C.byref(n), C.byref(npar)) == 0       # the stock problems are still there
    ens = E.rod_ensemble(4)
    hb = br.prepare(prob, ens["t_span"], ens["y0"].shape, t_eval=ens["t_eval"], **ens["options"])
    assert hb.n == 48 and hb.options.jac_finite_differences == 1 and hb.mass is None
    assert hb.var_index.tolist() == [1] + [0] * 46 + [1]
    with pytest.raises(ValueError, match="finite-difference"):
        br.prepare(prob, ens["t_span"], ens["y0"].shape, **dict(ens["options"], mass_matrix=np.eye(48)))
    with pytest.raises(ValueError, match="8 unknowns"):                         # a thread-per-system functor stays at n <= 8
        plugin.generate_header("big", 9, 0, "f[0] = 0;")


def test_rod_numpy_oracle_reproduces_the_reference_fixture():
    """The band example's fixture comes from the reference itself (jac=None with the tridiagonal jac_sparsity); the numpy
    restatement with the dense finite-difference Jacobian gives the same run: bands of a finite-difference Jacobian do not
    depend on how the columns are grouped."""
    ens, ref = U.load_golden("user_rod")
    out = rod_numpy_oracle(ens)
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["n_eval"], ref["n_eval"])
    assert U.scaled_error(out["y"], ref["y"], 1e-6, 1e-8) <= 1.0
    assert U.count_agreement(out["counters"], ref["counters"]) == 1.0
