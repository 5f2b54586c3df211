"""Events: scipy's solve_ivp(events=...) semantics on the per-step record (dae_scipy_b200.api.find_events), checked
against `tests/golden/events_pendulum3.npz`, which oracle/gen_golden.py made by running scipy.integrate.solve_ivp with
the reference's RadauDAE and the same event functions."""
import json
import os

import numpy as np
import pytest

from dae_scipy_b200.api import LinearEvent, find_events
from _util import GOLDEN_DIR, run_kernel_source_on_cpu


def pendulum_events():
    def x_zero(t, y):
        return y[0]

    def vy_up(t, y):
        return y[3]
    vy_up.direction = 1.0
    vy_up.terminal = 2
    return [x_zero, vy_up]


def _fixture():
    z = np.load(os.path.join(GOLDEN_DIR, "events_pendulum3.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    opts = dict(meta["options"], var_index=z["var_index"], mass_matrix="problem")
    ens = dict(problem=meta["problem"], t_span=tuple(meta["t_span"]), y0=z["y0"], params=z["params"], t_eval=z["t_eval"],
               options=opts)
    return ens, z


def test_event_location_matches_scipy_with_the_reference_solver():
    ens, ref = _fixture()
    out = run_kernel_source_on_cpu(ens, dense_cap=800)
    t_events, y_events, terminated, t_stop, y_stop = find_events(out["sol"], pendulum_events(), ens["t_span"][0], ens["y0"])
    assert terminated.tolist() == (ref["status"] == 1).tolist()
    for i in range(ens["y0"].shape[0]):
        for e in range(2):
            want = ref["t_events"][i, e]
            want = want[np.isfinite(want)]
            assert t_events[i][e].shape == want.shape
            np.testing.assert_allclose(t_events[i][e], want, rtol=0, atol=2e-6)       # solutions agree to ~10 x tolerance
        np.testing.assert_allclose(t_stop[i], ref["t_final"][i], atol=2e-6)
        np.testing.assert_allclose(y_stop[i][:4], ref["y_final"][i][:4], atol=1e-4)
        assert abs(y_stop[i][3]) < 1e-9                                              # the terminal event function is zero there
        np.testing.assert_allclose(y_events[i][0][:, 0], 0.0, atol=1e-9)             # x = 0 at the x-crossings


def test_non_terminal_events_and_direction_filter():
    ens, _ = _fixture()
    out = run_kernel_source_on_cpu(ens, dense_cap=800)

    def vy_down(t, y):
        return y[3]
    vy_down.direction = -1.0
    t_events, _, terminated, _, _ = find_events(out["sol"], [vy_down], ens["t_span"][0], ens["y0"])
    assert not terminated.any()
    sol = out["sol"]
    for i in range(ens["y0"].shape[0]):
        te = t_events[i][0]
        assert te.size >= 1 and (np.diff(te) > 0).all()
        eps = 1e-4
        vy = sol(np.concatenate([te - eps, te + eps]))[i, 3]
        assert (vy[:te.size] > 0).all() and (vy[te.size:] < 0).all()                  # only downward crossings


# ---------------------------------------------------------------- events located by the kernel itself (LinearEvent)
def linear_pendulum_events():
    """The same two events as `pendulum_events`, stated as linear functionals of the state for the device."""
    return [LinearEvent(0, n=5), LinearEvent(3, n=5, direction=1.0, terminal=2)]


def test_device_side_events_match_scipy_with_the_reference_solver():
    """The event block of the kernel source (run on the CPU): same occurrences as scipy's solve_ivp driving the reference
    solver, the system STOPS at its terminal event (status 1, t_final / y_final at the event, later t_eval points not
    produced), and the roots equal the host-side location on the record of the uninterrupted run to rounding."""
    ens, ref = _fixture()
    evs = linear_pendulum_events()
    out = run_kernel_source_on_cpu(ens, events=evs, max_events=4)
    assert np.array_equal(out["status"], ref["status"]) and (out["status"] == 1).all()
    assert np.array_equal(out["n_eval"], ref["n_eval"])
    np.testing.assert_allclose(out["t_final"], ref["t_final"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["y_final"][:, :4], ref["y_final"][:, :4], atol=1e-4)
    assert np.abs(out["y_final"][:, 3]).max() < 1e-9                           # the terminal event function is zero there
    full = run_kernel_source_on_cpu(ens, dense_cap=800)                          # the same integration, not interrupted
    t_host, y_host, term, t_stop, y_stop = find_events(full["sol"], evs, ens["t_span"][0], ens["y0"])
    assert term.all()
    for i in range(ens["y0"].shape[0]):
        for e in range(2):
            want = ref["t_events"][i, e]
            want = want[np.isfinite(want)]
            k = out["event_count"][i, e]
            assert k == want.size == t_host[i][e].size
            np.testing.assert_allclose(out["event_t"][i, e, :k], want, rtol=0, atol=2e-6)
            np.testing.assert_allclose(out["event_t"][i, e, :k], t_host[i][e], rtol=0, atol=1e-12)
            np.testing.assert_allclose(out["event_y"][i, e, :k], y_host[i][e], rtol=0, atol=1e-9)
            assert np.isnan(out["event_t"][i, e, k:]).all()
        np.testing.assert_allclose(out["event_y"][i, 0, :out["event_count"][i, 0], 0], 0.0, atol=1e-12)   # x = 0 at its crossings
    np.testing.assert_allclose(out["t_final"], t_stop, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["y_final"], y_stop, rtol=0, atol=1e-9)
    # counters: the interrupted run did not take the steps after the event
    assert (out["counters"][:, 3] < full["counters"][:, 3]).all()


def test_device_side_events_direction_level_cap_and_no_events():
    ens, _ = _fixture()
    plain = run_kernel_source_on_cpu(ens)
    # an event that never fires changes nothing
    never = run_kernel_source_on_cpu(ens, events=[LinearEvent(0, level=50.0, n=5, terminal=True)])
    for k in ("y", "status", "n_eval", "counters", "y_final", "t_final"):
        assert np.array_equal(never[k], plain[k], equal_nan=True)
    assert (never["event_count"] == 0).all()
    # non-terminal, downward only, level crossing, only one occurrence recorded of several counted
    evs = [LinearEvent(3, n=5, direction=-1.0), LinearEvent([0.0, 1.0, 0.0, 0.0, 0.0], level=-0.5)]
    out = run_kernel_source_on_cpu(ens, events=evs, max_events=1)
    for k in ("y", "status", "n_eval", "counters", "y_final", "t_final"):       # observers: the integration is untouched
        assert np.array_equal(out[k], plain[k], equal_nan=True)
    full = run_kernel_source_on_cpu(ens, dense_cap=800)
    t_host, y_host, term, _, _ = find_events(full["sol"], evs, ens["t_span"][0], ens["y0"])
    assert not term.any()
    for i in range(ens["y0"].shape[0]):
        for e in range(2):
            assert out["event_count"][i, e] == t_host[i][e].size
            if t_host[i][e].size:
                np.testing.assert_allclose(out["event_t"][i, e, 0], t_host[i][e][0], rtol=0, atol=1e-12)
        if t_host[i][1].size:
            np.testing.assert_allclose(out["event_y"][i, 1, 0, 1], -0.5, atol=1e-12)     # y = level at the crossing
    assert (out["event_count"][:, 0] >= 1).all()


def test_linear_event_packing_and_validation():
    from dae_scipy_b200.api import pack_linear_events
    coef, level, direction, terminal, functor = pack_linear_events(linear_pendulum_events(), 5)
    assert coef.shape == (2, 5) and coef[0, 0] == 1 and coef[1, 3] == 1 and level.tolist() == [0.0, 0.0]
    assert direction.tolist() == [0, 1] and terminal.tolist() == [0, 2] and functor.tolist() == [-1, -1]
    assert pack_linear_events([LinearEvent(1, n=3, terminal=True)], 3)[3].tolist() == [1]
    with pytest.raises(ValueError):
        pack_linear_events([LinearEvent(0, n=5)] * 5, 5)            # at most BRDAE_MAX_EVENTS
    with pytest.raises(ValueError):
        pack_linear_events([LinearEvent(0, n=4)], 5)
    with pytest.raises(ValueError):
        LinearEvent(2)                                              # an index needs n
    ev = LinearEvent([1.0, -1.0], level=0.25)
    assert ev(0.0, np.array([2.0, 0.5])) == 1.25                     # an ordinary event callable as well


def test_device_side_events_random_combinations_against_scipy_helpers():
    """Seeded differential run: random linear events (single components and dense functionals, random levels, directions,
    terminal after the 1st / 2nd / 3rd occurrence or never, forward and backward integration, several events at once) located
    by the kernel source, against scipy's find_active_events / handle_events / brentq applied on the host to the step record
    of the uninterrupted run: same occurrences, same terminations, times to 1e-9."""
    import dae_scipy_b200 as br
    E = br.ensembles
    rng = np.random.default_rng(20261021)
    scale = np.array([1, 1, 3, 3, 10.0])
    n_fired = n_term = 0
    for case in range(14):
        ens = E.pendulum_ensemble(3, t_end=float(rng.choice([1.0, 3.0])), seed=int(rng.integers(1 << 30)), n_t_eval=int(rng.integers(2, 40)))
        if rng.random() < 0.3:
            ens["t_span"] = (ens["t_span"][1], ens["t_span"][0])
            ens["t_eval"] = ens["t_eval"][::-1].copy()
        evs = []
        for _ in range(int(rng.integers(1, 5))):
            coef = np.zeros(5)
            if rng.random() < 0.5:
                coef[int(rng.integers(5))] = 1.0
            else:
                coef = rng.standard_normal(5) / scale
            level = float(coef @ ens["y0"].mean(axis=0) + rng.standard_normal() * 0.3 * np.abs(coef * scale).sum())
            evs.append(LinearEvent(coef, level=level, direction=float(rng.choice([-1, 0, 1])),
                                   terminal=[False, False, True, 2, 3][int(rng.integers(5))]))
        out = run_kernel_source_on_cpu(ens, events=evs, max_events=64)
        full = run_kernel_source_on_cpu(ens, dense_cap=3000)
        t_host, _, term, t_stop, _ = find_events(full["sol"], evs, ens["t_span"][0], ens["y0"])
        assert term.tolist() == (out["status"] == 1).tolist(), case
        for i in range(3):
            for e in range(len(evs)):
                k = out["event_count"][i, e]
                assert k == t_host[i][e].size, (case, i, e)
                np.testing.assert_allclose(out["event_t"][i, e, :k], t_host[i][e], rtol=0, atol=3e-9)
                n_fired += int(k)
            if term[i]:
                n_term += 1
                np.testing.assert_allclose(out["t_final"][i], t_stop[i], rtol=0, atol=3e-9)
    assert n_fired > 50 and n_term > 5            # the run did exercise occurrences and terminations
