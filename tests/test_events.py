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
