"""Events: scipy's solve_ivp(events=...) semantics on the per-step record (dae_scipy_b200.api.find_events), checked
against `tests/golden/events_pendulum3.npz`, which oracle/gen_golden.py made by running scipy.integrate.solve_ivp with
the reference's RadauDAE and the same event functions."""
import json
import os

import numpy as np
import pytest

import dae_scipy_b200 as br
from dae_scipy_b200.api import find_events
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
