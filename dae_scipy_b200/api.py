"""solve_ivp_batched — the batched drop-in for `solve_ivp(fun, t_span, y0, method=RadauDAE, ...)`.

Mirror of the reference's user-facing call (reference README.md:19-21; option surface
scipyDAE/radauDAE.py:263-278; driver radauDAE.py:990-1195 / scipy ivp.py:560-760) for an
ensemble of B independent systems.  What changes with respect to the reference call:

  * `fun`/`jac` closures become a problem name (a CUDA device functor, see problems.py);
    per-system parameters are a [B, p] array;
  * `y0` is [B, n]; results carry a leading system axis: `y` is [B, n, T];
  * `status`, `success`, `nfev`, `njev`, `nlu` and the RadauDAE counters are per-system arrays.

Everything numerical happens in libbrdae.so (hand-written CUDA for sm_100a) through the C ABI
of include/brdae.h; torch tensors are only the device buffers.  Without the built library or
without a CUDA device the call raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import warnings
from dataclasses import dataclass

import numpy as np

from . import _abi
from .ode_solver import DeviceFun, RadauDAE          # RadauDAE: the `method=` marker, and a scipy OdeSolver for one system
from .problems import Problem, get_problem


class LinearEvent:
    """An event function that is linear in the state, g(t, y) = coef . y - level, with scipy's `terminal` / `direction`
    attributes (solve_ivp(events=...), scipy ivp.py:25-60): such events are located ON THE DEVICE, inside the step that
    crosses them, and a terminal one stops the integration of that system there.  `coef` is an [n] vector or the index
    of one unknown; `terminal` may be True or the occurrence count that terminates (scipy >= 1.13).  It is also an
    ordinary callable, so it works wherever a Python event function does."""

    def __init__(self, coef, level=0.0, direction=0.0, terminal=False, n=None):
        if np.ndim(coef) == 0:
            if n is None:
                raise ValueError("LinearEvent(index, ...) needs n, the number of unknowns.")
            c = np.zeros(int(n))
            c[int(coef)] = 1.0
            coef = c
        self.coef = np.ascontiguousarray(coef, dtype=np.float64)
        self.level = float(level)
        self.direction = float(direction)
        self.terminal = terminal

    def __call__(self, t, y):
        return float(np.dot(self.coef, np.asarray(y)[:self.coef.size]) - self.level)


class FunctorEvent:
    """The k-th event function compiled into a user problem (`compile_problem(..., events=[...])`): any, also nonlinear,
    g_k(t, y), located on the device like a LinearEvent, with the same `direction` / `terminal` semantics.  `fun` is an
    optional Python restatement `fun(t, y)`; with it the object is an ordinary event callable too (host-side location)."""

    def __init__(self, k, direction=0.0, terminal=False, fun=None):
        self.k = int(k)
        self.direction = float(direction)
        self.terminal = terminal
        self.fun = fun

    def __call__(self, t, y):
        if self.fun is None:
            raise TypeError("this FunctorEvent exists on the device only (no Python restatement `fun` was given).")
        return self.fun(t, y)


def pack_linear_events(events, n, n_event_functions=0):
    """[LinearEvent | FunctorEvent] -> the arrays of brdae_batch (coef [E, n], level [E], direction [E], terminal count [E],
    functor index [E] with -1 for the linear form)."""
    if len(events) > _abi.MAX_EVENTS:
        raise ValueError(f"at most {_abi.MAX_EVENTS} events can be located on the device.")
    for ev in events:
        if isinstance(ev, FunctorEvent):
            if not 0 <= ev.k < n_event_functions:
                raise ValueError(f"FunctorEvent({ev.k}): the problem was compiled with {n_event_functions} event function(s).")
        elif ev.coef.shape != (n,):
            raise ValueError(f"LinearEvent coefficients must have shape {(n,)}.")
    coef = np.ascontiguousarray(np.stack([np.zeros(n) if isinstance(ev, FunctorEvent) else ev.coef for ev in events]))
    level = np.array([0.0 if isinstance(ev, FunctorEvent) else ev.level for ev in events], dtype=np.float64)
    functor = np.array([ev.k if isinstance(ev, FunctorEvent) else -1 for ev in events], dtype=np.int32)
    direction = np.array([int(np.sign(ev.direction)) for ev in events], dtype=np.int32)
    terminal = np.array([int(ev.terminal) if ev.terminal is not None else 0 for ev in events], dtype=np.int32)   # True -> 1
    if np.any(terminal < 0):
        raise ValueError("`terminal` must be a bool or a positive occurrence count.")
    return coef, level, direction, terminal, functor


def warn_extraneous(extraneous):
    """scipy common.py:27-43 — same warning text."""
    if extraneous:
        warnings.warn("The following arguments have no effect for a chosen solver: "
                      f"{', '.join(f'`{x}`' for x in extraneous)}.", stacklevel=3)


@dataclass
class HostBatch:
    """Validated, host-side description of one call (small uniform arrays + options)."""
    problem: Problem
    n: int
    B: int
    t0: float
    t_bound: float
    t_eval: np.ndarray          # float64 [T], in user order (monotone in the direction of integration)
    rtol: float
    atol: np.ndarray            # float64 [n]
    var_index: np.ndarray       # int32 [n]
    mass: np.ndarray | None     # float64 [n, n] shared matrix, None = the problem's own (device functor)
    jac_const: np.ndarray | None
    options: _abi.BrdaeOptions
    events: tuple | None = None    # device-side events: (coef [E, n], level [E], direction, terminal, functor index [E] int32)

    def struct(self, ptr) -> _abi.BrdaeBatch:
        """Fill a brdae_batch; `ptr` maps field name -> address of the batch-sized buffers."""
        b = _abi.BrdaeBatch()
        b.problem = self.problem.problem_id
        b.n = self.n
        b.n_params = self.problem.n_params
        b.n_t_eval = self.t_eval.size
        b.n_systems = self.B
        b.rtol = self.rtol
        b.t0 = self.t0
        b.t_bound = self.t_bound
        b.mass = self.mass.ctypes.data if self.mass is not None else None
        b.jac_const = self.jac_const.ctypes.data if self.jac_const is not None else None
        b.var_index = self.var_index.ctypes.data
        b.atol = self.atol.ctypes.data
        for k in ("y0", "params", "t_eval", "y_eval", "n_eval", "t_final", "y_final", "status", "counters",
                  "dense_t", "dense_y", "dense_q"):
            setattr(b, k, ptr.get(k))
        b.dense_cap = int(ptr.get("dense_cap") or 0)
        b.report_cap = int(ptr.get("report_cap") or 0)
        if b.report_cap:
            for k in ("report_count", "report_code", "report_it", "report_val"):
                setattr(b, k, ptr.get(k))
        if self.events is not None and ptr.get("event_cap"):
            coef, level, direction, terminal, functor = self.events
            b.n_events, b.event_cap = coef.shape[0], int(ptr["event_cap"])
            b.event_coef, b.event_level = coef.ctypes.data, level.ctypes.data
            b.event_direction, b.event_terminal = direction.ctypes.data, terminal.ctypes.data
            b.event_functor = functor.ctypes.data
            for k in ("event_count", "event_t", "event_y"):
                setattr(b, k, ptr.get(k))
        return b


def prepare(problem, t_span, y0_shape, *, t_eval=None, rtol=1e-3, atol=1e-6, mass_matrix="problem",
            var_index=None, jac="analytic", params=None, **options) -> HostBatch:
    """Argument validation of solve_ivp + RadauDAE.__init__, for a batch.

    Raises the same ValueErrors as the reference path: t_eval checks (scipy ivp.py:603-621 /
    radauDAE.py:1018-1028), tolerance checks (scipy common.py:45-60), mass-matrix and Jacobian
    shape checks (radauDAE.py:447-462, 499-510), first_step/max_step (common.py:10-24).
    """
    prob = get_problem(problem)
    if len(y0_shape) != 2:
        raise ValueError("`y0` must be 2-dimensional [B, n] for the batched solver.")
    B, n = int(y0_shape[0]), int(y0_shape[1])
    if prob.n and n != prob.n:
        raise ValueError(f"problem {prob.name!r} has {prob.n} unknowns, `y0` has {n}.")
    if not prob.n and (n <= 0 or n % 5):
        raise ValueError("`y0` of the n-pendulum must have 5*nodes columns.")
    t0, tf = float(t_span[0]), float(t_span[1])

    if t_eval is None:
        te = np.empty(0)
    else:
        te = np.asarray(t_eval, dtype=np.float64)
        if te.ndim != 1:
            raise ValueError("`t_eval` must be 1-dimensional.")
        if np.any(te < min(t0, tf)) or np.any(te > max(t0, tf)):
            raise ValueError("Values in `t_eval` are not within `t_span`.")
        d = np.diff(te)
        if tf > t0 and np.any(d <= 0) or tf < t0 and np.any(d >= 0):
            raise ValueError("Values in `t_eval` are not properly sorted.")
        te = np.ascontiguousarray(te)

    rtol = float(rtol)
    if rtol < 100 * np.finfo(float).eps:
        warnings.warn("At least one element of `rtol` is too small. "
                      f"Setting `rtol = np.maximum(rtol, {100 * np.finfo(float).eps})`.", stacklevel=3)
        rtol = 100 * np.finfo(float).eps
    at = np.asarray(atol, dtype=np.float64)
    if at.ndim > 0 and at.shape != (n,):
        raise ValueError("`atol` has wrong shape.")
    if np.any(at < 0):
        raise ValueError("`atol` must be positive.")
    at = np.ascontiguousarray(np.broadcast_to(at, (n,)))

    if var_index is None:
        vi = np.zeros(n, dtype=np.int32)                   # radauDAE.py:389-390: all differential
    elif isinstance(var_index, str):
        if var_index != "problem":
            raise ValueError("`var_index` must be None, an [n] array or the string 'problem'.")
        vi = np.ascontiguousarray(prob.var_index_array(n), dtype=np.int32)      # the reference script's own choice
    else:
        v = np.asarray(var_index)
        if v.ndim != 1 or v.size != n:
            raise ValueError("`var_index` must be a 1-D array with one entry per unknown.")
        if np.any(v != np.round(v)) or np.any(v < 0) or np.any(v > 3):
            raise ValueError("`var_index` entries must be integers between 0 and 3.")
        vi = np.ascontiguousarray(v, dtype=np.int32)

    mass = None
    if isinstance(mass_matrix, str):
        if mass_matrix != "problem":
            raise ValueError("`mass_matrix` must be None, an [n, n] array or the string 'problem'.")
    elif mass_matrix is None:
        mass = np.eye(n)                                   # radauDAE.py:448-449
        if options.get("first_step") is None:
            # Without a mass matrix the reference takes its first step from scipy's select_initial_step (radauDAE.py:307-311),
            # which is not restated on the device (and does not run under scipy 1.18: SURVEY 4.3).  Starting silently from
            # the with-mass default 1e-6 would change the step sequence, so the combination is refused.
            raise ValueError("`first_step` is required when `mass_matrix` is None (the reference would call "
                             "scipy's select_initial_step here); pass first_step=... or the mass matrix.")
    elif callable(mass_matrix):
        raise ValueError("`mass_matrix` should be a constant matrix, but is callable")
    else:
        if hasattr(mass_matrix, "toarray"):                # scipy.sparse matrix (radauDAE.py:452-457 keeps it as CSC; the
            mass_matrix = mass_matrix.toarray()            # reference's own chain script passes one): the kernels take it dense
        mass = np.ascontiguousarray(np.asarray(mass_matrix, dtype=np.float64))
        if mass.shape != (n, n):
            raise ValueError("`mass_matrix` is expected to have shape {}, but actually has {}."
                             .format((n, n), mass.shape))
        # the problem's own constant matrix -> the kernel with the compile-time mass structure
        if not prob.per_system_mass and np.array_equal(mass, prob.mass(n=n)):
            mass = None

    jac_const = None
    fd = False
    if isinstance(jac, str):
        if jac != "analytic":
            raise ValueError("`jac` must be 'analytic' or a constant [n, n] matrix.")
        if not prob.has_jac:
            fd = True          # a user problem compiled without a Jacobian body: the reference's default, num_jac
    elif jac is None:
        # the reference's default (radauDAE.py:468-481): scipy num_jac forward differences of the RHS
        # functor with adaptive column factors, on the device (thread-per-system problems)
        # (the n-pendulum chain: the reference's own default for that problem, recursive_pendulum.py:128, 165 -- runs on the
        # band kernel with one right-hand-side evaluation per group of 17 columns; jac='analytic' selects the much faster
        # block-elimination kernel)
        fd = True
    elif callable(jac):
        raise ValueError("Python Jacobian callables cannot run on the device; use jac='analytic'.")
    else:
        jac_const = np.ascontiguousarray(np.asarray(jac, dtype=np.float64))
        if jac_const.shape != (n, n):
            raise ValueError(f"`jac` is expected to have shape {(n, n)}, but actually has {jac_const.shape}.")

    first_step = options.get("first_step")
    if first_step is not None:
        if first_step <= 0:
            raise ValueError("`first_step` must be positive.")
        if first_step > abs(tf - t0):
            raise ValueError("`first_step` exceeds bounds.")
    ignored = {k: options.pop(k) for k in list(options) if k in _abi.IGNORED_OPTS}
    del ignored
    extraneous = {k: options.pop(k) for k in list(options) if k not in _abi.SOLVER_OPTION_NAMES}
    warn_extraneous(extraneous)
    opts = _abi.make_options(jac_finite_differences=fd, **options)
    if n > 8 and prob.name != "npendulum" and prob.band is None:
        raise ValueError("thread-per-system kernels support at most 8 unknowns (larger banded systems: compile_band_problem).")
    if prob.band is not None and (mass is not None or jac_const is not None):
        raise ValueError("a band problem runs with its own diag(0/1) mass matrix and a finite-difference Jacobian.")
    return HostBatch(problem=prob, n=n, B=B, t0=t0, t_bound=tf, t_eval=te, rtol=rtol, atol=at,
                     var_index=vi, mass=mass, jac_const=jac_const, options=opts)


class BatchedOdeResult:
    """OdeResult with a leading system axis (scipy ivp.py `OdeResult`; radauDAE.py:1183-1194).

    Tensors stay on the device they were computed on; `.numpy()` copies everything to the host.
      t [T]; y [B, n, T] (NaN after a failure point); n_eval [B]; status [B]; success [B];
      t_final [B]; y_final [B, n]; nfev, njev, nlu, nstep, naccpt, nrejct, nfailed, nlusolve,
      nlusolve_errorest [B].
    """

    def __init__(self, t, y_eval_tnb, n_eval, status, t_final, y_final_nb, counters_cb):
        self.sol = None                     # BatchedOdeSolution when dense_output=True
        self.t_events = self.y_events = None    # per system, per event (lists of arrays) when events were given
        self._event_cut = None                  # (terminated [B], t_stop [B], y_stop [B, n], ascending)
        self.event_count = self.event_t = self.event_y = None   # device-side events: [E, B], [E, cap, B], [E, cap, n, B]
        self.reports = None                     # report=True: dict of device tensors, see solve_ivp_batched
        self.tsub = self.ysub = None            # return_substeps=True: stage points of every step (host arrays)
        self.t = t
        self._y_eval = y_eval_tnb           # [T, n, B] as written by the kernel (coalesced SoA)
        self.n_eval = n_eval
        self.status = status
        self.t_final = t_final
        self._y_final = y_final_nb          # [n, B]
        self.counters = counters_cb         # [9, B]

    @property
    def y(self):
        return self._y_eval.permute(2, 1, 0)    # [B, n, T] view, no copy

    @property
    def y_final(self):
        return self._y_final.t()

    @property
    def success(self):
        return self.status >= 0

    def __getattr__(self, name):
        if name in _abi.COUNTER_NAMES:
            return self.counters[_abi.COUNTER_NAMES.index(name)]
        raise AttributeError(name)

    def message(self, i):
        return _abi.MESSAGES.get(int(self.status[i]), "unknown status")

    def numpy(self):
        out = {"t": np.asarray(self.t), "y": self.y.cpu().numpy(), "n_eval": self.n_eval.cpu().numpy(),
               "status": self.status.cpu().numpy(), "t_final": self.t_final.cpu().numpy(),
               "y_final": self.y_final.cpu().numpy(), "counters": self.counters.t().cpu().numpy()}
        for i, name in enumerate(_abi.COUNTER_NAMES):
            out[name] = out["counters"][:, i]
        if self.event_count is not None:        # events located on the device: arrays, and scipy's per-system lists
            cnt, et, ey = self.event_count.cpu().numpy(), self.event_t.cpu().numpy(), self.event_y.cpu().numpy()
            out["event_count"], out["event_t"], out["event_y"] = cnt.T, et.transpose(2, 0, 1), ey.transpose(3, 0, 1, 2)
            cap = et.shape[1]
            k = np.minimum(cnt, cap)
            out["t_events"] = [[out["event_t"][i, e, :k[e, i]] for e in range(cnt.shape[0])] for i in range(cnt.shape[1])]
            out["y_events"] = [[out["event_y"][i, e, :k[e, i]] for e in range(cnt.shape[0])] for i in range(cnt.shape[1])]
        if self.reports is not None:            # the reference's solver.reports (bReport=True), one row per system
            r = self.reports
            out["reports"] = {"n": r["count"].cpu().numpy(), "code": r["code"].t().cpu().numpy(),
                              "newton_iterations": r["it"][:, 0].t().cpu().numpy(), "bad_iterations": r["it"][:, 1].t().cpu().numpy(),
                              "t": r["val"][:, 0].t().cpu().numpy(), "dt": r["val"][:, 1].t().cpu().numpy(),
                              "err1": r["val"][:, 2].t().cpu().numpy(), "err2": r["val"][:, 3].t().cpu().numpy()}
        if self._event_cut is not None:         # a terminal event ends the reported result (scipy: status 1)
            term, t_stop, y_stop, asc = self._event_cut
            t = out["t"]
            for i in np.nonzero(term)[0]:
                keep = (t <= t_stop[i]) if asc else (t >= t_stop[i])
                keep &= np.arange(t.size) < out["n_eval"][i]
                out["y"][i][:, ~keep] = np.nan
                out["n_eval"][i] = int(keep.sum())
                out["status"][i] = 1
                out["t_final"][i] = t_stop[i]
                out["y_final"][i] = y_stop[i]
            out["t_events"], out["y_events"] = self.t_events, self.y_events
        out["success"] = out["status"] >= 0
        return out

    def system(self, i):
        """System i as the `OdeResult` the reference's `solve_ivp(fun_i, ..., method=RadauDAE)` returns (scipy ivp.py
        `OdeResult`; radauDAE.py:1183-1194): `t` (the `t_eval` points reached; with `t_eval=None` and a step record, the
        steps taken), `y` [n, len(t)], `sol`, `t_events`, `y_events`, `nfev`, `njev`, `nlu`, `status`, `message`, `success`
        -- plus the reference solver's own counters.  Host copies of the whole result are made once and cached."""
        from scipy.integrate._ivp.ivp import OdeResult
        if getattr(self, "_host", None) is None:
            self._host = self.numpy()
        h, i = self._host, int(i)
        k = int(h["n_eval"][i])
        t, y = h["t"][:k], h["y"][i][:, :k]
        if h["t"].size == 0 and self.sol is not None:            # t_eval=None: the steps themselves (ivp.py:699-701)
            ns = int(self.sol.n_steps[i])
            t, y = self.sol.ts[i, :ns + 1], self.sol.ys[i, :, :ns + 1]
        status = int(h["status"][i])
        extra = {name: int(h[name][i]) for name in _abi.COUNTER_NAMES if name not in ("nfev", "njev", "nlu")}
        return OdeResult(t=t, y=y, sol=None if self.sol is None else self.sol.system(i),
                         t_events=None if h.get("t_events") is None else h["t_events"][i],
                         y_events=None if h.get("y_events") is None else h["y_events"][i],
                         nfev=int(h["nfev"][i]), njev=int(h["njev"][i]), nlu=int(h["nlu"][i]), status=status,
                         message=_abi.MESSAGES.get(status, "unknown status"), success=status >= 0,
                         t_final=float(h["t_final"][i]), y_final=h["y_final"][i], **extra)


class BatchedOdeSolution:
    """What `solve_ivp(dense_output=True)` returns as `sol` (scipy OdeSolution over the reference's
    RadauDenseOutput cubics, radauDAE.py:951-982), for every system of the ensemble.

    The kernels record, per system, the end time, end state and interpolation coefficients Q of each of the
    first `max_dense_steps` accepted steps; `ts`/`ys` are also what the reference returns as `OdeResult.t`/`.y`
    when `t_eval` is None.  Arrays are host copies in the row-per-system layout:
      n_steps [B]; ts [B, K+1] (ts[:, 0] = t0, NaN beyond n_steps); ys [B, n, K+1]; Q [B, K, n, 3];
      truncated [B] (the system took more than K steps: `sol` is only valid up to ts[i, K]).
    `sol(t)` -> [B, n, len(t)] evaluates all systems at the times `t` (OdeSolution semantics: segment by
    searchsorted, times outside the recorded range are extrapolated from the first/last step); `sol.system(i)`
    returns the callable of one system."""

    def __init__(self, t0, y0, dense_t, dense_y, dense_q, n_steps, truncated, ascending):
        B, K = dense_t.shape
        self.n_steps, self.truncated, self.ascending = n_steps, truncated, ascending
        self.ts = np.concatenate([np.full((B, 1), t0), dense_t], axis=1)
        self.ys = np.concatenate([y0[:, :, None], dense_y], axis=2)
        self.Q = dense_q
        beyond = np.arange(K + 1)[None, :] > n_steps[:, None]
        self.ts[beyond] = np.nan
        self.ys[np.broadcast_to(beyond[:, None, :], self.ys.shape)] = np.nan

    def __call__(self, t):
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        B, K1 = self.ts.shape
        valid = np.arange(K1)[None, :] <= self.n_steps[:, None]
        sgn = 1.0 if self.ascending else -1.0
        # searchsorted(ts, t, side='left') per system on the recorded boundaries (ascending after the sign flip)
        below = (sgn * self.ts[:, :, None] < sgn * t[None, None, :]) & valid[:, :, None]
        seg = np.clip(below.sum(axis=1) - 1, 0, np.maximum(self.n_steps, 1)[:, None] - 1)        # [B, m]
        rows = np.arange(B)[:, None]
        t_old, t_new = self.ts[rows, seg], self.ts[rows, seg + 1]
        x = (t[None, :] - t_old) / (t_new - t_old)
        Q = self.Q[rows, seg]                                                                   # [B, m, n, 3]
        y = Q[..., 0] * x[..., None] + Q[..., 1] * (x * x)[..., None] + Q[..., 2] * (x * x * x)[..., None]
        y = y + np.moveaxis(self.ys, 2, 1)[rows, seg]
        return np.moveaxis(y, 1, 2)                                                             # [B, n, m]

    def system(self, i):
        return lambda t: self(np.atleast_1d(t))[i] if np.ndim(t) else self([t])[i, :, 0]

    def substeps(self):
        """The collocation (stage) points of every recorded step -- what the reference returns with `return_substeps=True`
        (radauDAE.py:661-663, 1097-1134, 1175-1191: `tsub`, `ysub`): `tsub` [B, 3 K] = t_old + h C and `ysub` [B, n, 3 K] =
        y_old + Z_i for the three Radau IIA stages of each step, NaN beyond a system's last step.  The stage values are
        recovered from the step's interpolation coefficients, Z^T = Q P^-1, i.e. the step's cubic evaluated at x = C_i (the
        third stage is the end of the step exactly)."""
        C = np.array([(4 - 6 ** 0.5) / 10, (4 + 6 ** 0.5) / 10, 1.0])          # radauDAE.py:14
        B, K1 = self.ts.shape
        K = K1 - 1
        h = self.ts[:, 1:] - self.ts[:, :-1]                                       # [B, K]
        tsub = self.ts[:, :-1, None] + h[:, :, None] * C[None, None, :]            # [B, K, 3]
        X = np.stack([C, C * C, C * C * C])                                        # [3 powers, 3 stages]
        ysub = np.einsum("bknp,ps->bkns", self.Q, X) + np.moveaxis(self.ys, 2, 1)[:, :-1, :, None]     # [B, K, n, 3]
        ysub[:, :, :, 2] = np.moveaxis(self.ys, 2, 1)[:, 1:, :]                    # C_3 = 1: the recorded end state itself
        beyond = np.arange(K)[None, :] >= self.n_steps[:, None]
        tsub[beyond] = np.nan
        ysub[beyond] = np.nan
        n = ysub.shape[2]
        return tsub.reshape(B, 3 * K), np.ascontiguousarray(np.moveaxis(ysub, 2, 1)).reshape(B, n, 3 * K)

    def step(self, i, k):
        """The cubic of step k of system i as a callable t -> y[n] (RadauDenseOutput, radauDAE.py:959-982)."""
        t_old, h = self.ts[i, k], self.ts[i, k + 1] - self.ts[i, k]
        Q, y_old = self.Q[i, k], self.ys[i, :, k]

        def cubic(t):
            x = (np.asarray(t, dtype=np.float64) - t_old) / h
            p = np.stack([x, x * x, x * x * x])
            return (Q @ p) + (y_old if p.ndim == 1 else y_old[:, None])
        return cubic


def find_events(sol, events, t0, y0):
    """Event location on the recorded solution, with the semantics of `scipy.integrate.solve_ivp(events=...)`
    (scipy ivp.py: `prepare_events`, `find_active_events`, `handle_events` -- scipy's own helpers are used; the
    reference's `solve_ivp_custom` wraps the same ones, radauDAE.py:1059-1129).  `events` are Python callables
    `event(t, y)` with the optional attributes `terminal` and `direction`; they are evaluated on the host, step by
    step, on each system's record, and roots are bracketed inside a step and refined on its cubic interpolant.
    Returns (t_events, y_events, terminated, t_stop, y_stop): `t_events[i][e]` is the array of times at which
    event e fired for system i; a terminal event cuts that system's record at `t_stop[i]`.  The integration on
    the device is not interrupted: it is the reported result that ends at the event."""
    from scipy.integrate._ivp.ivp import find_active_events, handle_events, prepare_events
    events, max_events, direction = prepare_events(events)
    B, n = sol.ys.shape[0], sol.ys.shape[1]
    t_events = [[[] for _ in events] for _ in range(B)]
    y_events = [[[] for _ in events] for _ in range(B)]
    terminated = np.zeros(B, dtype=bool)
    t_stop = np.full(B, np.nan)
    y_stop = np.full((B, n), np.nan)
    for i in range(B):
        g = [ev(t0, y0[i]) for ev in events]
        count = np.zeros(len(events))
        for k in range(int(sol.n_steps[i])):
            t_old, t = sol.ts[i, k], sol.ts[i, k + 1]
            g_new = [ev(t, sol.ys[i, :, k + 1]) for ev in events]
            active = find_active_events(g, g_new, direction)
            if active.size > 0:
                cubic = sol.step(i, k)
                count[active] += 1
                idx, roots, terminate = handle_events(cubic, events, active, count, max_events, t_old, t)
                for e, te in zip(idx, roots):
                    t_events[i][e].append(te)
                    y_events[i][e].append(cubic(te))
                if terminate:
                    terminated[i] = True
                    t_stop[i] = roots[-1]
                    y_stop[i] = cubic(roots[-1])
                    break
            g = g_new
    t_events = [[np.asarray(te) for te in sys_t] for sys_t in t_events]
    y_events = [[np.asarray(ye).reshape(-1, n) for ye in sys_y] for sys_y in y_events]
    return t_events, y_events, terminated, t_stop, y_stop


def _as_device_f64(x, device, torch):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def solve_ivp_batched(problem, t_span, y0, params=None, *, method=RadauDAE, t_eval=None,
                      dense_output=False, max_dense_steps=None, events=None, max_events=8, report=False, max_reports=4096,
                      return_substeps=False, device=None, stream=None, **options):
    """Integrate B independent systems  M y' = f(t, y)  from t_span[0] to t_span[1] on one GPU.

    Parameters follow `scipy.integrate.solve_ivp(..., method=RadauDAE, **options)` of the
    reference (all RadauDAE keywords keep their names and defaults: rtol, atol, first_step,
    max_step, mass_matrix, var_index, newton_tol, max_newton_ite, max_bad_ite, min_factor,
    max_factor, safety_factor, factor_on_non_convergence, scale_residuals, scale_newton_norm,
    scale_error, zero_algebraic_error, jacobianRecomputeFactor, bAlwaysApply2ndEstimate,
    bUsePredictiveController, bUsePredictiveNewtonStoppingCriterion, bUseExtrapolatedGuess,
    bPerformAllNewtonIterations, constant_dt, nmax_step), with these batch-specific ones:

    problem : str | Problem — device functor providing f and its analytic Jacobian.
    y0      : [B, n] array or tensor.       params : [B, p] array/tensor, None = defaults.
    mass_matrix : 'problem' (default: the problem's own matrix, per-system for the transistor),
                  an [n, n] array shared by the ensemble, or None (identity, as in the reference).
    var_index   : [n] integers 0..3 or None (all differential), as in the reference; 'problem' = the one the reference's
                  script for this problem passes.
    jac         : 'analytic' (default: the problem's analytic Jacobian functor), a constant [n, n] matrix, or
                  None = finite differences as in the reference's default (scipy num_jac).
    events      : `LinearEvent`s (g = coef . y - level) and, for user problems compiled with event functions, `FunctorEvent`s
                  (any g_k(t, y)) are located ON THE DEVICE inside the step that crosses them: no step
                  record is needed, `result.event_count` [E, B], `result.event_t` [E, max_events, B] and `result.event_y`
                  [E, max_events, n, B] stay on the device (`.numpy()` also builds scipy's `t_events` / `y_events` lists),
                  and a terminal event STOPS the integration of that system (status 1, `t_final` / `y_final` = the event
                  point).  Thread-per-system problems; the n-pendulum chain and arbitrary Python callables use the host path:
                  callable or list of callables `event(t, y)` with optional `terminal` / `direction` attributes, as in
                  scipy's solve_ivp; located on the host on the per-step record (implies dense_output=True, so
                  `max_dense_steps` is required); `result.t_events[i][e]`, `result.y_events[i][e]`, and a terminal
                  event cuts that system's reported result (status 1), see `find_events`.
    report, max_reports : the reference's `bReport=True` (`solver.reports`, radauDAE.py:424-443): one record per factorisation,
                  Jacobian update, non-converged attempt, rejected and accepted step (codes -10, -11, 5, 2, 0), the first
                  `max_reports` per system, in `result.reports` (device tensors `count` [B], `code` [R, B], `it` [R, 2, B],
                  `val` [R, 4, B] = t, h, err1, err2; `.numpy()["reports"]` gives the reference's keys, one row per system).
    return_substeps : as in the reference's solve_ivp_custom (radauDAE.py:992, 1175-1191): also return the collocation points of
                  every step, `result.tsub` [B, 3 K] and `result.ysub` [B, n, 3 K] (implies dense_output=True; see
                  `BatchedOdeSolution.substeps`).
    dense_output, max_dense_steps : record every accepted step (up to `max_dense_steps` per system, required
                  with dense_output=True: the record takes 8 (4 n + 1) bytes per step and system) and return
                  `result.sol`, a BatchedOdeSolution.  Times known in advance are cheaper through `t_eval`.
    """
    import torch

    if not (method is RadauDAE or method == "RadauDAE"):
        raise ValueError("`method` must be RadauDAE.")
    if callable(events):
        events = (events,)
    device_events = None
    if (events and all(isinstance(ev, (LinearEvent, FunctorEvent)) for ev in events) and get_problem(problem).name != "npendulum"
            and get_problem(problem).band is None):
        device_events, events = tuple(events), None       # located by the kernel; nothing to do on the host
        if int(max_events) < 1:
            raise ValueError("`max_events` (occurrences recorded per event and system) must be at least 1.")
    if (events or return_substeps) and not dense_output:
        dense_output = True                    # Python event functions are located on the per-step record; so are the stage points
    if dense_output and (max_dense_steps is None or int(max_dense_steps) < 1):
        raise ValueError("dense_output=True (and events) need `max_dense_steps` (steps recorded per system).")
    if not torch.cuda.is_available():
        raise RuntimeError("solve_ivp_batched needs a CUDA device; there is no CPU fallback.")
    lib = _abi.load(get_problem(problem).lib_path)      # the stock library, or the plugin build of a user problem
    dev = torch.device(device) if device is not None else (
        y0.device if isinstance(y0, torch.Tensor) and y0.is_cuda else torch.device("cuda", torch.cuda.current_device()))

    hb = prepare(problem, t_span, tuple(y0.shape), t_eval=t_eval, params=params, **options)
    prob, n, B, T = hb.problem, hb.n, hb.B, hb.t_eval.size
    if device_events:
        hb.events = pack_linear_events(device_events, n, prob.n_event_functions)
    y0_d = _as_device_f64(y0, dev, torch)
    if not torch.isfinite(y0_d).all():
        raise ValueError("All components of the initial state `y0` must be finite.")
    y0_soa = y0_d.t().contiguous()                                         # [n, B]
    if prob.n_params:
        p = prob.default_params(B) if params is None else params
        if tuple(p.shape) != (B, prob.n_params):
            raise ValueError(f"`params` must have shape {(B, prob.n_params)}.")
        par_soa = _as_device_f64(p, dev, torch).t().contiguous()           # [p, B]
    else:
        par_soa = None

    with torch.cuda.device(dev):
        te_d = torch.as_tensor(hb.t_eval).to(dev) if T else None
        y_eval = torch.empty((max(T, 1), n, B), dtype=torch.float64, device=dev)
        n_eval = torch.empty(B, dtype=torch.int32, device=dev)
        t_final = torch.empty(B, dtype=torch.float64, device=dev)
        y_final = torch.empty((n, B), dtype=torch.float64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        counters = torch.empty((_abi.NCOUNTERS, B), dtype=torch.int32, device=dev)
        ptr = {"y0": y0_soa.data_ptr(), "params": par_soa.data_ptr() if par_soa is not None else None,
               "t_eval": te_d.data_ptr() if T else None, "y_eval": y_eval.data_ptr() if T else None,
               "n_eval": n_eval.data_ptr(), "t_final": t_final.data_ptr(), "y_final": y_final.data_ptr(),
               "status": status.data_ptr(), "counters": counters.data_ptr()}
        if dense_output:
            K = int(max_dense_steps)
            dense_t = torch.empty((K, B), dtype=torch.float64, device=dev)
            dense_y = torch.empty((K, n, B), dtype=torch.float64, device=dev)
            dense_q = torch.empty((K, n, 3, B), dtype=torch.float64, device=dev)
            ptr.update(dense_cap=K, dense_t=dense_t.data_ptr(), dense_y=dense_y.data_ptr(), dense_q=dense_q.data_ptr())
        if report:
            R_ = int(max_reports)
            if R_ < 1:
                raise ValueError("`max_reports` must be at least 1.")
            rep_count = torch.zeros(B, dtype=torch.int32, device=dev)
            rep_code = torch.zeros((R_, B), dtype=torch.int32, device=dev)
            rep_it = torch.full((R_, 2, B), -1, dtype=torch.int32, device=dev)
            rep_val = torch.full((R_, 4, B), float("nan"), dtype=torch.float64, device=dev)
            ptr.update(report_cap=R_, report_count=rep_count.data_ptr(), report_code=rep_code.data_ptr(),
                       report_it=rep_it.data_ptr(), report_val=rep_val.data_ptr())
        if device_events:
            E_, cap = len(device_events), int(max_events)
            event_count = torch.zeros((E_, B), dtype=torch.int32, device=dev)
            event_t = torch.full((E_, cap, B), float("nan"), dtype=torch.float64, device=dev)
            event_y = torch.full((E_, cap, n, B), float("nan"), dtype=torch.float64, device=dev)
            ptr.update(event_cap=cap, event_count=event_count.data_ptr(), event_t=event_t.data_ptr(), event_y=event_y.data_ptr())
        batch = hb.struct(ptr)
        ws_bytes = int(lib.brdae_workspace_bytes(C.byref(batch), C.byref(hb.options)))
        ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
        cur = torch.cuda.current_stream(dev)
        cu_stream = stream if stream is not None else cur
        if cu_stream != cur:
            # the inputs above were produced on torch's current stream: order the solve after them, and keep the
            # allocator from reusing any buffer the kernel touches before the user's stream is done with it
            cu_stream.wait_stream(cur)
            for ten in (y0_soa, par_soa, te_d, y_eval, n_eval, t_final, y_final, status, counters, ws):
                if ten is not None:
                    ten.record_stream(cu_stream)
            for name in ("dense_t", "dense_y", "dense_q", "event_count", "event_t", "event_y", "rep_count", "rep_code", "rep_it", "rep_val"):
                ten = locals().get(name)
                if ten is not None:
                    ten.record_stream(cu_stream)
        _abi.check(lib.brdae_solve(C.byref(batch), C.byref(hb.options), ws.data_ptr(), ws_bytes,
                                   C.c_void_p(cu_stream.cuda_stream)), "brdae_solve")
        ws.record_stream(cu_stream)
    res = BatchedOdeResult(hb.t_eval, y_eval[:T], n_eval, status, t_final, y_final, counters)
    res._keepalive = (y0_soa, par_soa, te_d, ws)
    if device_events:
        res.event_count, res.event_t, res.event_y = event_count, event_t, event_y
    if report:
        res.reports = {"count": rep_count, "code": rep_code, "it": rep_it, "val": rep_val}
    if dense_output:
        cu_stream.synchronize()
        y0_h = y0_d.cpu().numpy()
        res.sol = make_ode_solution(hb, y0_h, dense_t.cpu().numpy(), dense_y.cpu().numpy(),
                                    dense_q.cpu().numpy(), status.cpu().numpy(), counters.cpu().numpy())
        if return_substeps:
            res.tsub, res.ysub = res.sol.substeps()
        if events:
            if res.sol.truncated.any():
                raise RuntimeError("`max_dense_steps` is too small for event location: a system took more steps than recorded.")
            res.t_events, res.y_events, term, t_stop, y_stop = find_events(res.sol, events, hb.t0, y0_h)
            res._event_cut = (term, t_stop, y_stop, hb.t_bound >= hb.t0)
    return res


def make_ode_solution(hb, y0, dense_t_kb, dense_y_knb, dense_q_kn3b, status, counters_cb):
    """Kernel-layout step record ([K][B], [K][n][B], [K][n][3][B]) -> BatchedOdeSolution."""
    K = dense_t_kb.shape[0]
    nstep = counters_cb[_abi.COUNTER_NAMES.index("nstep")]
    accepted = nstep - (status == _abi.STATUS_TOO_SMALL_STEP) - (status == _abi.STATUS_LU_FAILED)   # the failing step() call was counted
    return BatchedOdeSolution(hb.t0, y0, np.ascontiguousarray(dense_t_kb.T), np.ascontiguousarray(dense_y_knb.transpose(2, 1, 0)),
                              np.ascontiguousarray(dense_q_kn3b.transpose(3, 0, 1, 2)), np.minimum(accepted, K).astype(np.int64),
                              accepted > K, hb.t_bound >= hb.t0)


class PinnedArray:
    """numpy array over page-locked host memory from brdae_alloc_pinned (freed with the object)."""

    def __init__(self, shape, dtype):
        self._lib = _abi.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _abi.check(self._lib.brdae_alloc_pinned(max(self.nbytes, 1), C.byref(p)), "brdae_alloc_pinned")
        self._ptr = p
        buf = (C.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._ptr:
                self._lib.brdae_free_pinned(self._ptr)
                self._ptr = None
        except Exception:
            pass


def host_buffers(n, B, T, n_params=0, pinned=True):
    """Reusable result buffers for `solve_ivp_batched_host(..., buffers=...)` in the row-per-system
    layout (`y` [B, T, n], `y_final` [B, n], `counters` [B, 9], ...), page-locked when `pinned`
    (PCIe-rate copies, no first-touch page faults on every call)."""
    # "y" is [B][T][n] as the library writes it; solve_ivp_batched_host exposes it as a [B, n, T] view
    spec = {"y": ((B, max(T, 1), n), np.float64), "n_eval": ((B,), np.int32), "t_final": ((B,), np.float64),
            "y_final": ((B, n), np.float64), "status": ((B,), np.int32), "counters": ((B, _abi.NCOUNTERS), np.int32)}
    bufs, keep = {}, []
    for k, (shape, dt) in spec.items():
        if pinned:
            pa = PinnedArray(shape, dt)
            keep.append(pa)
            bufs[k] = pa.array
        else:
            bufs[k] = np.empty(shape, dtype=dt)
    bufs["_owners"] = keep
    bufs["_shape"] = (n, B, T)
    return bufs


def solve_ivp_batched_host(problem, t_span, y0, params=None, *, t_eval=None, buffers=None, **options):
    """Same call with numpy arrays in and out, through `brdae_solve_host_rows`: `y0` [B, n] and
    `params` [B, p] are handed to the library as they are (row per system), results come back as
    `y` [B, n, T] (a transposed view of the library's [B, T, n] buffer) etc.; device allocation, copies and the layout permutations happen inside the
    library.  This is the entry a numpy/scipy user of the reference binds to and what bench.py
    times as the end-to-end path.  `buffers` (from `host_buffers`) lets repeated calls reuse
    page-locked result arrays; the returned arrays are then views of those buffers."""
    lib = _abi.load(get_problem(problem).lib_path)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    hb = prepare(problem, t_span, y0.shape, t_eval=t_eval, params=params, **options)
    prob, n, B, T = hb.problem, hb.n, hb.B, hb.t_eval.size
    # scipy's check (base.py:14-17: all components of y0 finite).  The thread-per-system entry makes it on the device copy of y0
    # (error code BRDAE_ERR_NONFINITE_Y0 -> the same ValueError below: summing 40 MB on one host core was 4 % of an end-to-end
    # step of the 1M-pendulum ensemble); the chunked entry of the chain / band problems keeps the host test.  A finite sum
    # proves that every entry is finite; the exact element-wise test only runs if the sum overflowed or found one.
    large = hb.problem.name == "npendulum" or hb.problem.band is not None
    if large and not np.isfinite(y0.sum()) and not np.isfinite(y0).all():
        raise ValueError("All components of the initial state `y0` must be finite.")
    if buffers is None:
        buffers = host_buffers(n, B, T, prob.n_params, pinned=False)
    out = buffers
    if out["_shape"] != (n, B, T):
        raise ValueError("`buffers` were made for another problem size.")
    par = None
    if prob.n_params:
        par = prob.default_params(B) if params is None else np.ascontiguousarray(params, dtype=np.float64)
        if par.shape != (B, prob.n_params):
            raise ValueError(f"`params` must have shape {(B, prob.n_params)}.")
    ptr = {"y0": y0.ctypes.data, "params": par.ctypes.data if par is not None else None,
           "t_eval": hb.t_eval.ctypes.data if T else None, "y_eval": out["y"].ctypes.data if T else None,
           "n_eval": out["n_eval"].ctypes.data, "t_final": out["t_final"].ctypes.data,
           "y_final": out["y_final"].ctypes.data, "status": out["status"].ctypes.data,
           "counters": out["counters"].ctypes.data}
    batch = hb.struct(ptr)
    rc = lib.brdae_solve_host_rows(C.byref(batch), C.byref(hb.options))
    if rc == _abi.ERR_NONFINITE_Y0:
        raise ValueError("All components of the initial state `y0` must be finite.")
    _abi.check(rc, "brdae_solve_host_rows")
    res = {"t": hb.t_eval, "y": out["y"][:, :T, :].transpose(0, 2, 1), "n_eval": out["n_eval"], "status": out["status"],
           "success": out["status"] >= 0, "t_final": out["t_final"], "y_final": out["y_final"],
           "counters": out["counters"]}
    for i, name in enumerate(_abi.COUNTER_NAMES):
        res[name] = out["counters"][:, i]
    return res


def fp64_peak_tflops(ms_target: float = 50.0) -> float:
    """Measured FP64 FMA throughput of the current device in TFLOP/s (roofline denominator)."""
    lib = _abi.load()
    f = C.c_double(0.0)
    _abi.check(lib.brdae_fp64_peak_probe(C.c_double(ms_target), C.byref(f), None), "brdae_fp64_peak_probe")
    return f.value / 1e12
