"""ORACLE (test infrastructure, NOT product code) — numpy restatement of RadauDAE.

This file restates, for one system at a time and on the CPU, the algorithm of the
reference solver `scipyDAE/radauDAE.py` (class ``RadauDAE`` plus the ``solve_ivp`` driver
loop around it).  It exists only so that the CUDA path can be checked and timed against
something that travels to the GPU box (``/root/reference`` does not).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it; the product package must never do so.

Parity pin: ``tests/test_oracle_numpy.py`` checks this restatement
  * bit-for-bit against the live reference whenever ``/root/reference`` is importable
    (this container), and
  * against the committed fixtures ``tests/golden/*.npz`` that ``oracle/gen_golden.py``
    produced by running the reference itself.

The linear algebra is the same third-party code the reference calls
(`scipy.linalg.lu_factor/lu_solve` -> LAPACK d/zgetrf, d/zgetrs; scipy 1.18.1, numpy 2.3.5),
and every floating-point expression keeps the reference's operand order, so on one machine
the two produce identical bits.

Reference map (file:line in /root/reference/scipyDAE/radauDAE.py):
  tableau constants ............ 11-48      -> module constants below
  predict_factor ............... 52-98      -> _gustafsson_factor
  RadauDAE.__init__ ............ 263-422    -> _Setup.__init__
  _step_impl ................... 515-791    -> _advance_one_step
  solve_collocation_system ..... 793-949    -> _simplified_newton
  dense output ................. 951-982    -> _Cubic
  driver loop / t_eval slicing . 1079-1157 (and scipy ivp.py:659-732) -> solve()
Only the dense (non-sparse) branch and the callable/constant Jacobian modes are restated.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import lu_factor, lu_solve

# ---------------------------------------------------------------------------------------
# Radau IIA (3 stages, order 5) constants — radauDAE.py:11-48
# ---------------------------------------------------------------------------------------
_SQ6 = 6 ** 0.5
C = np.array([(4 - _SQ6) / 10, (4 + _SQ6) / 10, 1])
E = np.array([-13 - 7 * _SQ6, -13 + 7 * _SQ6, -1]) / 3
MU_REAL = 3 + 3 ** (2 / 3) - 3 ** (1 / 3)
MU_COMPLEX = (3 + 0.5 * (3 ** (1 / 3) - 3 ** (2 / 3))
              - 0.5j * (3 ** (5 / 6) + 3 ** (7 / 6)))
T = np.array([
    [0.09443876248897524, -0.14125529502095421, 0.03002919410514742],
    [0.25021312296533332, 0.20412935229379994, -0.38294211275726192],
    [1, 1, 0]])
TI = np.array([
    [4.17871859155190428, 0.32768282076106237, 0.52337644549944951],
    [-4.17871859155190428, -0.32768282076106237, 0.47662355450055044],
    [0.50287263494578682, -2.57192694985560522, 0.59603920482822492]])
TI_REAL = TI[0]
TI_COMPLEX = TI[1] + 1j * TI[2]
P = np.array([
    [13 / 3 + 7 * _SQ6 / 3, -23 / 3 - 22 * _SQ6 / 3, 10 / 3 + 5 * _SQ6],
    [13 / 3 - 7 * _SQ6 / 3, -23 / 3 + 22 * _SQ6 / 3, 10 / 3 - 5 * _SQ6],
    [1 / 3, -8 / 3, 10 / 3]])

EPS = np.finfo(float).eps

STATUS_FINISHED = 0
STATUS_FAILED = -1        # solver.status == 'failed'
STATUS_NMAX_STEP = 2      # solve_ivp_custom only (radauDAE.py:1081-1084)

MSG_TOO_SMALL = "Required step size is less than spacing between numbers."

# sub-codes that refine STATUS_FAILED the way the batched C-ABI reports them
FAIL_NONE = 0
FAIL_TOO_SMALL_STEP = 1
FAIL_LU = 2


def _rms(x):
    """scipy.integrate._ivp.common.norm (scipy common.py:63-65)."""
    return np.linalg.norm(x) / x.size ** 0.5


def _gustafsson_factor(h_abs, h_abs_old, err, err_old, predictive):
    """radauDAE.py:52-98."""
    if not predictive:
        mult = 1
    elif err_old is None or h_abs_old is None or err == 0:
        mult = 1
    else:
        mult = h_abs / h_abs_old * (err_old / err) ** 0.25
    with np.errstate(divide='ignore'):
        return min(1, mult) * err ** -0.25


class _Cubic:
    """RadauDenseOutput (radauDAE.py:959-982) without the scipy base class."""

    def __init__(self, t_old, t, y_old, Q):
        self.t_old = t_old
        self.h = t - t_old
        self.Q = Q
        self.y_old = y_old

    def __call__(self, t):
        t = np.asarray(t)
        x = (t - self.t_old) / self.h
        if t.ndim == 0:
            p = np.cumprod(np.tile(x, 3))
        else:
            p = np.cumprod(np.tile(x, (3, 1)), axis=0)
        y = np.dot(self.Q, p)
        if y.ndim == 2:
            y += self.y_old[:, None]
        else:
            y += self.y_old
        return y


@dataclass
class Counters:
    nfev: int = 0
    njev: int = 0
    nlu: int = 0
    nstep: int = 0
    naccpt: int = 0
    nrejct: int = 0
    nfailed: int = 0
    nlusolve: int = 0
    nlusolve_err: int = 0

    def as_array(self):
        return np.array([self.nfev, self.njev, self.nlu, self.nstep, self.naccpt,
                         self.nrejct, self.nfailed, self.nlusolve, self.nlusolve_err],
                        dtype=np.int32)


COUNTER_NAMES = ("nfev", "njev", "nlu", "nstep", "naccpt", "nrejct", "nfailed",
                 "nlusolve", "nlusolve_err")


@dataclass
class OracleResult:
    t: np.ndarray
    y: np.ndarray
    status: int
    fail_code: int
    message: str | None
    counters: Counters
    t_final: float
    y_final: np.ndarray
    steps_t: list = field(default_factory=list)     # accepted step end times
    interpolants: list = field(default_factory=list)   # dense_output=True: the per-step RadauDenseOutput objects

    def sol(self, t):
        """OdeSolution of solve_ivp(dense_output=True) (scipy base.py OdeSolution, ascending or descending
        ts): segment by searchsorted, clipped to the first/last step."""
        ts = np.array([self.interpolants[0].t_old] + [c.t_old + c.h for c in self.interpolants])
        t = np.atleast_1d(np.asarray(t, dtype=float))
        if ts[-1] >= ts[0]:
            ind = np.searchsorted(ts, t, side="left")
        else:
            ind = np.searchsorted(-ts, -t, side="right")
        seg = np.clip(ind - 1, 0, len(self.interpolants) - 1)
        return np.stack([self.interpolants[k](x) for k, x in zip(seg, t)], axis=1)
    log: list = field(default_factory=list)         # per-attempt records when trace=True

    @property
    def success(self):
        return self.status >= 0


class _Setup:
    """Everything RadauDAE.__init__ derives from its arguments (radauDAE.py:263-422)."""

    def __init__(self, fun, t0, y0, t_bound, *, jac, mass_matrix=None, var_index=None,
                 rtol=1e-3, atol=1e-6, first_step=None, max_step=np.inf,
                 factor_on_non_convergence=0.5, safety_factor=0.9, newton_tol=None,
                 constant_dt=False, max_newton_ite=6, min_factor=0.2, max_factor=10,
                 max_bad_ite=None, scale_residuals=True, scale_newton_norm=True,
                 scale_error=True, zero_algebraic_error=False,
                 jacobianRecomputeFactor=1e-3, bAlwaysApply2ndEstimate=True,
                 bUsePredictiveController=True,
                 bUsePredictiveNewtonStoppingCriterion=True, bUseExtrapolatedGuess=True,
                 bPerformAllNewtonIterations=False):
        self.cnt = Counters()
        y0 = np.asarray(y0)
        if not np.issubdtype(y0.dtype, np.floating):
            y0 = y0.astype(float)
        y0 = y0.astype(float, copy=False)
        if y0.ndim != 1:
            raise ValueError("`y0` must be 1-dimensional.")
        if not np.isfinite(y0).all():
            raise ValueError("All components of the initial state `y0` must be finite.")
        self.n = n = y0.size
        self.t, self.y, self.t_bound = t0, y0, t_bound
        self.direction = np.sign(t_bound - t0) if t_bound != t0 else 1

        def counted_fun(t, y):
            self.cnt.nfev += 1
            return np.asarray(fun(t, y), dtype=float)
        self.fun = counted_fun

        if max_step <= 0:
            raise ValueError("`max_step` must be positive.")
        self.max_step = max_step
        if np.any(rtol < 100 * EPS):
            rtol = np.maximum(rtol, 100 * EPS)
        atol = np.asarray(atol)
        if atol.ndim > 0 and atol.shape != (n,):
            raise ValueError("`atol` has wrong shape.")
        if np.any(atol < 0):
            raise ValueError("`atol` must be positive.")
        self.rtol, self.atol = rtol, atol

        self.f = self.fun(self.t, self.y)                                   # :304
        if first_step is None:                                              # :307-315
            if mass_matrix is None:
                raise TypeError("first_step=None without a mass matrix hits the "
                                "select_initial_step signature mismatch of the reference "
                                "under scipy>=1.11 (SURVEY 4.3); pass first_step")
            self.h_abs = self.direction * min(1e-6, abs(t_bound - t0))
        else:
            if first_step <= 0:
                raise ValueError("`first_step` must be positive.")
            if first_step > np.abs(t_bound - t0):
                raise ValueError("`first_step` exceeds bounds.")
            self.h_abs = first_step
        self.h_abs_old = None
        self.err_old = None

        if newton_tol is None:                                              # :319-324
            self.newton_tol = max(10 * EPS / rtol, min(0.03, rtol ** 0.5))
        else:
            self.newton_tol = newton_tol

        # Jacobian modes (radauDAE.py:482-511): callable, or constant matrix
        if jac is None:
            # finite differences, radauDAE.py:468-481: scipy's own num_jac (third-party, common.py:268-452)
            # on the uncounted vectorised wrapper of scipy base.py:143-149, threshold = atol, the
            # adaptive column factors carried from call to call
            from scipy.integrate._ivp.common import num_jac
            self.jac_factor = None

            def fun_vectorized(t, y):
                f = np.empty_like(y)
                for i, yi in enumerate(y.T):
                    f[:, i] = np.asarray(fun(t, yi), dtype=float)
                return f

            def fd_jac(t, y, f):
                self.cnt.njev += 1
                J, self.jac_factor = num_jac(fun_vectorized, t, y, f, self.atol, self.jac_factor, None)
                return J
            J = fd_jac(t0, y0, self.f)
            self.jac = fd_jac
        elif callable(jac):
            J = np.asarray(jac(t0, y0), dtype=float)
            self.cnt.njev = 1

            def counted_jac(t, y, _f=None):
                self.cnt.njev += 1
                return np.asarray(jac(t, y), dtype=float)
            self.jac = counted_jac
        else:
            J = np.asarray(jac, dtype=float)
            self.jac = None
        if J.shape != (n, n):
            raise ValueError(f"`jac` is expected to have shape {(n, n)}, but actually has {J.shape}.")
        self.J = J

        self.I = np.identity(n)
        if mass_matrix is None:                                             # :447-462
            self.M = self.I
        elif callable(mass_matrix):
            raise ValueError("`mass_matrix` should be a constant matrix, but is callable")
        else:
            self.M = np.asarray(mass_matrix, dtype=float)
            if self.M.shape != (n, n):
                raise ValueError("`mass_matrix` is expected to have shape {}, but actually has {}."
                                 .format((n, n), self.M.shape))

        if var_index is None:                                               # :389-398
            self.var_index = np.zeros((n,))
        else:
            self.var_index = np.asarray(var_index)
            assert self.var_index.ndim == 1 and self.var_index.size == n
        self.var_exp = self.var_index - 1
        self.var_exp[self.var_exp < 0] = 0
        if max_bad_ite is not None:                                         # :400-410
            self.nmax_bad = max_bad_ite
        else:
            self.nmax_bad = 1 if np.any(self.var_index > 1) else 0
        self.alg_idx = np.where(self.var_index != 0)[0]

        self.scale_residuals = scale_residuals
        self.scale_newton_norm = scale_newton_norm
        self.scale_error = scale_error
        self.zero_algebraic_error = zero_algebraic_error
        self.hscale = None if scale_residuals else self.I

        self.maxit = max_newton_ite
        self.min_factor, self.max_factor = min_factor, max_factor
        self.nonconv_factor = factor_on_non_convergence
        self.safety_factor = safety_factor
        self.constant_dt = constant_dt
        self.jac_recompute = jacobianRecomputeFactor
        self.always_2nd = bAlwaysApply2ndEstimate
        self.predictive_ctrl = bUsePredictiveController
        self.predictive_newton = bUsePredictiveNewtonStoppingCriterion
        self.extrapolate = bUseExtrapolatedGuess
        self.all_newton = bPerformAllNewtonIterations

        self.current_jac = True                                             # :416
        self.LU_real = None
        self.LU_complex = None
        self.sol = None
        self.t_old = None
        self.trace = None

    # LU closures (radauDAE.py:364-370)
    def lu(self, A):
        self.cnt.nlu += 1
        return lu_factor(A, overwrite_a=True)

    def lusolve(self, LU, b):
        self.cnt.nlusolve += 1
        return lu_solve(LU, b, overwrite_b=True)


def _simplified_newton(s: _Setup, t, y, h, Z0, norm_scale, LU_real, LU_complex, rscale):
    """radauDAE.py:793-949.  Returns (converged, n_iter, n_bad, Z, rate)."""
    n = y.shape[0]
    M_real = MU_REAL / h
    M_complex = MU_COMPLEX / h
    W = TI.dot(Z0)
    Z = Z0
    F = np.empty((3, n))
    ch = h * C
    dW_norm_old = None
    dW = np.empty_like(W)
    converged = False
    rate = None
    nbad = 0
    tol = s.newton_tol
    k = -1
    for k in range(s.maxit):
        for i in range(3):
            F[i] = s.fun(t + ch[i], y + Z[i])
        if not np.all(np.isfinite(F)):
            break
        f_real = F.T.dot(TI_REAL) - M_real * s.M.dot(W[0])
        f_complex = F.T.dot(TI_COMPLEX) - M_complex * s.M.dot(W[1] + 1j * W[2])
        f_real = rscale @ f_real
        f_complex = rscale @ f_complex
        dW_real = s.lusolve(LU_real, f_real)
        dW_complex = s.lusolve(LU_complex, f_complex)
        s.cnt.nlusolve -= 1
        dW[0] = dW_real
        dW[1] = dW_complex.real
        dW[2] = dW_complex.imag
        dW_norm = _rms(dW / norm_scale)
        W += dW
        Z = T.dot(W)
        if dW_norm_old is not None:
            rate = dW_norm / dW_norm_old
            dW_true = rate / (1 - rate) * dW_norm
            dW_true_max = rate ** (s.maxit - k) / (1 - rate) * dW_norm
        if dW_norm < tol:
            converged = True
            break
        if rate is not None:
            if rate >= 1:
                if rate < 100:
                    if nbad < s.nmax_bad:
                        nbad += 1
                        continue
                if not s.all_newton:
                    break
            if dW_true < tol:
                converged = True
                break
            if (dW_true_max > tol) and s.predictive_newton:
                if nbad < s.nmax_bad:
                    nbad += 1
                    continue
                if not s.all_newton:
                    break
        dW_norm_old = dW_norm
    return converged, k + 1, nbad, Z, rate


def _advance_one_step(s: _Setup):
    """radauDAE.py:515-791.  Returns (ok, fail_code, message)."""
    t, y, f = s.t, s.y, s.f
    n = y.size
    cnt = s.cnt
    cnt.nstep += 1
    max_step, atol, rtol = s.max_step, s.atol, s.rtol

    min_step = max(1e-20, 10 * np.abs(np.nextafter(t, s.direction * np.inf) - t))
    if s.h_abs > max_step:
        h_abs, h_abs_old, err_old = max_step, None, None
    elif s.h_abs < min_step:
        h_abs, h_abs_old, err_old = min_step, None, None
    else:
        h_abs, h_abs_old, err_old = s.h_abs, s.h_abs_old, s.err_old

    if abs(s.t_bound - (t + s.direction * h_abs)) < 1e-2 * h_abs:
        h_abs = abs(s.t_bound - t)
        s.LU_real = None
        s.LU_complex = None

    J, LU_real, LU_complex = s.J, s.LU_real, s.LU_complex
    current_jac = s.current_jac
    rejected = False
    accepted = False
    while not accepted:
        if h_abs < min_step:
            return False, FAIL_TOO_SMALL_STEP, MSG_TOO_SMALL
        h = h_abs * s.direction
        t_new = t + h
        if s.direction * (t_new - s.t_bound) > 0:
            t_new = s.t_bound
        h = t_new - t
        h_abs = np.abs(h)

        if (s.sol is None) or not s.extrapolate:
            Z0 = np.zeros((3, y.shape[0]))
        else:
            Z0 = s.sol(t + h * C).T - y
        newton_scale = atol + np.abs(y) * rtol
        if s.scale_newton_norm:
            newton_scale = newton_scale / (h ** s.var_exp)

        converged = False
        while not converged:
            if LU_real is None or LU_complex is None:
                if s.scale_residuals:
                    s.hscale = np.diag(h ** (-np.minimum(1, s.var_index)))
                try:
                    LU_real = s.lu(s.hscale @ (MU_REAL / h * s.M - J))
                    LU_complex = s.lu(s.hscale @ (MU_COMPLEX / h * s.M - J))
                    cnt.nlu -= 1
                except ValueError as e:
                    return False, FAIL_LU, 'LU decomposition failed ({})'.format(e)
            converged, n_iter, n_bad, Z, rate = _simplified_newton(
                s, t, y, h, Z0, newton_scale, LU_real, LU_complex, s.hscale)
            safety = s.safety_factor * (2 * s.maxit + 1) / (2 * s.maxit + n_iter)
            if not converged:
                if current_jac:
                    break
                J = s.jac(t, y, f)
                current_jac = True
                LU_real = None
                LU_complex = None

        if not converged:
            cnt.nfailed += 1
            if s.trace is not None:
                s.trace.append(("nonconv", t, h, n_iter, n_bad, np.nan))
            h_abs *= s.nonconv_factor
            LU_real = None
            LU_complex = None
            continue

        y_new = y + Z[-1]
        if s.constant_dt:
            accepted = True
            error_norm = 0.
        else:
            ZE = Z.T.dot(E) / h
            err_scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            if s.scale_error:
                err_scale = err_scale / (h ** s.var_exp)
            error = s.lusolve(LU_real, f + s.M.dot(ZE))
            cnt.nlusolve_err += 1
            cnt.nlusolve -= 1
            if s.zero_algebraic_error:
                error[s.alg_idx] = 0.
                error_norm = np.linalg.norm(error / err_scale) / (n - s.alg_idx.size) ** 0.5
            else:
                error_norm = _rms(error / err_scale)
            if (s.always_2nd and error_norm > 1) or (rejected and error_norm > 1):
                error = s.lusolve(LU_real, s.fun(t, y + error) + s.M.dot(ZE))
                cnt.nlusolve_err += 1
                cnt.nlusolve -= 1
                if s.zero_algebraic_error:
                    error[s.alg_idx] = 0.
                    error_norm = np.linalg.norm(error / err_scale) / (n - s.alg_idx.size) ** 0.5
                else:
                    error_norm = _rms(error / err_scale)
            if error_norm > 1:
                factor = _gustafsson_factor(h_abs, h_abs_old, error_norm, err_old,
                                            s.predictive_ctrl)
                if s.trace is not None:
                    s.trace.append(("reject", t, h, n_iter, n_bad, float(error_norm)))
                h_abs *= max(s.min_factor, safety * factor)
                LU_real = None
                LU_complex = None
                rejected = True
                cnt.nrejct += 1
            else:
                if s.trace is not None:
                    s.trace.append(("accept", t, h, n_iter, n_bad, float(error_norm)))
                cnt.naccpt += 1
                accepted = True

    recompute_jac = s.jac is not None and n_iter > 2 and rate > s.jac_recompute
    if s.constant_dt:
        factor = s.max_step / h_abs
    else:
        factor = _gustafsson_factor(h_abs, h_abs_old, error_norm, err_old, s.predictive_ctrl)
        factor = min(s.max_factor, safety * factor)
    if not recompute_jac and factor < 1.2:
        factor = 1
    else:
        LU_real = None
        LU_complex = None

    f_new = s.fun(t_new, y_new)
    if recompute_jac:
        J = s.jac(t_new, y_new, f_new)
        current_jac = True
    elif s.jac is not None:
        current_jac = False

    s.h_abs_old = s.h_abs
    s.err_old = error_norm
    s.h_abs = h_abs * factor
    y_old = y
    s.t, s.y, s.f = t_new, y_new, f_new
    s.LU_real, s.LU_complex = LU_real, LU_complex
    s.current_jac = current_jac
    s.J = J
    s.t_old = t
    s.sol = _Cubic(t, t_new, y_old, np.dot(Z.T, P))
    return True, FAIL_NONE, None


def solve(fun, t_span, y0, *, t_eval=None, nmax_step=np.inf, trace=False, dense_output=False, **options):
    """Driver loop with t_eval slicing (scipy ivp.py:659-732; radauDAE.py:1079-1157).

    Returns an OracleResult; `y` is laid out [n, T_reached] like OdeResult.y.
    """
    t0, tf = float(t_span[0]), float(t_span[1])
    if t_eval is not None:
        t_eval = np.asarray(t_eval)
        if t_eval.ndim != 1:
            raise ValueError("`t_eval` must be 1-dimensional.")
        if np.any(t_eval < min(t0, tf)) or np.any(t_eval > max(t0, tf)):
            raise ValueError("Values in `t_eval` are not within `t_span`.")
        d = np.diff(t_eval)
        if tf > t0 and np.any(d <= 0) or tf < t0 and np.any(d >= 0):
            raise ValueError("Values in `t_eval` are not properly sorted.")
        if tf > t0:
            te_i = 0
        else:
            t_eval = t_eval[::-1]
            te_i = t_eval.shape[0]

    s = _Setup(fun, t0, y0, tf, **options)
    if trace:
        s.trace = []
    if t_eval is None:
        ts, ys = [t0], [s.y]
    else:
        ts, ys = [], []
    steps_t = []
    interpolants = []
    status = None
    fail_code = FAIL_NONE
    message = None
    nsteps = 0
    while status is None:
        nsteps += 1
        if nsteps > nmax_step:
            status = STATUS_NMAX_STEP
            message = "Too many time steps have been performed"
            break
        # OdeSolver.step (scipy base.py:179-210)
        if s.n == 0 or s.t == s.t_bound:
            s.t_old = s.t
            s.t = s.t_bound
            finished = True
            const_out = True
        else:
            ok, fail_code, message = _advance_one_step(s)
            if not ok:
                status = STATUS_FAILED
                break
            finished = s.direction * (s.t - s.t_bound) >= 0
            const_out = s.t == s.t_old
        if finished:
            status = STATUS_FINISHED
        t = s.t
        steps_t.append(t)
        if dense_output and not const_out:
            interpolants.append(s.sol)
        if t_eval is None:
            ts.append(t)
            ys.append(s.y)
        else:
            if s.direction > 0:
                te_new = np.searchsorted(t_eval, t, side='right')
                step_pts = t_eval[te_i:te_new]
            else:
                te_new = np.searchsorted(t_eval, t, side='left')
                step_pts = t_eval[te_new:te_i][::-1]
            if step_pts.size > 0:
                if const_out:      # ConstantDenseOutput (scipy base.py:262-283)
                    vals = np.empty((s.n, step_pts.size))
                    vals[:] = s.y[:, None]
                else:
                    vals = s.sol(step_pts)
                ts.append(step_pts)
                ys.append(vals)
                te_i = te_new

    if status == STATUS_FINISHED:
        message = "The solver successfully reached the end of the integration interval."
    if t_eval is None:
        t_out = np.array(ts)
        y_out = np.vstack(ys).T
    elif ts:
        t_out = np.hstack(ts)
        y_out = np.hstack(ys)
    else:
        t_out = np.empty(0)
        y_out = np.empty((s.n, 0))
    return OracleResult(t=t_out, y=y_out, status=status, fail_code=fail_code, message=message,
                        counters=s.cnt, t_final=float(s.t), y_final=np.array(s.y),
                        steps_t=steps_t, interpolants=interpolants, log=s.trace or [])
