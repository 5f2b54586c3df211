"""ORACLE (test infrastructure, NOT product code) — the reference's test problems, parametrised.

Each builder returns ``(fun, jac, mass, var_index)`` for ONE system with the parameter
vector used by the batched CUDA path (`dae_scipy_b200.problems`).  The expressions keep the
operand order of the reference closures so that, with the reference's default parameters,
`fun`/`jac` return the same bits as the reference's own closures:

  pendulum (index 3/2/1) .... scipyDAE/tests/pendulum.py:129-184, IC :225-231
  robertson DAE ............. scipyDAE/tests/robertson.py:78-87
  transistor amplifier ...... scipyDAE/tests/transistor_amplifier.py:33-56
  n-pendulum chain (index 3)  scipyDAE/tests/recursive_pendulum.py:56-126, IC :168-177

The reference has no analytic Jacobian for the transistor (complex step, :61-65) nor for the
n-pendulum (`jac_dae=None`, :128); the analytic ones here are checked against complex-step /
finite differences in tests/test_oracle_numpy.py and are handed to the reference when the
golden fixtures are generated (oracle/gen_golden.py), so both sides use the same Jacobian.
"""
from __future__ import annotations

import numpy as np

GRAVITY = 9.81

PENDULUM_DEFAULTS = dict(m=1.0, g=9.81, r0=1.0)
ROBERTSON_DEFAULTS = dict(k1=0.04, k2=1e4, k3=3e7)
TRANSISTOR_DEFAULTS = dict(C1=1e-6, C2=2e-6, C3=3e-6, R0=1e3, R1=9e3, R2=9e3, R3=9e3,
                           R4=9e3, R5=9e3, Ub=6.0, Ua=0.4, w=200 * np.pi)
TRANSISTOR_ORDER = ("C1", "C2", "C3", "R0", "R1", "R2", "R3", "R4", "R5", "Ub", "Ua", "w")


# ----------------------------------------------------------------------------- pendulum
def pendulum_ic(theta_0, theta_dot0, m=1.0, g=9.81, r0=1.0):
    """Consistent initial condition, pendulum.py:225-231."""
    x0 = r0 * np.sin(theta_0)
    y0 = -r0 * np.cos(theta_0)
    vx0 = r0 * theta_dot0 * np.cos(theta_0)
    vy0 = r0 * theta_dot0 * np.sin(theta_0)
    lbda_0 = (m * r0 * theta_dot0 ** 2 + m * g * np.cos(theta_0)) / r0
    return np.array([x0, y0, vx0, vy0, lbda_0])


def pendulum(index=3, m=1.0, g=9.81, r0=1.0):
    mass = np.eye(5)
    mass[-1, -1] = 0
    if index == 3:
        def fun(t, X):
            x, y, vx, vy, lbda = X[0], X[1], X[2], X[3], X[4]
            return np.array([vx, vy, -x * lbda / m, -g - (y * lbda) / m,
                             x ** 2 + y ** 2 - r0 ** 2])

        def jac(t, X):
            x, y, vx, vy, lbda = X[0], X[1], X[2], X[3], X[4]
            return np.array([[0., 0., 1., 0., 0.],
                             [0., 0., 0., 1., 0.],
                             [-lbda / m, 0., 0., 0., -x / m],
                             [0., -lbda / m, 0., 0., -y / m],
                             [2 * x, 2 * y, 0., 0., 0.]])
        var_index = np.array([0, 0, 0, 0, 3])
    elif index == 2:
        def fun(t, X):
            x, y, vx, vy, lbda = X[0], X[1], X[2], X[3], X[4]
            return np.array([vx, vy, -x * lbda / m, -g - (y * lbda) / m, x * vx + y * vy])

        def jac(t, X):
            x, y, vx, vy, lbda = X[0], X[1], X[2], X[3], X[4]
            return np.array([[0., 0., 1., 0., 0.],
                             [0., 0., 0., 1., 0.],
                             [-lbda / m, 0., 0., 0., -x / m],
                             [0., -lbda / m, 0., 0., -y / m],
                             [vx, vy, x, y, 0.]])
        var_index = np.array([0, 0, 0, 0, 2])
    elif index == 1:
        def fun(t, X):
            x, y, vx, vy, lbda = X[0], X[1], X[2], X[3], X[4]
            return np.array([vx, vy, -x * lbda / m, -g - (y * lbda) / m,
                             lbda * (x ** 2 + y ** 2) / m + g * y - (vx ** 2 + vy ** 2)])

        def jac(t, X):
            x, y, vx, vy, lbda = X[0], X[1], X[2], X[3], X[4]
            return np.array([[0., 0., 1., 0., 0.],
                             [0., 0., 0., 1., 0.],
                             [-lbda / m, 0., 0., 0., -x / m],
                             [0., -lbda / m, 0., 0., -y / m],
                             [2 * x * lbda / m, 2 * y * lbda / m + g, -2 * vx, -2 * vy,
                              (x ** 2 + y ** 2) / m]])
        var_index = np.array([0, 0, 0, 0, 1])
    else:
        raise ValueError("pendulum index must be 1, 2 or 3")
    return fun, jac, mass, var_index


# ---------------------------------------------------------------------------- robertson
def robertson(k1=0.04, k2=1e4, k3=3e7):
    mass = np.array([[1., 0., 0.], [0., 1., 0.], [0., 0., 0.]])

    def fun(t, y):
        return np.array([-k1 * y[0] + k2 * y[1] * y[2],
                         k1 * y[0] - k2 * y[1] * y[2] - k3 * (y[1] ** 2),
                         y[0] + y[1] + y[2] - 1])

    def jac(t, y):
        return np.array([[-k1, k2 * y[2], k2 * y[1]],
                         [k1, -k2 * y[2] - 2 * k3 * y[1], -k2 * y[1]],
                         [1, 1, 1], ])
    return fun, jac, mass, np.array([0, 0, 1])


def robertson_ode(k1=0.04, k2=1e4, k3=3e7):
    """ODE formulation, robertson.py:69-77 (third equation y2' = k3 y1^2, identity mass)."""
    mass = np.eye(3)

    def fun(t, y):
        return np.array([-k1 * y[0] + k2 * y[1] * y[2],
                         k1 * y[0] - k2 * y[1] * y[2] - k3 * (y[1] ** 2),
                         k3 * (y[1] ** 2)])

    def jac(t, y):
        return np.array([[-k1, k2 * y[2], k2 * y[1]],
                         [k1, -k2 * y[2] - 2 * k3 * y[1], -k2 * y[1]],
                         [0., 2 * k3 * y[1], 0.], ])
    return fun, jac, mass, np.array([0, 0, 0])


ROBERTSON_Y0 = np.array([1., 0., 0.])


# ------------------------------------------------------------------------- batch reactor
def reactor(k=1.47, Ca0=0.0209, Cb0=0.0209 / 3):
    """Index-1 batch-reactor DAE of the reference's issue reproduction, tests/reactor.py:74-106 (Cc0 = 0):
    Ca' = -k Ca Cb, 0 = Cb - (Cb0 - (Ca0 - Ca)), 0 = Cc - (Ca0 - Ca).  The script has no Jacobian (jac=None);
    the analytic one here is for the like-for-like runs."""
    mass = np.diag([1., 0., 0.])

    def fun(t, y):
        return np.array([-k * y[0] * y[1],
                         y[1] - (Cb0 - (Ca0 - y[0])),
                         y[2] - (Ca0 - y[0])])

    def jac(t, y):
        return np.array([[-k * y[1], -k * y[0], 0.],
                         [-1., 1., 0.],
                         [1., 0., 1.]])
    return fun, jac, mass, np.array([0, 1, 1])


# ------------------------------------------------------------------------- reaction-diffusion rod (user band problem)
def rod(D=0.05, k=2.0, a=0.5, w=6.0, n=48):
    """The band example of dae_scipy_b200.ensembles (ROD_RHS_ROW) as a numpy closure: method of lines with algebraic boundary
    rows; no analytic Jacobian (jac=None with a tridiagonal `jac_sparsity`, like the reference's own chain script)."""
    diag = np.ones(n)
    diag[0] = diag[-1] = 0.0
    s = float(n - 1) * float(n - 1)

    def fun(t, y):
        f = np.empty_like(y)
        f[0] = y[0] - (1.0 + a * np.sin(w * t))
        f[-1] = y[-1] - y[-2]
        f[1:-1] = (D * s) * ((y[:-2] - 2.0 * y[1:-1]) + y[2:]) - k * ((y[1:-1] * y[1:-1]) * y[1:-1])
        return f
    vi = np.zeros(n, dtype=int)
    vi[0] = vi[-1] = 1
    return fun, None, np.diag(diag), vi


# ------------------------------------------------------------------------- Jay 1995
def jay(nonlinear=1.0):
    """Index-3 DAE of Jay 1995, jay_dae.py:22-68 (y1, y2, z1, z2, lambda).  The reference has no
    analytic Jacobian (finite differences); the one below is checked by complex step in the tests."""
    nl = bool(nonlinear)
    mass = np.eye(5)
    mass[-1, -1] = 0

    def fun(t, y):
        y1, y2, z1, z2, la = y
        if nl:
            dy3 = -y1 * (y2 ** 2) * (z2 ** 3) * (la ** 2)
        else:
            dy3 = -y1 * (y2 ** 2) * (z2 ** 2) * la
        return np.array([2 * y1 * y2 * z1 * z2, -y1 * y2 * (z2 ** 2), (y1 * y2 + z1 * z2) * la, dy3,
                         y1 * y2 ** 2 - 1.0], dtype=y.dtype)

    def jac(t, y):
        y1, y2, z1, z2, la = y
        J = np.zeros((5, 5))
        J[0] = [2 * y2 * z1 * z2, 2 * y1 * z1 * z2, 2 * y1 * y2 * z2, 2 * y1 * y2 * z1, 0.]
        J[1] = [-y2 * (z2 ** 2), -y1 * (z2 ** 2), 0., -2 * y1 * y2 * z2, 0.]
        J[2] = [y2 * la, y1 * la, z2 * la, z1 * la, y1 * y2 + z1 * z2]
        if nl:
            J[3] = [-(y2 ** 2) * (z2 ** 3) * (la ** 2), -2 * y1 * y2 * (z2 ** 3) * (la ** 2), 0.,
                    -3 * y1 * (y2 ** 2) * (z2 ** 2) * (la ** 2), -2 * y1 * (y2 ** 2) * (z2 ** 3) * la]
        else:
            J[3] = [-(y2 ** 2) * (z2 ** 2) * la, -2 * y1 * y2 * (z2 ** 2) * la, 0., -2 * y1 * (y2 ** 2) * z2 * la,
                    -y1 * (y2 ** 2) * (z2 ** 2)]
        J[4] = [y2 ** 2, 2 * y1 * y2, 0., 0., 0.]
        return J
    return fun, jac, mass, np.array([0, 0, 0, 0, 3])


JAY_Y0 = np.ones(5)


# ------------------------------------------------------------------------- polynomial index-2 DAE
def polydae2(a4=0., a3=0., a2=0., a1=1., a0=0.):
    """validation_dense_output.py:23-59 with p = 2: y' = P'(t) + (z - y), 0 = y - P(t), where
    P(t) = a4 t^4 + ... + a0 (highest power first, like np.polyval)."""
    a = np.array([a4, a3, a2, a1, a0], dtype=float)
    der = np.array([4 * a4, 3 * a3, 2 * a2, a1], dtype=float)
    mass = np.eye(2)
    mass[-1, -1] = 0

    def fun(t, x):
        y, z = x
        return np.array([np.polyval(der, t) + (z - y), y - np.polyval(a, t)])

    def jac(t, x):
        return np.array([[-1., 1.], [1., 0.]])
    return fun, jac, mass, np.array([0, 2])


# ------------------------------------------------------------------------- linear ODE with mass
def linear7(*c):
    """linear_ODE_with_mass_matrix.py:32-39: f(t, x) = c * x componentwise (c = 1 with the mass matrix
    diag(1/eig), c = eig without)."""
    c = np.array(c if c else np.ones(7), dtype=float)

    def fun(t, x):
        return c * x

    def jac(t, x):
        return np.diag(c)
    return fun, jac, np.eye(7), np.zeros(7, dtype=int)


LINEAR7_EIG = np.array([float(i) * (-1) ** (i) for i in range(1, 8)])


# --------------------------------------------------------------------------- transistor
def transistor_mass(C1, C2, C3):
    return -np.array([[-C1, C1, 0, 0, 0],
                      [C1, -C1, 0, 0, 0],
                      [0, 0, -C2, 0, 0],
                      [0, 0, 0, -C3, C3],
                      [0, 0, 0, C3, -C3]])


def transistor_y0(R1=9e3, R2=9e3, Ub=6.0, **_):
    return np.array([0., Ub * R1 / (R1 + R2), Ub * R1 / (R1 + R2), Ub, 0.])


def transistor(C1=1e-6, C2=2e-6, C3=3e-6, R0=1e3, R1=9e3, R2=9e3, R3=9e3, R4=9e3, R5=9e3,
               Ub=6.0, Ua=0.4, w=200 * np.pi):
    mass = transistor_mass(C1, C2, C3)

    def Ue(t):
        return Ua * np.sin(w * t)

    def fun(t, y):
        U1, U2, U3, U4, U5 = y[0], y[1], y[2], y[3], y[4]
        f = 1e-6 * (np.exp((U2 - U3) / 0.026) - 1)
        return np.array([(Ue(t) - U1) / R0,
                         Ub / R2 - U2 * (1 / R1 + 1 / R2) - 0.01 * f,
                         f - U3 / R3,
                         Ub / R4 - U4 / R4 - 0.99 * f,
                         -U5 / R5])

    def jac(t, y):
        U2, U3 = y[1], y[2]
        df = 1e-6 * np.exp((U2 - U3) / 0.026) / 0.026
        return np.array([[-1 / R0, 0., 0., 0., 0.],
                         [0., -(1 / R1 + 1 / R2) - 0.01 * df, 0.01 * df, 0., 0.],
                         [0., df, -df - 1 / R3, 0., 0.],
                         [0., -0.99 * df, 0.99 * df, -1 / R4, 0.],
                         [0., 0., 0., 0., -1 / R5]])
    return fun, jac, mass, None


# --------------------------------------------------------------------------- n-pendulum
def npendulum_constants(n):
    """Masses and rod lengths, recursive_pendulum.py:63-73."""
    m_node = 1. / (n - 1)
    L_rod = 2. / n
    r0 = np.array([L_rod for _ in range(n)])
    m = np.array([m_node for _ in range(n)])
    m[-1] = 1.
    return m, r0, r0 ** 2


def npendulum_ic(n, initial_angle=np.pi / 4):
    """recursive_pendulum.py:168-177 (lambda0 = 0)."""
    _, r0, _ = npendulum_constants(n)
    theta_0 = np.array([initial_angle for _ in range(n)])
    x0 = np.cumsum(r0 * np.sin(theta_0))
    y0 = np.cumsum(-r0 * np.cos(theta_0))
    return np.vstack((x0, y0, np.zeros(n), np.zeros(n), np.zeros(n))).reshape((-1,), order='F')


def npendulum_tf():
    return 2 * (2 * np.pi * np.sqrt(2. / GRAVITY))


def npendulum(n):
    g = GRAVITY
    m, r0, r0s = npendulum_constants(n)

    def fun(t, X):
        x, y, vx, vy, lbda = X[0::5], X[1::5], X[2::5], X[3::5], X[4::5]
        dXdt = np.empty((5, x.size), order='C', dtype=X.dtype)
        dt_x, dt_y, dt_vx, dt_vy, constraints = dXdt
        dt_x[:] = vx
        dt_y[:] = vy
        dx = np.empty_like(x)
        dy = np.empty_like(x)
        dx[0] = x[0]
        dx[1:] = np.diff(x)
        dy[0] = y[0]
        dy[1:] = np.diff(y)
        rs = dx ** 2 + dy ** 2
        dt_vx[:-1] = (1 / m[:-1]) * (+lbda[:-1] * dx[:-1] / rs[:-1] - lbda[1:] * dx[1:] / rs[1:])
        dt_vy[:-1] = (1 / m[:-1]) * (+lbda[:-1] * dy[:-1] / rs[:-1] - lbda[1:] * dy[1:] / rs[1:]
                                     - m[:-1] * g)
        dt_vx[-1] = (1 / m[-1]) * (+lbda[-1] * dx[-1] / rs[-1])
        dt_vy[-1] = (1 / m[-1]) * (+lbda[-1] * dy[-1] / rs[-1] - m[-1] * g)
        constraints[:] = rs - r0s
        return dXdt.reshape((-1,), order='F')

    def jac(t, X):
        """Analytic Jacobian (dense [5n,5n]); not in the reference (FD there)."""
        x, y, lbda = X[0::5], X[1::5], X[4::5]
        N = 5 * n
        J = np.zeros((N, N))
        dx = np.empty_like(x)
        dy = np.empty_like(x)
        dx[0] = x[0]
        dx[1:] = np.diff(x)
        dy[0] = y[0]
        dy[1:] = np.diff(y)
        rs = dx ** 2 + dy ** 2
        # a_i = lbda_i*dx_i/rs_i, b_i = lbda_i*dy_i/rs_i and their partials wrt (dx_i, dy_i, lbda_i)
        a_dx = lbda * (rs - 2 * dx * dx) / rs ** 2
        a_dy = lbda * (-2 * dx * dy) / rs ** 2
        b_dx = a_dy
        b_dy = lbda * (rs - 2 * dy * dy) / rs ** 2
        a_l = dx / rs
        b_l = dy / rs
        for i in range(n):
            r = 5 * i
            J[r + 0, r + 2] = 1.
            J[r + 1, r + 3] = 1.
            im = 1 / m[i]
            # +a_i term: depends on x_i, x_{i-1}, y_i, y_{i-1}, lbda_i
            J[r + 2, r + 0] += im * a_dx[i]
            J[r + 2, r + 1] += im * a_dy[i]
            J[r + 2, r + 4] += im * a_l[i]
            J[r + 3, r + 0] += im * b_dx[i]
            J[r + 3, r + 1] += im * b_dy[i]
            J[r + 3, r + 4] += im * b_l[i]
            if i > 0:
                J[r + 2, r - 5 + 0] -= im * a_dx[i]
                J[r + 2, r - 5 + 1] -= im * a_dy[i]
                J[r + 3, r - 5 + 0] -= im * b_dx[i]
                J[r + 3, r - 5 + 1] -= im * b_dy[i]
            if i < n - 1:
                # -a_{i+1}: dx_{i+1} = x_{i+1} - x_i
                J[r + 2, r + 5 + 0] -= im * a_dx[i + 1]
                J[r + 2, r + 5 + 1] -= im * a_dy[i + 1]
                J[r + 2, r + 5 + 4] -= im * a_l[i + 1]
                J[r + 2, r + 0] += im * a_dx[i + 1]
                J[r + 2, r + 1] += im * a_dy[i + 1]
                J[r + 3, r + 5 + 0] -= im * b_dx[i + 1]
                J[r + 3, r + 5 + 1] -= im * b_dy[i + 1]
                J[r + 3, r + 5 + 4] -= im * b_l[i + 1]
                J[r + 3, r + 0] += im * b_dx[i + 1]
                J[r + 3, r + 1] += im * b_dy[i + 1]
            # constraint rs_i - r0s_i
            J[r + 4, r + 0] = 2 * dx[i]
            J[r + 4, r + 1] = 2 * dy[i]
            if i > 0:
                J[r + 4, r - 5 + 0] = -2 * dx[i]
                J[r + 4, r - 5 + 1] = -2 * dy[i]
        return J

    diag_mass = np.ones(5 * n)
    var_index = np.zeros(5 * n)
    for i in range(n):
        diag_mass[4 + i * 5] = 0
        var_index[2 + i * 5] = 2
        var_index[3 + i * 5] = 2
        var_index[4 + i * 5] = 3
    return fun, jac, np.diag(diag_mass), var_index


# --------------------------------------------------------------------------- registry
def build(problem, params=(), n=None):
    """(fun, jac, mass, var_index) of one system of `problem` with its parameter vector."""
    params = tuple(float(x) for x in params)
    if problem.startswith("pendulum"):
        return pendulum(int(problem[-1]), *params)
    if problem == "robertson":
        return robertson(*params)
    if problem == "robertson_ode":
        return robertson_ode(*params)
    if problem == "transistor":
        return transistor(*params)
    if problem == "npendulum":
        return npendulum(n // 5)
    if problem == "jay":
        return jay(*params)
    if problem == "polydae2":
        return polydae2(*params)
    if problem == "linear7":
        return linear7(*params)
    if problem == "reactor":
        return reactor(*params)
    if problem == "rod":
        return rod(*params, n=n)
    raise ValueError(problem)
