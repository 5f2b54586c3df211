"""Host-side registry of the problems compiled into libbrdae.so as device functors.

Each entry mirrors what the reference's `generateSystem` helpers hand to `solve_ivp`
(`t_span`, mass matrix, `var_index`, initial state) for the batched case, where the right-hand
side and Jacobian are not Python closures but CUDA functors selected by name:

  pendulum3/2/1 .... scipyDAE/tests/pendulum.py:101-236   params (m, g, r0)
  robertson ........ scipyDAE/tests/robertson.py:66-92     params (k1, k2, k3)
  transistor ....... scipyDAE/tests/transistor_amplifier.py:33-70
                     params (C1,C2,C3,R0,R1,R2,R3,R4,R5,Ub,Ua,w)
  npendulum ........ scipyDAE/tests/recursive_pendulum.py:34-183  (n nodes, 5n unknowns)
Only numpy is used here; nothing in this module integrates anything.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Problem:
    name: str
    problem_id: int
    n: int                       # state size; 0 = set per call (npendulum: 5 * nodes)
    param_names: tuple
    param_defaults: tuple
    var_index: tuple             # the reference's var_index for this formulation
    t_span: tuple                # the reference script's t_span
    per_system_mass: bool = False  # mass matrix built from each system's parameters
    doc: str = ""
    # user problems (dae_scipy_b200/plugin.py): own diag(0/1) mass matrix, whether an analytic Jacobian functor
    # exists, and the plugin library / generated functor header that hold them
    mass_diag: tuple | None = None
    has_jac: bool = True
    lib_path: str | None = None
    header_path: str | None = None
    n_event_functions: int = 0      # (nonlinear) event functions compiled into the functor: FunctorEvent(k), k < this
    band: tuple | None = None       # user band problem (more than 8 unknowns, CTA-per-system kernel): (kl, ku) of df/dy

    @property
    def n_params(self):
        return len(self.param_names)

    def default_params(self, B):
        return np.tile(np.asarray(self.param_defaults, dtype=np.float64), (B, 1)).reshape(B, self.n_params)

    def mass(self, params=None, n=None):
        """The reference's mass matrix for this problem (one system)."""
        if self.mass_diag is not None:
            return np.diag(np.asarray(self.mass_diag, dtype=np.float64))
        if self.name.startswith("pendulum"):
            M = np.eye(5)
            M[-1, -1] = 0
            return M
        if self.name == "robertson":
            return np.array([[1., 0., 0.], [0., 1., 0.], [0., 0., 0.]])
        if self.name == "robertson_ode":
            return np.eye(3)
        if self.name == "jay":
            M = np.eye(5)
            M[-1, -1] = 0
            return M
        if self.name == "polydae2":
            return np.array([[1., 0.], [0., 0.]])
        if self.name == "linear7":
            return np.eye(7)
        if self.name == "transistor":
            p = np.asarray(self.param_defaults if params is None else params, dtype=np.float64)
            C1, C2, C3 = p[0], p[1], p[2]
            return -np.array([[-C1, C1, 0, 0, 0], [C1, -C1, 0, 0, 0], [0, 0, -C2, 0, 0],
                              [0, 0, 0, -C3, C3], [0, 0, 0, C3, -C3]])
        if self.name == "npendulum":
            d = np.ones(n)
            d[4::5] = 0
            return np.diag(d)
        raise KeyError(self.name)

    def var_index_array(self, n=None):
        if self.name == "npendulum":
            return np.tile(np.array([0, 0, 2, 2, 3], dtype=np.int32), n // 5)
        return np.asarray(self.var_index, dtype=np.int32)


_PEND = dict(param_names=("m", "g", "r0"), param_defaults=(1.0, 9.81, 1.0), t_span=(0.0, 10.0))
PROBLEMS = {
    "pendulum3": Problem("pendulum3", 0, 5, var_index=(0, 0, 0, 0, 3), **_PEND),
    "pendulum2": Problem("pendulum2", 1, 5, var_index=(0, 0, 0, 0, 2), **_PEND),
    "pendulum1": Problem("pendulum1", 2, 5, var_index=(0, 0, 0, 0, 1), **_PEND),
    "robertson": Problem("robertson", 3, 3, ("k1", "k2", "k3"), (0.04, 1e4, 3e7), (0, 0, 1), (0.0, 5e6)),
    "transistor": Problem("transistor", 4, 5,
                          ("C1", "C2", "C3", "R0", "R1", "R2", "R3", "R4", "R5", "Ub", "Ua", "w"),
                          (1e-6, 2e-6, 3e-6, 1e3, 9e3, 9e3, 9e3, 9e3, 9e3, 6.0, 0.4, 200 * np.pi),
                          (0, 0, 0, 0, 0), (0.0, 1.0), per_system_mass=True),
    "npendulum": Problem("npendulum", 5, 0, (), (), (), (0.0, 2 * (2 * np.pi * np.sqrt(2.0 / 9.81)))),
    # known-answer problems of the reference's own checks
    "robertson_ode": Problem("robertson_ode", 6, 3, ("k1", "k2", "k3"), (0.04, 1e4, 3e7), (0, 0, 0), (0.0, 5e6)),
    "jay": Problem("jay", 7, 5, ("nonlinear",), (1.0,), (0, 0, 0, 0, 3), (0.0, 0.1)),
    "polydae2": Problem("polydae2", 8, 2, ("a4", "a3", "a2", "a1", "a0"), (0.0, 0.0, 0.0, 1.0, 0.0), (0, 2), (0.0, 2.0)),
    "linear7": Problem("linear7", 9, 7, tuple(f"c{k}" for k in range(7)), (1.0,) * 7, (0,) * 7, (0.0, 1.0)),
}


def get_problem(problem) -> Problem:
    if isinstance(problem, Problem):
        return problem
    try:
        return PROBLEMS[problem]
    except KeyError:
        raise ValueError(f"unknown problem {problem!r}; available: {sorted(PROBLEMS)}") from None


# ------------------------------------------------------------------ initial conditions
def pendulum_initial_state(theta_0, theta_dot0, m=1.0, g=9.81, r0=1.0):
    """Consistent initial condition of the pendulum, vectorised over systems
    (scipyDAE/tests/pendulum.py:225-231).  Returns [B, 5]."""
    theta_0 = np.atleast_1d(np.asarray(theta_0, dtype=np.float64))
    theta_dot0 = np.broadcast_to(np.asarray(theta_dot0, dtype=np.float64), theta_0.shape)
    x0 = r0 * np.sin(theta_0)
    y0 = -r0 * np.cos(theta_0)
    vx0 = r0 * theta_dot0 * np.cos(theta_0)
    vy0 = r0 * theta_dot0 * np.sin(theta_0)
    lbda_0 = (m * r0 * theta_dot0 ** 2 + m * g * np.cos(theta_0)) / r0
    return np.stack([x0, y0, vx0, vy0, lbda_0 * np.ones_like(x0)], axis=1)


def robertson_initial_state(B):
    """robertson.py:89."""
    y0 = np.zeros((B, 3))
    y0[:, 0] = 1.0
    return y0


def transistor_initial_state(params):
    """Consistent initial state, transistor_amplifier.py:45; params [B, 12]."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    R1, R2, Ub = params[:, 4], params[:, 5], params[:, 9]
    y0 = np.zeros((params.shape[0], 5))
    y0[:, 1] = Ub * R1 / (R1 + R2)
    y0[:, 2] = Ub * R1 / (R1 + R2)
    y0[:, 3] = Ub
    return y0


def npendulum_initial_state(n_nodes, initial_angle):
    """Chain hanging at `initial_angle` [B] from the vertical, at rest, lambda = 0
    (recursive_pendulum.py:168-177).  Returns [B, 5 n_nodes]."""
    ang = np.atleast_1d(np.asarray(initial_angle, dtype=np.float64))
    L_rod = 2.0 / n_nodes
    out = np.zeros((ang.size, 5 * n_nodes))
    for b, th in enumerate(ang):
        theta = np.array([th for _ in range(n_nodes)])
        r0 = np.array([L_rod for _ in range(n_nodes)])
        out[b, 0::5] = np.cumsum(r0 * np.sin(theta))
        out[b, 1::5] = np.cumsum(-r0 * np.cos(theta))
    return out
