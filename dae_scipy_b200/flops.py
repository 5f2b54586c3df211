"""Algorithmic FLOP and byte model of the hot path, evaluated from the
per-system counters the kernels return — never from a profiler.

Counting rule: +, -, *, /, sqrt and each libm call count 1; an FMA counts 2; complex multiply 6,
complex add 2.  Dense size n, z = nnz(M):
    F_LU   = 5 (2/3 n^3 - 1/2 n^2 - 1/6 n)      real + complex factorisation (complex = 4 x real)
    F_form = 5 n^2 + 3 z                          forming both iteration matrices
    F_newt = (2n^2 - n) + 4 (2n^2 - n) + 60 n + 6 z   one simplified-Newton iteration (without its 3 RHS)
    F_err  = 2 n^2 + 15 n + 2 z                   one error estimate
    FLOPs  = nfev F_f + njev F_J + nlu (F_form + F_LU) + nlusolve F_newt + nlusolve_err F_err
             + naccpt 15 n + n_out (6 n + 3)
For the banded CTA-per-system kernel (K2) 2/3 n^3 becomes 2 N kl (kl + ku) and 2 n^2 becomes
2 N (2 kl + ku) with kl = 9, ku = 7.  The chain kernel K4 does not run that algorithm: it eliminates the velocities and
factorises 3 x 3 node blocks (csrc/chain_k4.cuh), and its model counts what that algorithm needs per node
(`flop_constants("npendulum", n, chain_block=True)`):
    factorisation (both systems)  ~700   (X = G U 18, E -= L X 36, Gauss-Jordan inverse ~75, block set-up ~10: 140 real, x 4 complex)
    Newton iteration              ~610   (real solve 57, complex 4 x 57, residuals / stage combinations / norm / update 60 per unknown, 6 z)
    error estimate                ~140   (real solve 57, 15 per unknown, 2 z)
so the FP64 fraction of a chain run is quoted for the work the kernel actually has to do, not for the band LU it replaced.
"""
from __future__ import annotations


# per right-hand-side / Jacobian evaluation; n-pendulum figures are per node
F_RHS = {"pendulum3": 12, "pendulum2": 12, "pendulum1": 20, "robertson": 15, "transistor": 22, "npendulum": 17}
F_JAC = {"pendulum3": 4, "pendulum2": 4, "pendulum1": 12, "robertson": 8, "transistor": 12, "npendulum": 30}
MASS_NNZ = {"pendulum3": 4, "pendulum2": 4, "pendulum1": 4, "robertson": 2, "transistor": 9}


def flop_constants(problem: str, n: int, chain_block: bool = False):
    if problem == "npendulum" and chain_block:
        nodes = n // 5
        return dict(F_f=F_RHS[problem] * nodes, F_J=F_JAC[problem] * nodes, F_LU=690 * nodes, F_form=10 * nodes,
                    F_newt=(5 * 57 + 60 * 5 + 6 * 4) * nodes, F_err=(57 + 15 * 5 + 2 * 4) * nodes)
    if problem == "npendulum":
        nodes = n // 5
        kl, ku = 9, 7
        z = 4 * nodes
        lu_real = 2 * n * kl * (kl + ku)
        solve_real = 2 * n * (2 * kl + ku)
        f_lu = 5 * lu_real
        f_form = 5 * n * (kl + ku + 1) + 3 * z
        f_newt = 5 * solve_real + 60 * n + 6 * z
        f_err = solve_real + 15 * n + 2 * z
        return dict(F_f=F_RHS[problem] * nodes, F_J=F_JAC[problem] * nodes, F_LU=f_lu, F_form=f_form,
                    F_newt=f_newt, F_err=f_err)
    z = MASS_NNZ[problem]
    f_lu = 5 * (2 / 3 * n ** 3 - n ** 2 / 2 - n / 6)
    return dict(F_f=F_RHS[problem], F_J=F_JAC[problem], F_LU=f_lu, F_form=5 * n * n + 3 * z,
                F_newt=(2 * n * n - n) + 4 * (2 * n * n - n) + 60 * n + 6 * z,
                F_err=2 * n * n + 15 * n + 2 * z)


def algorithmic_flops(problem: str, n: int, counter_sums, n_out_total: int, chain_block: bool = False) -> float:
    """counter_sums: the 9 counters summed over the ensemble, in _abi.COUNTER_NAMES order."""
    c = flop_constants(problem, n, chain_block)
    nfev, njev, nlu, _nstep, naccpt, _nrej, _nfail, nsolve, nsolve_err = [float(x) for x in counter_sums]
    return (nfev * c["F_f"] + njev * c["F_J"] + nlu * (c["F_form"] + c["F_LU"]) + nsolve * c["F_newt"]
            + nsolve_err * c["F_err"] + naccpt * 15 * n + float(n_out_total) * (6 * n + 3))


def algorithmic_bytes(n: int, n_params: int, T: int, B: int) -> float:
    """HBM bytes a launch must move: y0 + params in; y_eval, t_final, y_final, status, n_eval,
    counters out."""
    per_sys = 8 * (n + n_params) + 8 * n * T + 8 * (n + 1) + 4 * (2 + 9)
    return float(per_sys) * B


def chain_state_bytes_per_node_sweep() -> dict:
    """Bytes of rope state the chain kernel K4 moves per node (and per rope) in each kind of sweep: the slot range the sweep
    stages in shared memory (TMA bulk copy of slots [lo, hi) of the node, csrc/chain_k4.cuh: Stager) plus the slots it
    stores.  8-byte slots; the figures are the `S.begin(R, lo, count, ...)` ranges of the sweeps."""
    newton = 8 * (54 + 69) + 8 * (15 + 15)          # forward [O_Y, O_Z) + backward [O_W, SPN); stores z, position rows | W
    factor = 8 * 7 + 8 * 27                         # rod quantities in, three inverse blocks out
    errest = 8 * (41 + 65) + 8 * (9 + 5)            # first estimate: forward [O_F, O_VCR) + backward [O_Y, O_RP + 2)
    setup = 8 * 30 + 8 * 20                         # [O_Q, O_W) in; W and the Newton scales out
    accept = 8 * 20 + 8 * 37                        # [O_Y, O_ND) in; y_old, y, Q, f and the rod quantities out
    return dict(newton=newton, factor=factor, error_estimate=errest, setup=setup, accept=accept)
