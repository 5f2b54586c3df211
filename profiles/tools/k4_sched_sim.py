#!/usr/bin/env python
"""Trace-driven study of the K4 (lane-per-rope chain kernel) block schedule, CPU only.

The C oracle's per-attempt report of the first ropes of the bench ensemble gives, for every rope, the sequence of blocks it
walks through: S set-up, F factorisation, N Newton iteration, E error estimate/accept, J Jacobian update.  Warps of `rpw`
ropes are simulated under candidate schedules of the warp loop; the figure of merit is sweeps executed by the warp per rope
(a sweep = one pass over the nodes; the cost of every block in sweeps is below) and the lane efficiency
useful / (lanes x executed).

  python profiles/tools/k4_sched_sim.py [n_nodes] [ropes]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dae_scipy_b200 as br
from oracle import c_oracle as co

# cost of a block in node sweeps (memory passes over the rope): E = forward + backward of the error solve + accept sweep,
# S = scale + guess, F = one sweep (compute-heavy: two 3x3 inversions per node), N = forward + backward
COST = {"E": 3.0, "S": 2.0, "F": 1.5, "N": 2.0, "J": 1.0}


def traces(n_nodes, ropes):
    ens = br.ensembles.npendulum_ensemble(ropes, n_nodes=n_nodes)
    opts = dict(ens["options"])
    opts.pop("mass_matrix")
    r = co.solve_batch("npendulum", ens["t_span"], ens["y0"], np.zeros((ropes, 0)), t_eval=ens["t_eval"], report_cap=8192, **opts)
    out = []
    for i in range(ropes):
        k = r.reports["n"][i]
        code = r.reports["code"][i, :k]
        nit = r.reports["newton_iterations"][i, :k]
        # report codes (radauDAE.py:24-28): -10 factorisation, -11 jacobian update, 5 non-converged, 2 rejected, 0 accepted
        seq = []
        for c, it in zip(code, nit):
            if c == -10:
                seq.append("F")
            elif c == -11:
                seq.append("J")
            elif c == 5:
                seq.append(("Nfail", int(it)))
            else:
                seq.append(("E", int(it)))
        # expand into the block walk: S [F] N^k E | S [F] N^k (fail) | J ...
        walk = []
        pend_f = False
        for ev in seq:
            if ev == "F":
                pend_f = True
            elif ev == "J":
                # the failed Newton that led here is not in the report: the kernel has already spent its iterations
                walk.append("J")
            else:
                kind, it = ev
                walk.append("S")
                if pend_f:
                    walk.append("F")
                    pend_f = False
                walk.extend("N" * max(it, 1))
                if kind == "E":
                    walk.append("E")
        out.append("".join(walk))
    return out, r


def simulate(trs, rpw, policy, reps=3, frac=0.5):
    """One warp after the other (warps are independent); returns (executed sweeps per rope, lane efficiency)."""
    nw = len(trs) // rpw
    executed = useful = 0.0
    for w in range(nw):
        lanes = [trs[w * rpw + l] for l in range(rpw)]
        pos = [0] * rpw

        def want(b):
            return [l for l in range(rpw) if pos[l] < len(lanes[l]) and lanes[l][pos[l]] == b]

        def run(b, ls):
            nonlocal executed, useful
            if not ls:
                return
            executed += COST[b]
            useful += COST[b] * len(ls) / rpw
            for l in ls:
                pos[l] += 1

        while any(pos[l] < len(lanes[l]) for l in range(rpw)):
            alive = sum(pos[l] < len(lanes[l]) for l in range(rpw))
            run("E", want("E"))
            run("J", want("J"))
            run("S", want("S"))
            run("F", want("F"))
            if policy == "fixed":
                for _ in range(reps):
                    run("N", want("N"))
            elif policy == "vote":            # at least `reps` passes, then for as long as more than frac of the live ropes iterate
                k = 0
                while True:
                    ls = want("N")
                    if not ls or (k >= reps and len(ls) <= frac * alive):
                        break
                    run("N", ls)
                    k += 1
            elif policy == "all":
                while want("N"):
                    run("N", want("N"))
    return executed / nw, useful / executed


if __name__ == "__main__":
    n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    ropes = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    trs, r = traces(n_nodes, ropes)
    per = {b: np.mean([t.count(b) for t in trs]) for b in "SFNEJ"}
    ideal = sum(COST[b] * per[b] for b in per)
    print(f"n = {n_nodes}, {ropes} ropes: blocks per rope {per}; ideal sweeps per rope {ideal:.0f}")
    for rpw in (32, 16, 8):
        print(f" ropes per warp {rpw}")
        for policy, kw in (("fixed", dict(reps=2)), ("fixed", dict(reps=3)), ("fixed", dict(reps=4)), ("fixed", dict(reps=5)),
                           ("vote", dict(reps=1, frac=0.5)), ("vote", dict(reps=1, frac=0.25)), ("vote", dict(reps=1, frac=0.125)),
                           ("vote", dict(reps=3, frac=0.25)), ("all", {})):
            ex, eff = simulate(trs, rpw, policy, **kw)
            print(f"   {policy:6s} {str(kw):32s} sweeps executed per warp {ex:9.1f}  (x{ex / ideal:5.2f} of one rope alone)  lane efficiency {eff:.3f}")
