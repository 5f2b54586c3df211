#!/usr/bin/env python
"""Trace-driven study of the K1 block-dispatch policy (CPU only).

Event traces ('F' factorisation, 'N' Newton iteration, 'E' error estimate/accept) of the first
systems of the pendulum ensemble come from the C oracle; a group of G lanes with refill is then
simulated under several dispatch policies and the lane efficiency (useful lane-cost / executed
lane-cost) is printed.  Block costs are the measured SASS instruction counts of the kernel."""
import ctypes as C
import sys
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dae_scipy_b200 as br
from oracle import c_oracle as co

COST = {"F": 1500.0, "N": 800.0, "E": 1300.0}


def traces(nsys, t_end=2.0):
    ens = br.ensembles.pendulum_ensemble(nsys, t_end=t_end)
    L = co.lib()
    L.orc_trace.restype = C.c_int
    opts = dict(ens["options"]); opts.pop("mass_matrix")
    vi = np.ascontiguousarray(opts.pop("var_index"), dtype=np.int32)
    rtol = opts.pop("rtol"); atol = np.full(5, opts.pop("atol"))
    o = co.make_options(**opts)
    out = []
    buf = C.create_string_buffer(1 << 16)
    for i in range(nsys):
        y0 = np.ascontiguousarray(ens["y0"][i]); p = np.ascontiguousarray(ens["params"][i])
        n = L.orc_trace(C.c_int(0), C.c_int(5), C.c_int(3), co._dptr(y0), co._dptr(p), vi.ctypes.data_as(C.POINTER(C.c_int32)),
                        co._dptr(atol), C.c_double(rtol), C.c_double(0.0), C.c_double(t_end), C.byref(o), buf, C.c_int(1 << 16))
        out.append(buf.raw[:n].decode())
    return out


def simulate(trs, G, policy):
    queue = list(range(len(trs)))[::-1]
    lane_sys = [None] * G
    pos = [0] * G
    def refill(l):
        if queue:
            lane_sys[l] = queue.pop(); pos[l] = 0
        else:
            lane_sys[l] = None
    for l in range(G):
        refill(l)
    useful = executed = 0.0
    last = None
    while True:
        want = {"F": [], "N": [], "E": []}
        alive = 0
        for l in range(G):
            s = lane_sys[l]
            if s is None:
                continue
            alive += 1
            want[trs[s][pos[l]]].append(l)
        if alive == 0:
            break
        run = policy({k: len(v) for k, v in want.items()}, alive, last)
        last = run
        executed += COST[run] * G
        useful += COST[run] * len(want[run])
        for l in want[run]:
            pos[l] += 1
            if pos[l] >= len(trs[lane_sys[l]]):
                refill(l)
    return useful / executed


def majority(n, alive, last):
    if n["N"] > 0 and n["N"] >= n["E"] and n["N"] >= n["F"]:
        return "N"
    if n["E"] > 0 and n["E"] >= n["F"]:
        return "E"
    return "F"


def weighted(n, alive, last):          # maximise useful cost per trip
    return max("NEF", key=lambda k: (n[k] * COST[k], k == "N"))


def sticky(th):
    def pol(n, alive, last):           # keep running N while at least th of the lanes want it, else the fullest other pool
        if n["N"] >= th * alive:
            return "N"
        if n["E"] >= n["F"] and n["E"] > 0:
            return "E"
        if n["F"] > 0:
            return "F"
        return "N"
    return pol


def drain_order(n, alive, last):       # E whenever anyone converged would starve; cycle F -> N* -> E by fullest, but prefer finishing a phase
    if last and n[last] >= 0.5 * alive:
        return last
    return majority(n, alive, last)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "cycle"):
    nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    trs = traces(nsys)
    ev = "".join(trs)
    print("systems", nsys, "events/system", len(ev) / nsys, {k: ev.count(k) / nsys for k in "FNE"})
    for G in (32, 128, 256):
        row = [f"G={G:3d}"]
        for name, pol in (("majority", majority), ("weighted", weighted), ("sticky.25", sticky(0.25)), ("sticky.35", sticky(0.35)),
                          ("sticky.5", sticky(0.5)), ("drain", drain_order)):
            row.append(f"{name} {simulate(trs, G, pol):.3f}")
        print("  ".join(row))


# ---------------------------------------------------------------------------------------------------------------
# The shipped policy (fixed cycle E -> F -> N x reps, one CTA of W warps) and an upper bound for migrating systems
# between the warps of a CTA (same lane index, so the lane-strided shared-memory columns stay conflict-free):
#     python profiles/tools/sched_sim.py cycle [nsys]
def simulate_cycle(trs, W=4, reps=4, migrate=None, migrate_cost=0.0):
    """Lane efficiency of the fixed cycle for one CTA of W warps.  `migrate`: None, or a key function ranking the W systems
    that share a lane index (they are re-dealt to the warps in that order before every cycle); `migrate_cost` is charged per
    warp and cycle.  A block is executed by a warp when any of its lanes waits for it."""
    queue = list(range(len(trs)))[::-1]
    sysid = [[None] * 32 for _ in range(W)]
    pos = [[0] * 32 for _ in range(W)]

    def refill(w, l):
        if queue:
            sysid[w][l] = queue.pop(); pos[w][l] = 0
        else:
            sysid[w][l] = None
    for w in range(W):
        for l in range(32):
            refill(w, l)
    useful = executed = 0.0

    def nxt(w, l):
        s = sysid[w][l]
        return None if s is None else trs[s][pos[w][l]]

    def advance(w, l):
        pos[w][l] += 1
        if pos[w][l] >= len(trs[sysid[w][l]]):
            refill(w, l)

    def block(kind):
        nonlocal useful, executed
        for w in range(W):
            lanes = [l for l in range(32) if nxt(w, l) == kind]
            if lanes:
                executed += COST[kind] * 32
                useful += COST[kind] * len(lanes)
                for l in lanes:
                    advance(w, l)
    while any(sysid[w][l] is not None for w in range(W) for l in range(32)):
        if migrate is not None:
            for l in range(32):
                col = sorted(((sysid[w][l], pos[w][l]) for w in range(W)), key=lambda sp: migrate(trs, sp[0], sp[1], reps))
                for w in range(W):
                    sysid[w][l], pos[w][l] = col[w]
            executed += migrate_cost * 32 * W
        block("E"); block("F")
        for _ in range(reps):
            block("N")
    return useful / executed


def key_newton_passes(trs, s, p, reps):
    """How many of the cycle's Newton passes the system will use (what a lane knows only approximately: an upper bound)."""
    if s is None:
        return -1
    t = trs[s]
    while p < len(t) and t[p] in "EF":
        p += 1
    n = 0
    while p < len(t) and t[p] == "N" and n < reps:
        p += 1; n += 1
    return -n


def key_state_only(trs, s, p, reps):
    """What a lane does know at the top of a cycle: whether it is in the middle of a Newton sequence or at its start."""
    if s is None:
        return 2
    return 0 if trs[s][p] == "N" else 1


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "cycle":
    nsys = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    trs = traces(nsys)
    for reps in (3, 4, 5):
        print(f"reps={reps}: fixed cycle {simulate_cycle(trs, 4, reps):.3f}   "
              f"migration by Newton passes to come (oracle, free) {simulate_cycle(trs, 4, reps, key_newton_passes):.3f}   "
              f"(cost of one Newton pass per cycle) {simulate_cycle(trs, 4, reps, key_newton_passes, COST['N']):.3f}   "
              f"migration by state only {simulate_cycle(trs, 4, reps, key_state_only):.3f}")
    sys.exit(0)
