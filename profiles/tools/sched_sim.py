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


if __name__ == "__main__":
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
