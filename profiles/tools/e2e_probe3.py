#!/usr/bin/env python
"""Phase times of the host-buffer entry (BRDAE_ROWS_TIMING) under its switches: pipelined / serial copy-in, chunked SoA path."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dae_scipy_b200 as br
ens = br.ensembles.pendulum_ensemble(1 << 20)
n, B, T = 5, 1 << 20, 21
bufs = br.host_buffers(n, B, T, 3, pinned=True)
for name, env in (("pipelined-in", {}), ("serial-in", {"BRDAE_ROWS_SERIAL_IN": "1"}), ("chunked SoA", {"BRDAE_ROWS_CHUNKED": "1"})):
    for k in ("BRDAE_ROWS_SERIAL_IN", "BRDAE_ROWS_CHUNKED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["BRDAE_ROWS_TIMING"] = "1"
    for _ in range(2):
        br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], buffers=bufs, **ens["options"])
    t0 = time.perf_counter()
    for _ in range(4):
        out = br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], buffers=bufs, **ens["options"])
    dt = (time.perf_counter() - t0) / 4
    print(f"== {name}: {dt*1e3:.2f} ms per call, {B/dt/1e6:.2f} M trajectories/s, finished {(out['status']==0).all()}", flush=True)
