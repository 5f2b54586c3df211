import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dae_scipy_b200 as br
B = 1 << 20
ens = br.ensembles.pendulum_ensemble(B)
bufs = br.host_buffers(5, B, 21, 3, pinned=True)
def run():
    return br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], buffers=bufs, **ens["options"])
for _ in range(2): run()
os.environ["BRDAE_ROWS_TIMING"] = "1"
for mode in ("pipelined", "serial", "pipelined", "serial"):
    if mode == "serial": os.environ["BRDAE_ROWS_SERIAL_IN"] = "1"
    else: os.environ.pop("BRDAE_ROWS_SERIAL_IN", None)
    for _ in range(2):
        t0 = time.perf_counter(); run(); print(f"{mode}: python call {1e3*(time.perf_counter()-t0):.1f} ms", file=sys.stderr)
