"""Where the end-to-end time of solve_ivp_batched_host goes (tuning helper)."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dae_scipy_b200 as br
import torch
B = 1 << 20
ens = br.ensembles.pendulum_ensemble(B)
n, T = 5, 21
bufs = br.host_buffers(n, B, T, 3, pinned=True)
def run():
    return br.solve_ivp_batched_host(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], buffers=bufs, **ens["options"])
for chunk in (os.environ.get("CHUNKS", "1048576,524288,262144,131072,65536").split(",")):
    os.environ["BRDAE_ROWS_CHUNK"] = chunk
    for _ in range(2): run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): run()
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk {chunk}: {dt*1e3:.1f} ms per call, {B/dt/1e6:.2f} M traj/s")
# python-side share
t0 = time.perf_counter()
for _ in range(4):
    y0 = np.ascontiguousarray(ens["y0"], dtype=np.float64)
    hb = br.prepare(ens["problem"], ens["t_span"], y0.shape, t_eval=ens["t_eval"], params=ens["params"], **ens["options"])
    ok = np.isfinite(y0).all()
print(f"python-side validation: {(time.perf_counter()-t0)/4*1e3:.2f} ms per call")
# pinned inputs
py0 = br.api.PinnedArray(ens["y0"].shape, np.float64); py0.array[:] = ens["y0"]
pp = br.api.PinnedArray(ens["params"].shape, np.float64); pp.array[:] = ens["params"]
os.environ["BRDAE_ROWS_CHUNK"] = "131072"
def run2():
    return br.solve_ivp_batched_host(ens["problem"], ens["t_span"], py0.array, pp.array, t_eval=ens["t_eval"], buffers=bufs, **ens["options"])
for _ in range(2): run2()
t0 = time.perf_counter()
for _ in range(4): run2()
dt = (time.perf_counter() - t0) / 4
print(f"pinned inputs, chunk 131072: {dt*1e3:.1f} ms per call, {B/dt/1e6:.2f} M traj/s")
