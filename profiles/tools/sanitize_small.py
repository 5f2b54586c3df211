#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (memcheck / synccheck / racecheck): the TMA-staged chain kernel at
several ropes-per-warp geometries with refills, the band kernel (analytic and finite-difference Jacobian), K1 (pendulum,
Robertson, transistor; finite differences; events), a user band problem.  Results are compared with the C oracle where it applies."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dae_scipy_b200 as br
from _util import assert_bit_identical, run_c_oracle, run_gpu

E = br.ensembles
for rpw in (2, 8, 32):
    os.environ["BRDAE_K4_RPW"] = str(rpw)
    os.environ["BRDAE_K4_MAX_SLOTS"] = str(2 * rpw)
    ens = E.npendulum_ensemble(3 * rpw + 1, n_nodes=5, t_end=0.3)
    assert_bit_identical(run_gpu(ens), run_c_oracle(ens), f"k4 rpw={rpw}")
    print("k4 staged, ropes per warp", rpw, "ok", flush=True)
del os.environ["BRDAE_K4_RPW"], os.environ["BRDAE_K4_MAX_SLOTS"]
ens = E.npendulum_ensemble(6, n_nodes=4, t_end=0.3)
assert_bit_identical(run_gpu(ens, jac=None), run_c_oracle(ens, jac=None), "k2 fd")
print("k2 finite differences ok", flush=True)
os.environ["BRDAE_NPENDULUM_KERNEL"] = "k2"
assert_bit_identical(run_gpu(ens), run_c_oracle(ens, chain_linear_algebra="band"), "k2 analytic")
del os.environ["BRDAE_NPENDULUM_KERNEL"]
print("k2 analytic ok", flush=True)
for name, ens, over in (("pendulum3", E.pendulum_ensemble(300, t_end=0.5), {}), ("robertson", E.robertson_ensemble(200, t_end=10.0), {}),
                        ("pendulum2 fd", E.pendulum_ensemble(100, t_end=0.5, index=2), dict(jac=None))):
    assert_bit_identical(run_gpu(ens, **over), run_c_oracle(ens, **over), name)
    print(name, "ok", flush=True)
out = run_gpu(E.transistor_ensemble(64, t_end=0.01))
assert (out["status"] == 0).all()
print("transistor ok", flush=True)
E.rod_problem()
out = run_gpu(E.rod_ensemble(10, t_end=0.3))
assert (out["status"] == 0).all()
print("user band problem ok", flush=True)
