"""Host-side logic that needs no GPU: argument validation (same errors as the reference path),
seeded ensembles, FLOP model, sharding and the torch.distributed gather/reduce (gloo, 2 ranks)."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

import dae_scipy_b200 as br
from dae_scipy_b200 import dist as brdist, flops
from _util import ROOT


def _prep(**over):
    ens = br.ensembles.pendulum_ensemble(4)
    opts = dict(ens["options"])
    te = over.pop("t_eval", ens["t_eval"])
    span = over.pop("t_span", ens["t_span"])
    shape = over.pop("shape", ens["y0"].shape)
    opts.update(over)
    return br.prepare(ens["problem"], span, shape, t_eval=te, **opts)


@pytest.mark.parametrize("kw,msg", [
    (dict(t_eval=np.zeros((2, 2))), "1-dimensional"),
    (dict(t_eval=[0.0, 3.0]), "not within `t_span`"),
    (dict(t_eval=[0.0, 1.0, 1.0]), "not properly sorted"),
    (dict(atol=[1e-6, 1e-6]), "`atol` has wrong shape"),
    (dict(atol=-1.0), "`atol` must be positive"),
    (dict(first_step=-1.0), "`first_step` must be positive"),
    (dict(first_step=5.0), "`first_step` exceeds bounds"),
    (dict(max_step=0.0), "`max_step` must be positive"),
    (dict(mass_matrix=np.eye(4)), "`mass_matrix` is expected to have shape"),
    (dict(mass_matrix=lambda t: 0), "callable"),
    (dict(jac=np.eye(3)), "`jac` is expected to have shape"),
    (dict(var_index=[0, 0, 0]), "var_index"),
    (dict(var_index=[0, 0, 0, 0, 4]), "between 0 and 3"),
    (dict(shape=(4, 6)), "unknowns"),
])
def test_validation_errors_mirror_the_reference(kw, msg):
    with pytest.raises(ValueError, match=msg):
        _prep(**kw)


def test_extraneous_options_warn_like_scipy():
    with pytest.warns(UserWarning, match="no effect for a chosen solver: `bogus`"):
        _prep(bogus=1)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        _prep(bPrint=True, bPrintProgress=False, bDebug=False, bReport=True, vectorized=False, jac_sparsity=None)


def test_option_defaults_follow_the_reference_constructor():
    hb = _prep()
    o = hb.options
    assert (o.max_newton_ite, o.max_bad_ite, o.min_factor, o.max_factor) == (6, -1, 0.2, 10.0)
    assert (o.safety_factor, o.factor_on_non_convergence, o.jac_recompute_factor) == (0.9, 0.5, 1e-3)
    assert (o.scale_residuals, o.scale_newton_norm, o.scale_error, o.zero_algebraic_error) == (1, 1, 1, 0)
    assert (o.always_2nd_estimate, o.predictive_controller, o.predictive_newton_stop, o.extrapolated_guess) == (1, 1, 1, 1)
    assert (o.all_newton_iterations, o.constant_dt, o.nmax_step) == (0, 0, 0)
    assert np.isinf(o.max_step) and o.newton_tol < 0
    assert hb.mass is None                      # 'problem' -> compile-time mass structure
    hb2 = _prep(mass_matrix=np.diag([1., 1, 1, 1, 0]))
    assert hb2.mass is None                     # the problem's own matrix given explicitly -> same kernel
    hb3 = _prep(mass_matrix=None)
    assert np.array_equal(hb3.mass, np.eye(5))  # reference default: identity
    assert _prep(var_index=None).var_index.tolist() == [0] * 5
    assert o.jac_finite_differences == 0
    assert _prep(jac=None).options.jac_finite_differences == 1      # reference default: finite differences
    e = br.ensembles.npendulum_ensemble(2, n_nodes=4)               # the chain too (band kernel, the reference's own default)
    hbc = br.prepare(e["problem"], e["t_span"], e["y0"].shape, t_eval=e["t_eval"], **dict(e["options"], jac=None))
    assert hbc.options.jac_finite_differences == 1


def test_ensembles_are_seeded_and_prefix_stable():
    a = br.ensembles.pendulum_ensemble(64)
    b = br.ensembles.pendulum_ensemble(64)
    assert np.array_equal(a["y0"], b["y0"])
    c = br.ensembles.pendulum_ensemble(1024)
    assert np.array_equal(c["y0"][:64], a["y0"])            # subsample of a larger run = same systems
    x, y = a["y0"][:, 0], a["y0"][:, 1]
    assert np.allclose(x * x + y * y, 1.0)                  # consistent with the constraint
    r = br.ensembles.robertson_ensemble(32)
    assert r["params"].shape == (32, 3) and (r["y0"] == [1, 0, 0]).all()
    t = br.ensembles.transistor_ensemble(8)
    assert t["params"].shape == (8, 12) and np.allclose(t["y0"][:, 3], 6.0)
    n = br.ensembles.npendulum_ensemble(3, n_nodes=6)
    assert n["y0"].shape == (3, 30) and len(n["options"]["var_index"]) == 30


def test_flop_model_matches_survey_figures():
    c = flops.flop_constants("pendulum3", 5)
    assert c["F_newt"] == (50 - 5) + 4 * (50 - 5) + 300 + 24 and c["F_err"] == 50 + 75 + 8
    assert abs(c["F_LU"] - 5 * (2 / 3 * 125 - 12.5 - 5 / 6)) < 1e-9
    # G1 counters: about 0.29 MFLOP per trajectory
    f = flops.algorithmic_flops("pendulum3", 5, [1228, 92, 91, 98, 98, 0, 0, 353, 168], 21)
    assert 0.2e6 < f < 0.4e6
    assert flops.algorithmic_bytes(5, 3, 21, 10) == 10 * (64 + 840 + 48 + 44)


def test_shard_bounds_cover_the_ensemble_exactly():
    for B in (0, 1, 7, 1 << 20):
        for w in (1, 2, 3, 8):
            cuts = [brdist.shard_bounds(B, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_and_reduce():
    """world_size = 2 on CPU: shard an ensemble, integrate each shard with the CPU checker standing
    in for the GPU kernel, gather results on rank 0 and all_reduce the statistics."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", WORLD_SIZE="2", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "rank0 ok" in outs[0]


def test_per_system_ode_result_view():
    """BatchedOdeResult.system(i): the OdeResult a user of the reference gets from solve_ivp, rebuilt from the batch arrays.
    The arrays come from the kernel source run on the CPU (kernel layout, CPU tensors), the expected values from the
    reference's own run of the same system (tests/golden/g1_pendulum3_single.npz)."""
    import torch
    from dae_scipy_b200.api import BatchedOdeResult
    from _util import load_golden, run_kernel_source_on_cpu
    ens, ref = load_golden("g1_pendulum3_single")
    out = run_kernel_source_on_cpu(ens, dense_cap=400)
    res = BatchedOdeResult(np.asarray(ens["t_eval"], dtype=float), torch.as_tensor(out["y"]).permute(2, 1, 0).contiguous(),
                           torch.as_tensor(out["n_eval"]), torch.as_tensor(out["status"]), torch.as_tensor(out["t_final"]),
                           torch.as_tensor(out["y_final"]).t().contiguous(), torch.as_tensor(out["counters"]).t().contiguous())
    res.sol = out["sol"]
    r = res.system(0)
    assert r.success and r.status == 0 and r.message.startswith("The solver successfully reached the end")
    assert r.t.shape == (21,) and r.y.shape == (5, 21)
    assert np.abs(r.y - ref["y"][0]).max() <= 10 * (1e-6 + 1e-6 * np.abs(ref["y"][0]).max())
    assert (r.nfev, r.njev, r.nlu) == tuple(int(v) for v in ref["counters"][0, :3])
    assert (r.naccpt, r.nrejct, r.nfailed) == tuple(int(v) for v in ref["counters"][0, 4:7])
    np.testing.assert_allclose(r.sol(r.t[5]), r.y[:, 5], rtol=0, atol=1e-13)         # the dense output goes through the t_eval values
    assert r.t_events is None and float(r.t_final) == 2.0
    # t_eval=None: the steps themselves, as the reference returns them
    res2 = BatchedOdeResult(np.empty(0), torch.empty((0, 5, 1), dtype=torch.float64), torch.zeros(1, dtype=torch.int32),
                            torch.as_tensor(out["status"]), torch.as_tensor(out["t_final"]),
                            torch.as_tensor(out["y_final"]).t().contiguous(), torch.as_tensor(out["counters"]).t().contiguous())
    res2.sol = out["sol"]
    r2 = res2.system(0)
    assert r2.t.shape == (r.naccpt + 1,) and r2.t[0] == 0.0 and r2.t[-1] == 2.0 and r2.y.shape == (5, r.naccpt + 1)
    np.testing.assert_array_equal(r2.y[:, -1], out["y_final"][0])


def test_var_index_of_the_problem():
    hb = br.prepare("pendulum3", (0.0, 1.0), (3, 5), var_index="problem")
    assert hb.var_index.tolist() == [0, 0, 0, 0, 3]
    assert br.prepare("npendulum", (0.0, 1.0), (2, 10), var_index="problem").var_index.tolist() == [0, 0, 2, 2, 3] * 2
    assert br.prepare("pendulum3", (0.0, 1.0), (3, 5)).var_index.tolist() == [0] * 5        # None: all differential, as in the reference
    with pytest.raises(ValueError):
        br.prepare("pendulum3", (0.0, 1.0), (3, 5), var_index="auto")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): one JSON line with the contract's
    keys, measured on a bounded sample with every host core."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "16"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "trajectories/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"] == "pendulum3_1M_tend2" and d["dtype"] == "f64" and d["vs_baseline"] is None
    from oracle import reference_runner          # the unmodified reference when it is importable, else its numpy port
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_runner.reference_location() else "port")
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    # the other ranks of a torchrun launch exit without work
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
