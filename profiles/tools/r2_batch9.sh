#!/bin/bash
# Round 2, GPU call 9: finite-difference Jacobian of the chain on the band kernel (bit-identical to the C oracle; reference fixtures),
# timing of the like-for-like chain run (jac=None), the default bench line with the D2H probe.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fd_ or npendulum" > gpurun_out/r2b9_pytest_fd.log 2>&1; echo "pytest fd rc=$?"; tail -4 gpurun_out/r2b9_pytest_fd.log
timeout -k 10 600 python - > gpurun_out/r2b9_fd_chain_timing.txt 2>&1 <<'PY'
import time, torch, numpy as np, sys
sys.path.insert(0, '.')
import dae_scipy_b200 as br
for n, B in ((20, 16384), (50, 4096)):
    ens = br.ensembles.npendulum_ensemble(B, n_nodes=n)
    y0 = torch.as_tensor(ens["y0"]).cuda()
    for jac in ("analytic", None):
        opts = dict(ens["options"], jac=jac) if jac is None else ens["options"]
        run = lambda: br.solve_ivp_batched(ens["problem"], ens["t_span"], y0, None, t_eval=ens["t_eval"], **opts)
        r = run(); torch.cuda.synchronize()
        t0 = time.perf_counter(); r = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"n={n} B={B} jac={jac}: {dt*1e3:.1f} ms, {B/dt:.0f} ropes/s, finished {int((r.status==0).sum())}, mean njev {r.counters[1].double().mean().item():.1f}", flush=True)
PY
cat gpurun_out/r2b9_fd_chain_timing.txt | tail -5
timeout -k 10 600 python bench.py > gpurun_out/r2b9_bench_default.json 2> gpurun_out/r2b9_bench_default.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2b9_bench_default.json').read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), 'e2e', d['e2e'], d['cpu_baseline'], d['cpu_baseline_native'])"
