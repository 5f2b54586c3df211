#!/usr/bin/env python
"""Times the chain kernel K4 on the chain workloads under run-time settings (environment variables read by k4_chain.cu at every
call), for the library selected with BRDAE_LIB (compile-time variants: dae_scipy_b200.build.build_variant).

    BRDAE_LIB=... python profiles/tools/k4_variants.py TAG n:ropes[,n:ropes...] "VAR=val VAR=val;VAR=val ..."

Prints one line per (workload, setting): ropes/s and ms (CUDA events, 1 warm-up + 2 timed passes), and checks that the results
of every setting are bit-identical to the first one (scheduling must not change results)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dae_scipy_b200 as br

tag = sys.argv[1]
wls = [tuple(int(v) for v in w.split(":")) for w in sys.argv[2].split(",")]
settings = [dict(kv.split("=") for kv in s.split()) for s in sys.argv[3].split(";")] if len(sys.argv) > 3 else [{}]
dev = torch.device("cuda", 0)
out = []
for n_nodes, ropes in wls:
    ens = br.ensembles.npendulum_ensemble(ropes, n_nodes=n_nodes)
    y0 = torch.as_tensor(ens["y0"]).to(dev)
    ref = None
    for st in settings:
        for k in list(os.environ):
            if k.startswith("BRDAE_K4_") or k == "BRDAE_NEWTON_REPS":
                del os.environ[k]
        os.environ.update(st)
        run = lambda: br.solve_ivp_batched(ens["problem"], ens["t_span"], y0, None, t_eval=ens["t_eval"], device=dev, **ens["options"])
        r = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            r = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        y = r._y_eval.cpu().numpy()
        c = r.counters.cpu().numpy()
        if ref is None:
            ref = (y, c)
        same = bool(np.array_equal(y, ref[0], equal_nan=True) and np.array_equal(c, ref[1]))
        line = dict(tag=tag, n=n_nodes, ropes=ropes, setting=st, ms=round(ms, 1), ropes_per_s=round(ropes / ms * 1e3), same_as_first=same,
                    finished=int((r.status == 0).sum().item()))
        print(json.dumps(line), flush=True)
        out.append(line)
