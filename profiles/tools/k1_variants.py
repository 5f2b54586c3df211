#!/usr/bin/env python
"""Times a K1 workload under run-time settings (BRDAE_NEWTON_REPS ...) for the library selected with BRDAE_LIB.

    python profiles/tools/k1_variants.py TAG pendulum|robertson|transistor[:B] "VAR=val;VAR=val ..." """
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dae_scipy_b200 as br

tag = sys.argv[1]
kind, _, b = sys.argv[2].partition(":")
B = int(b) if b else {"pendulum": 1 << 20, "robertson": 1 << 22, "transistor": 1 << 18}[kind]
settings = [dict(kv.split("=") for kv in s.split()) for s in sys.argv[3].split(";")] if len(sys.argv) > 3 else [{}]
dev = torch.device("cuda", 0)
ens = br.ensembles.ENSEMBLES[kind](B)
y0 = torch.as_tensor(ens["y0"]).to(dev)
par = torch.as_tensor(ens["params"]).to(dev)
ref = None
for st in settings:
    for k in list(os.environ):
        if k == "BRDAE_NEWTON_REPS":
            del os.environ[k]
    os.environ.update(st)
    run = lambda: br.solve_ivp_batched(ens["problem"], ens["t_span"], y0, par, t_eval=ens["t_eval"], device=dev, **ens["options"])
    for _ in range(2):
        r = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    c = r.counters.cpu().numpy()
    if ref is None:
        ref = c
    print(json.dumps(dict(tag=tag, workload=kind, B=B, setting=st, ms=round(ms, 2), per_s=round(B / ms * 1e3), same_counters=bool(np.array_equal(c, ref)))), flush=True)
