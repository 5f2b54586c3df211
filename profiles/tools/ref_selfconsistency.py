#!/usr/bin/env python
"""How reproducible are the REFERENCE's own step counts under rounding-level noise?  

The reference's LU factorisations and substitutions run in OpenBLAS, whose kernels (and therefore the last bits of the
results) depend on the CPU family it dispatches to.  This script integrates the first systems of the BASELINE ensembles
with the UNMODIFIED reference (oracle/reference_runner.py: scipyDAE.radauDAE.RadauDAE from oracle/_ref or
/root/reference) once per OPENBLAS_CORETYPE, in fresh processes, and reports for every kernel family against the default
one: the fraction of systems with identical (accepted, rejected, failed) step counts and the largest difference of the
trajectories in units of atol + rtol |y|.  That band is the ceiling any independent implementation can be held to on
the count criterion of north_star; the C restatement and the CUDA kernels are measured against the same default run.

    python profiles/tools/ref_selfconsistency.py [--systems 512] [--out gpurun_out/ref_selfconsistency.json]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    "C2 pendulum3 t_end=2": ("pendulum", dict(t_end=2.0), 4),
    "C3 robertson": ("robertson", dict(), 2),
    "C4 transistor t_end=0.2": ("transistor", dict(t_end=0.2), 5),
}
CORETYPES = ["", "Haswell", "SkylakeX", "Zen", "Sandybridge", "Nehalem"]


def child(key, B, path):
    import dae_scipy_b200 as br
    from oracle import reference_runner
    gen, kw, _ = CONFIGS[key]
    ens = br.ensembles.ENSEMBLES[gen](B, **kw)
    out = reference_runner.run_reference_parallel(ens)
    try:
        from threadpoolctl import threadpool_info
        arch = sorted({i.get("architecture", "?") for i in threadpool_info() if i.get("internal_api") == "openblas"})
    except Exception:
        arch = ["?"]
    np.savez(path, arch=np.array(json.dumps(arch)), **out)


def stats(a, b, rtol, atol):
    same = float((a["counters"][:, 4:7] == b["counters"][:, 4:7]).all(axis=1).mean())
    both = ~(np.isnan(a["y"]) | np.isnan(b["y"]))
    err = np.where(both, np.abs(a["y"] - b["y"]) / (atol + rtol * np.abs(b["y"])), 0.0)
    per_sys = err.max(axis=(1, 2))
    return dict(identical_counts=same, max_scaled_error=float(per_sys.max()), p99_scaled_error=float(np.quantile(per_sys, 0.99)),
                bitwise_equal=bool(np.array_equal(a["y"], b["y"], equal_nan=True)), status_equal=bool(np.array_equal(a["status"], b["status"])))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--systems", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_selfconsistency.json"))
    ap.add_argument("--child", nargs=3)
    ap.add_argument("--with-c-oracle", action="store_true", help="also compare the C restatement with the default reference run")
    args = ap.parse_args()
    if args.child:
        child(args.child[0], int(args.child[1]), args.child[2])
        return
    import dae_scipy_b200 as br
    report = {"cores": os.cpu_count(), "systems": args.systems, "configs": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for key, (gen, kw, n_diff) in CONFIGS.items():
            B = args.systems if gen != "transistor" else max(64, args.systems // 2)
            runs = {}
            for ct in CORETYPES:
                env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
                env.pop("OPENBLAS_CORETYPE", None)
                if ct:
                    env["OPENBLAS_CORETYPE"] = ct
                path = os.path.join(tmp, f"{gen}_{ct or 'default'}.npz")
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", key, str(B), path], env=env, capture_output=True, text=True)
                if r.returncode != 0 or not os.path.exists(path):
                    runs[ct or "default"] = {"error": r.stderr[-400:]}
                    continue
                z = np.load(path)
                runs[ct or "default"] = {k: z[k] for k in z.files}
            ens = br.ensembles.ENSEMBLES[gen](B, **kw)
            rtol, atol = ens["options"]["rtol"], ens["options"]["atol"]
            base = runs["default"]
            entry = {"systems": B, "default_architecture": json.loads(str(base["arch"])), "vs_default": {}}
            for ct, r in runs.items():
                if ct == "default":
                    continue
                entry["vs_default"][ct] = r if "error" in r else dict(stats(r, base, rtol, atol), architecture=json.loads(str(r["arch"])))
            if args.with_c_oracle:
                from _util import run_c_oracle
                entry["c_restatement_vs_default"] = stats(run_c_oracle(ens), base, rtol, atol)
            report["configs"][key] = entry
            print(key, json.dumps(entry["vs_default"], indent=None), entry.get("c_restatement_vs_default"), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
