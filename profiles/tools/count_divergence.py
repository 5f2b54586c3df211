#!/usr/bin/env python
"""Where do the step counts of the restatement first leave the reference's?  (VERDICT r1 weak #2, next-round task 7)

For the first systems of a BASELINE ensemble this script records the per-attempt report of the reference itself
(`RadauDAE(bReport=True)`, radauDAE.py:424-443) and of the C restatement (oracle/radau_ref.c `report`, the same record; the
CUDA kernels are bit-identical to it), finds the systems whose (accepted, rejected, failed) counts differ and prints the first
record at which the two reports disagree, with the records around it.

    python profiles/tools/count_divergence.py robertson 128
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dae_scipy_b200 as br  # noqa: E402
from oracle import c_oracle, problems_numpy as prob, reference_runner  # noqa: E402

CODES = {-10: "FACTORISATION", -11: "JACOBIAN_UPDATE", 5: "NON_CONV", 0: "ACCEPTED", 2: "REJECTED"}


def reference_report(ens, i):
    RadauDAE = reference_runner.radau_class()
    n = ens["y0"].shape[1]
    fun, jac, mass, _ = prob.build(ens["problem"], ens["params"][i], n)
    opts = dict(ens["options"])
    opts.pop("mass_matrix")
    vi = opts.pop("var_index")
    with contextlib.redirect_stdout(io.StringIO()):
        s = RadauDAE(fun, ens["t_span"][0], ens["y0"][i].copy(), ens["t_span"][1], jac=jac, mass_matrix=mass,
                     var_index=None if vi is None else np.asarray(vi, float), bReport=True, **opts)
        while s.status == "running":
            s.step()
    r = s.reports
    return dict(code=np.array(r["code"]), newt=np.array(r["newton_iterations"]), nbad=np.array(r["bad_iterations"]),
                t=np.array(r["t"], float), dt=np.array(r["dt"], float), err1=np.array(r["err1"], float), err2=np.array(r["err2"], float),
                counts=(s.naccpt, s.nrejct, s.nfailed))


def main():
    gen = sys.argv[1] if len(sys.argv) > 1 else "robertson"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    ens = br.ensembles.ENSEMBLES[gen](B)
    o = dict(ens["options"])
    o.pop("mass_matrix")
    c = c_oracle.solve_batch(ens["problem"], ens["t_span"], ens["y0"], ens["params"], t_eval=ens["t_eval"], report_cap=4000, **o)
    differ, first_kind = [], {}
    for i in range(B):
        ref = reference_report(ens, i)
        k = int(c.reports["n"][i])
        mine = dict(code=c.reports["code"][i, :k], newt=c.reports["newton_iterations"][i, :k], nbad=c.reports["bad_iterations"][i, :k],
                    t=c.reports["t"][i, :k], dt=c.reports["dt"][i, :k], err1=c.reports["err1"][i, :k], err2=c.reports["err2"][i, :k])
        same_counts = tuple(int(x) for x in c.counters[i, 4:7]) == tuple(ref["counts"])
        m = min(k, ref["code"].size)
        neq = np.nonzero((mine["code"][:m] != ref["code"][:m]) | (mine["newt"][:m] != ref["newt"][:m]))[0]
        if same_counts and neq.size == 0:
            continue
        j = int(neq[0]) if neq.size else m
        differ.append((i, j, ref, mine, same_counts))
    print(f"# {gen}: {B} systems, {sum(1 for d in differ if not d[4])} with different (accepted, rejected, failed) counts, "
          f"{len(differ)} whose per-attempt reports differ somewhere")
    for i, j, ref, mine, same in differ[:6]:
        print(f"\n## system {i}: counts reference {ref['counts']} vs restatement, same counts: {same}; first differing record {j} of {ref['code'].size}")
        print("| record | side | code | t | h | Newton its | bad its | err1 | err2 |")
        print("|---|---|---|---|---|---|---|---|---|")
        for q in range(max(0, j - 2), min(j + 2, ref["code"].size, mine["code"].size)):
            for side, r in (("reference", ref), ("restatement", mine)):
                print(f"| {q} | {side} | {CODES.get(int(r['code'][q]), r['code'][q])} | {r['t'][q]:.9g} | {r['dt'][q]:.9g} | {r['newt'][q]} | "
                      f"{r['nbad'][q]} | {r['err1'][q]:.6g} | {r['err2'][q]:.6g} |")
        if j < min(ref["code"].size, mine["code"].size):
            a, b = ref, mine
            kind = "Newton iteration count" if a["code"][j] == b["code"][j] else "accept / reject / failure decision"
            first_kind[kind] = first_kind.get(kind, 0) + 1
    print("\nfirst disagreement by kind:", first_kind)


if __name__ == "__main__":
    main()
