#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <library.so> <kernel-name-substring> [top_n]
Needs `ncu`, `cuobjdump`, `nvdisasm` on PATH and a library built with -lineinfo."""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
amap = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    if kern not in dis:
        continue
    cur, want = None, False
    for line in dis.splitlines():
        if ".text." in line:
            want = kern in line
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if want:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                amap[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
base = None
for r in data:
    if len(r) != len(hdr):
        continue
    a = int(r[ix["Address"]], 16)
    base = a if base is None else base
    key = amap.get(a - base)
    for name, col in (("inst", "Instructions Executed"), ("thr", "Thread Instructions Executed"), ("samples", "# Samples")):
        v = int(r[ix[col]] or 0)
        agg[key][name] += v
        tot[name] += v
    for s in stalls:
        v = int(r[ix[s]] or 0)
        agg[key][s] += v
        tot[s] += v
    agg[key]["sass"] += 1
print(f"kernel {kern}: warp instructions {tot['inst']:,}  avg active threads {tot['thr'] / max(tot['inst'], 1):.2f}  samples {tot['samples']:,}")
print("stall mix (% of samples): " + ", ".join(f"{s[6:]} {100 * tot[s] / max(tot['samples'], 1):.1f}" for s in sorted(stalls, key=lambda s: -tot[s])[:8]))
print(f"{'file:line':28s} {'inst%':>6s} {'thr/inst':>8s} {'samp%':>6s} {'sass':>5s}  top stalls")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    ts = sorted(stalls, key=lambda s: -c[s])[:3]
    name = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{name:28s} {100 * c['inst'] / tot['inst']:6.2f} {c['thr'] / max(c['inst'], 1):8.1f} {100 * c['samples'] / tot['samples']:6.2f} {c['sass']:5d}  "
          + ", ".join(f"{s[6:]} {100 * c[s] / max(c['samples'], 1):.0f}%" for s in ts))
