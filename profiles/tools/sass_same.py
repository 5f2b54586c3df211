#!/usr/bin/env python
"""Are the kernels of two objects the same machine code?  usage: sass_same.py old.o new.o [name-filter]

Compares `cuobjdump -sass` per kernel after dropping addresses, encodings and constant-bank offsets (a field added to
SolveArgs moves the offsets of the fields behind it and nothing else).  Used at the end of round 1 to show that the
event / harness-counter commits made after the last GPU run left every non-event K1 kernel instruction-identical to the
build the GPU suite and the bench had run (the only other differing lines: one `UMOV` of a SolveArgs array base per
finite-difference kernel)."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            s = re.sub(r"/\*[0-9a-fx ]+\*/", "", line).strip()
            funcs[cur].append(re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0x0][X]", s))
    return funcs


old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
flt = sys.argv[3] if len(sys.argv) > 3 else ""
for name in sorted(new):
    if flt not in name:
        continue
    if name not in old:
        print("new kernel      ", name[-70:])
        continue
    a, b = old[name], new[name]
    nd = sum(x != y for x, y in zip(a, b)) + abs(len(a) - len(b))
    print(f"{'identical' if nd == 0 else f'{nd} lines differ':16s} {len(b):6d} instructions  {name[-70:]}")
