#!/usr/bin/env python
"""Extract one kernel from `cuobjdump -sass` and summarise its loops.

    python tools/sass_fn.py LIB.so 'fused_x2_kernelILi8ELi3ELb0ELb1' [--dump START END]
"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
sel = [f for f in funcs if pat in f.split("\n")[0]]
if not sel:
    sys.exit("no function matches")
f = sel[0]
print("function:", f.split("\n")[0])
ins = []
for l in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions,", ins[-1][0] + 16, "bytes")
if "--dump" in sys.argv:
    i = sys.argv.index("--dump")
    a, b = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
    for addr, t in ins:
        if a <= addr <= b:
            print(f"{addr:06x}  {t}")
    sys.exit(0)
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a:
            loops.append((tgt, a))
loops = [l for l in loops if "--all" in sys.argv or (l[1] - l[0]) // 16 < 3000]
for tgt, a in loops[:12]:
    body = [x for x in ins if tgt <= x[0] <= a]
    c = Counter()
    for _, x in body:
        parts = x.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        c[op.split(".")[0]] += 1
    print(f"loop {tgt:#x} -> {a:#x}: {len(body)} instrs", dict(c.most_common(40)))
