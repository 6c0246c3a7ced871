#!/usr/bin/env python
"""List every loop (backward branch) of one kernel in a .so with its size and opcode histogram.

    python tools/sass_loops.py LIB.so SUBSTRING_OF_MANGLED_NAME [--min N] [--dump OUT.sass]
"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
mn = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 150
mx = int(sys.argv[sys.argv.index("--max") + 1]) if "--max" in sys.argv else 10**9
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
sel = [f for f in funcs if pat in f.split("\n")[0]]
if not sel:
    sys.exit("no function matches")
f = sel[0]
ins = []
for l in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f.split("\n")[0][-60:], len(ins), "instructions")
if "--dump" in sys.argv:
    open(sys.argv[sys.argv.index("--dump") + 1], "w").write("\n".join(f"{a:06x}  {t}" for a, t in ins))
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        if len(body) < mn or len(body) > mx:
            continue
        c = Counter()
        for _, x in body:
            parts = x.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            c[op.split(".")[0]] += 1
        print(f"loop {tgt:#x} -> {a:#x}: {len(body)} instrs", dict(c.most_common(50)))
