#!/usr/bin/env python
"""Hot-path listing of a loop of a kernel: like sass_hot.py, but prints a finer opcode histogram (IMAD.MOV of a constant
/ of a register, IMAD.SHL, predicated moves ...) and can dump the hot instruction stream.

    python tools/sass_hot2.py LIB.so SUBSTRING START END [--dump FILE]
"""
import re
import subprocess
import sys
from collections import Counter

lib, pat, start, end = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [f for f in funcs if pat in f.split("\n")[0]][0]
ins = []
for l in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
body = [x for x in ins if start <= x[0] <= end]
c = Counter()
hot = []
i = 0
vote_at = -100
while i < len(body):
    a, t = body[i]
    parts = t.split()
    pred = parts[0].startswith("@")
    op = parts[1] if pred else parts[0]
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
    key = op.split(".")[0]
    if op.startswith("IMAD.MOV"):
        key = "IMAD.MOV(const)" if re.search(r"RZ, RZ, (0x|RZ|-?\d)", t) else "IMAD.MOV(reg)"
    elif op.startswith("IMAD.SHL"):
        key = "IMAD.SHL"
    elif op.startswith("IMAD.IADD"):
        key = "IMAD.IADD"
    elif key == "MOV":
        key = "MOV(const)" if re.search(r"MOV R\d+, (0x|RZ|UR|c\[)", t) else "MOV(reg)"
    if pred and key.startswith(("MOV", "IMAD.MOV")):
        key = "@" + key
    c[key] += 1
    hot.append((a, t))
    if m and pred and "DIV" not in op and int(m.group(1), 16) > a and i - vote_at <= 40:
        tgt = int(m.group(1), 16)
        while i < len(body) and body[i][0] < tgt:
            i += 1
        vote_at = -100
        continue
    if op.startswith("VOTE"):
        vote_at = i
    i += 1
n = len(hot)
print(f"loop {start:#x}-{end:#x}: {len(body)} instrs, hot {n} ({n/3:.1f} per row, {n/24:.2f} per pixel)")
print(dict(c.most_common(80)))
if "--dump" in sys.argv:
    open(sys.argv[sys.argv.index("--dump") + 1], "w").write("\n".join(f"{a:06x}  {t}" for a, t in hot))
