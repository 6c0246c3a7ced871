#!/usr/bin/env python
"""Executed-instruction mix by opcode from the SASS source page of a .ncu-rep.

    python tools/ncu_opmix.py x.ncu-rep [px_per_launch]
"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
hdr = rows[hi]
isrc = hdr.index("Source")
iex = next(i for i, c in enumerate(hdr) if c == "# Instructions Executed" or c == "Instructions Executed")
c = Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= max(isrc, iex):
        continue
    try:
        n = int(r[iex].replace(",", ""))
    except ValueError:
        continue
    parts = r[isrc].split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    key = op.split(".")[0] if "--full" not in sys.argv else op
    if op.startswith("IMAD.MOV"):
        key = "IMAD.MOV"
    c[key] += n
    tot += n
print("total warp-instructions executed:", tot)
for k, v in c.most_common(45):
    extra = f"  {32 * v / px:6.2f} /px" if px else ""
    print(f"  {k:14s} {v:12d} {100 * v / tot:5.1f}%{extra}")
