#!/usr/bin/env python
"""Hot-path instruction count of the main loop of a fused kernel: walks the largest loop that is not itself
nested in another counted loop, and skips every region that a predicated forward branch issued right after a
VOTE jumps over (the cold paths).  Prints the opcode histogram per loop trip (3 image rows).

    python tools/sass_hot.py LIB.so SUBSTRING [LOOP_MIN LOOP_MAX]
"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "--range" else 600
hi = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[3] != "--range" else 1400
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [f for f in funcs if pat in f.split("\n")[0]][0]
ins = []
for l in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        n = (a - int(m.group(1), 16)) // 16 + 1
        if lo <= n <= hi:
            loops.append((n, int(m.group(1), 16), a))
if "--range" in sys.argv:
    k = sys.argv.index("--range")
    start, end = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
    n = (end - start) // 16 + 1
else:
    if not loops:
        sys.exit("no loop in range")
    n, start, end = sorted(loops)[0]
body = [x for x in ins if start <= x[0] <= end]
c = Counter()
i = 0
hot = 0
prev_vote = False
while i < len(body):
    a, t = body[i]
    parts = t.split()
    pred = parts[0].startswith("@")
    op = (parts[1] if pred else parts[0])
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
    c[op.split(".")[0]] += 1
    hot += 1
    if m and pred and "DIV" not in op and int(m.group(1), 16) > a and prev_vote_recent:
        tgt = int(m.group(1), 16)
        while i < len(body) and body[i][0] < tgt:
            i += 1
        prev_vote_recent = False
        continue
    if op.startswith("VOTE"):
        prev_vote_recent = True
        vote_at = i
    elif "prev_vote_recent" in dir() and prev_vote_recent and i - vote_at > 40:
        prev_vote_recent = False
    if "prev_vote_recent" not in dir():
        prev_vote_recent = False
    i += 1
print(f.split("\n")[0][-50:], f"loop {start:#x}-{end:#x}: {n} instrs, hot {hot} ({hot/3:.1f} per row, {hot/24:.2f} per pixel)")
print(dict(c.most_common(60)))
