#!/usr/bin/env python
"""Where the time of one profiled kernel goes, per SASS instruction: executed counts and stall samples from
`ncu --set full --import-source on`, grouped by how often the instruction runs (hot path vs cold blocks).

    python tools/ncu_hotspots.py gpurun_out/x.ncu-rep [TOP]
"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ins = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        ins.append((r[col["Source"]].strip(), int(r[col["Instructions Executed"]]), int(r[col["Warp Stall Sampling (All Samples)"]]),
                    int(r[col["Warp Stall Sampling (Not-issued Samples)"]])))
    except ValueError:
        pass
tot_exec = sum(i[1] for i in ins)
tot_samp = sum(i[2] for i in ins)
mx = max(i[1] for i in ins)
hot = [i for i in ins if i[1] >= 0.5 * mx]
cold = [i for i in ins if i[1] < 0.5 * mx]
print(f"{len(ins)} SASS instructions, {tot_exec} executed, {tot_samp} stall samples")
print(f"hot  (executed >= half of the maximum): {len(hot):5d} instructions, {sum(i[1] for i in hot)/tot_exec:6.1%} of executed, {sum(i[2] for i in hot)/tot_samp:6.1%} of samples")
print(f"cold (the rest)                       : {len(cold):5d} instructions, {sum(i[1] for i in cold)/tot_exec:6.1%} of executed, {sum(i[2] for i in cold)/tot_samp:6.1%} of samples")
by_op = Counter()
by_op_n = Counter()
for s, e, a, n in ins:
    parts = s.split()
    op = (parts[1] if parts and parts[0].startswith("@") else parts[0]).split(".")[0] if parts else "?"
    by_op[op] += a
    by_op_n[op] += e
print("\nstall samples by opcode (share of samples | share of executed):")
for op, a in by_op.most_common(18):
    print(f"  {op:10s} {a/tot_samp:6.1%} | {by_op_n[op]/tot_exec:6.1%}")
print(f"\ntop {top} instructions by stall samples:")
for s, e, a, n in sorted(ins, key=lambda i: -i[2])[:top]:
    print(f"  {a/tot_samp:5.2%}  exec {e:9d}  {s[:90]}")
