#!/usr/bin/env python
"""Condensed view of one .ncu-rep (the numbers profiles/*.md quote).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [px_per_launch]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "sm__inst_issued.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:120])
    for k in want:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    stalls = [(k, float(d[k].replace(',', ''))) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d[k]]
    stalls.sort(key=lambda kv: -kv[1])
    print("  stalls (warps per issue-active cycle):", ", ".join(f"{k[34:-23]}={v:.2f}" for k, v in stalls[:9]))
    if px:
        inst = float(d["smsp__inst_executed.sum"].replace(',', ''))
        print(f"  warp-instructions per pixel: {inst / px:.3f}  (= {32 * inst / px:.1f} lane-instructions per pixel)")
        t = float(d["gpu__time_duration.sum"].replace(',', ''))
        u = units[hdr.index("gpu__time_duration.sum")]
        t_us = t * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1)
        print(f"  {px / t_us:.0f} Mpx/s under ncu")
