#!/usr/bin/env python
"""Plots of the results tables -- the counterpart of the reference's src/*/results/visualisation.py (CleanData,
CPUvsOpenCLEndtoEnd, KernelvsOpenCLTotal, SpeedFactorEndtoEnd, MAE: lines 17-89), extended with what this repo measures:
Mpixel/s, share of the HBM peak, and the 1 -> 8 GPU curves.

    python tools/plot_results.py --csv results_extended.csv [--bench BENCH_1.json BENCH_2.json ...] [--out profiles/plots]

* --csv    a table written by FileHandler::WriteExtendedResultsToCSV (tools/rip_headless.bin images ...); the
           reference's own 11-column files work too (the extra charts are skipped)
* --bench  JSON lines printed by bench.py at different --gpus (or the driver's BENCH_rNN.json / SCALE_rNN.json):
           whole-job Mpixel/s resident, end to end, end to end with NV12 input, against linear scaling

Charts are written as SVG by a small built-in writer (no matplotlib in this image); with matplotlib installed the same
series are also drawn as PNG.
"""
from __future__ import annotations

import argparse
import csv
import json
import math
import os

COLORS = ["#1f77b4", "#d62728", "#2ca02c", "#ff7f0e", "#9467bd", "#8c564b", "#17becf", "#7f7f7f"]


def _ticks(lo: float, hi: float, log: bool) -> list[float]:
    if log:
        a, b = math.floor(math.log10(lo)), math.ceil(math.log10(hi))
        return [10.0 ** e for e in range(a, b + 1)]
    if hi <= lo:
        hi = lo + 1.0
    step = 10 ** math.floor(math.log10((hi - lo) / 4))
    for m in (1, 2, 5, 10):
        if (hi - lo) / (step * m) <= 6:
            step *= m
            break
    t0 = math.floor(lo / step) * step
    out = []
    v = t0
    while v <= hi + 1e-9 * step:
        out.append(round(v, 10))
        v += step
    return out


def _fmt(v: float) -> str:
    if v == 0:
        return "0"
    if abs(v) >= 1e6 or abs(v) < 1e-2:
        return f"{v:.0e}".replace("e+0", "e").replace("e-0", "e-")
    return f"{v:g}"


def svg_lines(path: str, title: str, xlabel: str, ylabel: str, series: list[tuple[str, list[float], list[float]]],
              logx: bool = False, logy: bool = False, hline: float | None = None) -> None:
    """series: (label, xs, ys).  Points with non-positive values are dropped on log axes."""
    W, H, L, R, T, B = 760, 460, 80, 220, 40, 60
    pts = [(x, y) for _, xs, ys in series for x, y in zip(xs, ys) if (not logx or x > 0) and (not logy or y > 0)]
    if not pts:
        return
    xlo, xhi = min(p[0] for p in pts), max(p[0] for p in pts)
    ylo, yhi = min(p[1] for p in pts), max(p[1] for p in pts)
    if hline is not None:
        ylo, yhi = min(ylo, hline), max(yhi, hline)
    if not logy:
        ylo = min(0.0, ylo)
    xt, yt = _ticks(xlo, xhi, logx), _ticks(ylo, yhi, logy)
    xlo, xhi, ylo, yhi = min(xlo, xt[0]), max(xhi, xt[-1]), min(ylo, yt[0]), max(yhi, yt[-1])
    fx = (lambda v: math.log10(v)) if logx else (lambda v: v)
    fy = (lambda v: math.log10(v)) if logy else (lambda v: v)
    sx = lambda v: L + (fx(v) - fx(xlo)) / max(fx(xhi) - fx(xlo), 1e-12) * (W - L - R)   # noqa: E731
    sy = lambda v: H - B - (fy(v) - fy(ylo)) / max(fy(yhi) - fy(ylo), 1e-12) * (H - T - B)   # noqa: E731
    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="12">',
         f'<rect width="{W}" height="{H}" fill="white"/>', f'<text x="{L}" y="22" font-size="15">{title}</text>']
    for v in xt:
        o.append(f'<line x1="{sx(v):.1f}" y1="{T}" x2="{sx(v):.1f}" y2="{H - B}" stroke="#ddd"/>')
        o.append(f'<text x="{sx(v):.1f}" y="{H - B + 16}" text-anchor="middle">{_fmt(v)}</text>')
    for v in yt:
        o.append(f'<line x1="{L}" y1="{sy(v):.1f}" x2="{W - R}" y2="{sy(v):.1f}" stroke="#ddd"/>')
        o.append(f'<text x="{L - 6}" y="{sy(v) + 4:.1f}" text-anchor="end">{_fmt(v)}</text>')
    o.append(f'<rect x="{L}" y="{T}" width="{W - L - R}" height="{H - T - B}" fill="none" stroke="black"/>')
    o.append(f'<text x="{(L + W - R) / 2}" y="{H - 16}" text-anchor="middle">{xlabel}</text>')
    o.append(f'<text x="18" y="{(T + H - B) / 2}" text-anchor="middle" transform="rotate(-90 18 {(T + H - B) / 2})">{ylabel}</text>')
    if hline is not None:
        o.append(f'<line x1="{L}" y1="{sy(hline):.1f}" x2="{W - R}" y2="{sy(hline):.1f}" stroke="red" stroke-dasharray="5,4"/>')
    for i, (label, xs, ys) in enumerate(series):
        c = COLORS[i % len(COLORS)]
        p = sorted((x, y) for x, y in zip(xs, ys) if (not logx or x > 0) and (not logy or y > 0))
        if not p:
            continue
        o.append(f'<polyline fill="none" stroke="{c}" stroke-width="2" points="' + " ".join(f"{sx(x):.1f},{sy(y):.1f}" for x, y in p) + '"/>')
        o += [f'<circle cx="{sx(x):.1f}" cy="{sy(y):.1f}" r="3" fill="{c}"/>' for x, y in p]
        o.append(f'<rect x="{W - R + 12}" y="{T + 18 * i}" width="12" height="12" fill="{c}"/>')
        o.append(f'<text x="{W - R + 30}" y="{T + 18 * i + 11}">{label}</text>')
    o.append("</svg>")
    with open(path, "w") as f:
        f.write("\n".join(o))


def clean_rows(path: str) -> list[dict]:
    """CleanData (visualisation.py:17-34): strip the column names, split the resolution, sort by pixel count."""
    rows = []
    with open(path, newline="") as f:
        rd = csv.reader(f)
        hdr = [h.strip() for h in next(rd)]
        for r in rd:
            if len(r) < len(hdr):
                continue
            d = {k: v.strip() for k, v in zip(hdr, r)}
            w, h = d["Resolution"].split("x")
            d["Pixel Count"] = int(w) * int(h)
            for k in hdr[3:]:
                if k != "Method":
                    try:
                        d[k] = float(d[k])
                    except ValueError:
                        pass
            rows.append(d)
    rows.sort(key=lambda d: d["Pixel Count"])
    return rows


def plots_from_csv(path: str, out: str) -> list[str]:
    rows = clean_rows(path)
    made = []
    methods = sorted({d.get("Method", "") for d in rows})
    for m in methods:
        rs = [d for d in rows if d.get("Method", "") == m]
        tag = m.lower() or "all"
        px = [d["Pixel Count"] for d in rs]
        col = lambda k: [d[k] for d in rs]   # noqa: E731
        # CPUvsOpenCLEndtoEnd + KernelvsOpenCLTotal (visualisation.py:36-58) in one chart
        p = os.path.join(out, f"{tag}_times.svg")
        svg_lines(p, f"{m or 'all methods'}: CPU path vs B200 path", "pixels per image", "average time (ms)",
                  [("CPU (Comparator)", px, col("avg_CPU_Time_ms")), ("GPU end to end", px, col("avg_OpenCL_Time_ms")),
                   ("upload + kernel + download", px, col("avg_OpenCL_kernel_operation_ms")), ("kernel", px, col("avg_OpenCL_kernel_ms")),
                   ("upload", px, col("avg_OpenCL_kernel_write_ms")), ("download", px, col("avg_OpenCL_kernel_read_ms"))], logx=True, logy=True)
        made.append(p)
        # SpeedFactorEndtoEnd / SpeedFactorOperation (visualisation.py:60-84)
        sp = [c / g if g > 0 else 0 for c, g in zip(col("avg_CPU_Time_ms"), col("avg_OpenCL_Time_ms"))]
        so = [c / g if g > 0 else 0 for c, g in zip(col("avg_CPU_Time_ms"), col("avg_OpenCL_kernel_operation_ms"))]
        p = os.path.join(out, f"{tag}_speedup.svg")
        svg_lines(p, f"{m or 'all methods'}: speed-up over the CPU path", "pixels per image", "speed-up factor",
                  [("end to end", px, sp), ("upload + kernel + download", px, so)], logx=True, logy=True, hline=1.0)
        made.append(p)
        if "Mpix_s" in rs[0]:
            p = os.path.join(out, f"{tag}_throughput.svg")
            svg_lines(p, f"{m}: throughput and share of the HBM peak", "pixels per image", "Mpixel/s  |  % of HBM peak (kernel only)",
                      [("Mpixel/s (upload + kernel + download)", px, col("Mpix_s")), ("% of HBM peak, kernel", px, col("pct_hbm_peak"))], logx=True, logy=True)
            made.append(p)
    # MAE (visualisation.py:86-89): one line per method; all zero here by construction (bit-exact)
    p = os.path.join(out, "error_mae.svg")
    svg_lines(p, "GPU vs CPU path: mean absolute error (0 = bit-exact)", "pixels per image", "MAE",
              [(m or "all", [d["Pixel Count"] for d in rows if d.get("Method", "") == m], [d["Error_MAE"] for d in rows if d.get("Method", "") == m])
               for m in methods], logx=True)
    made.append(p)
    return made


def bench_lines(paths: list[str]) -> list[dict]:
    out = []
    for p in paths:
        with open(p) as f:
            txt = f.read()
        try:
            j = json.loads(txt)
            cand = j if isinstance(j, list) else [j]
        except json.JSONDecodeError:
            cand = [json.loads(ln) for ln in txt.splitlines() if ln.strip().startswith("{")]
        for c in cand:
            for d in (c, c.get("parsed") if isinstance(c, dict) else None):
                if isinstance(d, dict) and "n_gpus" in d and "value" in d and d.get("impl") != "reference":
                    out.append(d)
            if isinstance(c, dict) and isinstance(c.get("runs"), list):   # a SCALE file: {"runs": [{"parsed": {...}}, ...]}
                out += [r["parsed"] for r in c["runs"] if isinstance(r.get("parsed"), dict) and "n_gpus" in r["parsed"]]
    return sorted(out, key=lambda d: d["n_gpus"])


def plots_from_bench(paths: list[str], out: str) -> list[str]:
    ls = bench_lines(paths)
    if not ls:
        return []
    n = [d["n_gpus"] for d in ls]
    res = [d["value"] / 1e3 for d in ls]
    e2e = [d.get("e2e", {}).get("value", 0) / 1e3 for d in ls]
    nv = [d.get("e2e_nv12", {}).get("value", 0) / 1e3 for d in ls]
    base = res[0] / n[0]
    made = []
    p = os.path.join(out, "scaling_resident.svg")
    svg_lines(p, "fused 4K pipeline, frames resident in HBM (weak scaling, 32 frames per GPU)", "GPUs", "Gpixel/s, whole job",
              [("measured", n, res), ("linear from the smallest run", n, [base * k for k in n])])
    made.append(p)
    p = os.path.join(out, "scaling_end_to_end.svg")
    b2 = e2e[0] / n[0] if e2e[0] else 0
    svg_lines(p, "fused 4K pipeline, host buffers in and out (H2D + kernel + D2H per step)", "GPUs", "Gpixel/s, whole job",
              [("RGB8 in (3 B/px over PCIe)", n, e2e), ("NV12 in (1.5 B/px over PCIe)", n, nv), ("linear from the smallest run (RGB8)", n, [b2 * k for k in n])])
    made.append(p)
    c5 = [(d["n_gpus"], d["config5"]) for d in ls if isinstance(d.get("config5"), dict)]
    if c5:
        p = os.path.join(out, "config5_8k.svg")
        svg_lines(p, "7680x4320 through one context over N devices (host buffers)", "devices in the context", "Gpixel/s",
                  [("one frame as N row bands", [k for k, _ in c5], [c["banded_1frame_e2e"]["mpx_s"] / 1e3 for _, c in c5]),
                   ("16 frames sharded over N devices", [k for k, _ in c5], [c["sharded_batch16_e2e"]["mpx_s"] / 1e3 for _, c in c5])])
        made.append(p)
    return made


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", action="append", default=[])
    ap.add_argument("--bench", nargs="*", default=[])
    ap.add_argument("--out", default="plots")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    made = []
    for c in a.csv:
        made += plots_from_csv(c, a.out)
    made += plots_from_bench(a.bench, a.out)
    for p in made:
        print("wrote", p)
    if not made:
        raise SystemExit("nothing to plot: give --csv and/or --bench")


if __name__ == "__main__":
    main()
