#!/usr/bin/env python
"""bench.py -- headline benchmark: fused gray -> 5x5 Gaussian -> Sobel on 4K RGB frames.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One process per GPU (the driver launches N > 1 with torch.distributed.run).  Frames are
independent, so each rank owns whole frames (BASELINE config 4: 256 frames of 3840x2160 RGB8 over 8
GPUs = 32 frames per GPU); there is no data-path collective -- torch.distributed only supplies the
barrier and the max-over-ranks reduction of the device-timed duration.  Weak scaling: every rank
always processes 32 frames per step.

A "step" is one pass of the fused kernel over the rank's resident 32-frame batch (796 MB in +
265 MB out, far larger than the 126 MB L2, so every step streams from HBM).  `value` is
whole-job Mpixel/s with inputs resident in HBM; `e2e` is the same metric through the public
host-buffer API (rip_process_host: pinned host frames -> H2D -> kernel -> D2H every step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, CN = 3840, 2160, 3
FRAMES_PER_GPU = 32
KSIZE, SIGMA = 5, 1.0
ALGO_BYTES_PER_PX = 4  # 3 in (RGB8) + 1 out (u8)   -- SURVEY.md 8(d)
WORKLOAD = (f"fused gray->gauss{KSIZE}x{KSIZE}(sigma={SIGMA})->sobel, synthetic RGB8 {W}x{H}, "
            f"{FRAMES_PER_GPU} frames per GPU per step (BASELINE config 4 shard: 256 frames over 8 GPUs)")


def synth_batch(n: int, seed0: int) -> np.ndarray:
    """uniform iid u8 frames, seed = 0xB200 + 1000*config + frame."""
    out = np.empty((n, H, W, CN), np.uint8)
    for i in range(n):
        out[i] = np.random.default_rng(seed0 + i).integers(0, 256, (H, W, CN), dtype=np.uint8)
    return out


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one fused-kernel launch from the committed ncu
    capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("fused_4k_32frames_dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(frames: np.ndarray, weights: np.ndarray, budget_s: float = 12.0) -> dict:
    """The oracle (CPU restatement of the reference's CPU paths) on a bounded sample, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)   # passed explicitly: torchrun sets OMP_NUM_THREADS=1
    t0 = time.perf_counter()
    O.fused(frames[0], KSIZE, weights=weights, threads=cores)
    t1 = time.perf_counter() - t0
    n = int(max(1, min(frames.shape[0], budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    for i in range(n):
        O.fused(frames[i], KSIZE, weights=weights, threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n * W * H / dt / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": "port",
            "sample": f"{n} of the {frames.shape[0]} 4K frames of one step, OpenMP over rows on {cores} threads "
                      f"({dt / n * 1e3:.0f} ms per frame); single-thread reference behaviour is ~{cores}x slower"}


def run_reference(args, rank: int) -> None:
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference
    needs OpenCV C++ + OpenCL and cannot be built here), all host threads, same config/metric."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = 2  # bounded sample: 2 of the 32 frames per step
    frames = synth_batch(per_step, 0xB200 + 4000)
    w = O.gauss_weights(KSIZE, SIGMA)
    for _ in range(args.warmup):
        O.fused(frames[0], KSIZE, weights=w, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(per_step):
            O.fused(frames[i], KSIZE, weights=w, threads=cores)
    dt = time.perf_counter() - t0
    mpx = args.steps * per_step * W * H / dt / 1e6
    line = {"impl": "reference", "metric": "fused_4k_throughput", "value": mpx, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step_sample": f"{per_step} frames per step (bounded sample of the 32-frame step)"},
            "frames_per_s": mpx * 1e6 / (W * H),
            "cpu_baseline": {"value": mpx, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} 4K frames per step x {args.steps} steps, OpenMP on {cores} threads"},
            "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: min(steps, 10))")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import rip_b200 as rip
    if rip.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: librip_cuda has no CPU fallback")
    dev = local_rank % rip.device_count()

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one process per GPU: host thread and pinned frame buffers on the GPU's own NUMA node (matters for the
    # end-to-end leg at 8 GPUs, where every rank streams 50 GB/s from host memory)
    numa = rip.bind_host_to_device_numa(dev) if world > 1 else {}

    L = rip.lib()
    weights = rip.gauss_weights(KSIZE, SIGMA)
    n = FRAMES_PER_GPU
    px_per_step_rank = n * W * H

    # ---- inputs: synthetic frames in pinned host memory, uploaded once for the resident leg ----
    pin_in = rip.PinnedBuffer(n * H * W * CN)
    pin_out = rip.PinnedBuffer(n * H * W)
    frames = pin_in.array.reshape(n, H, W, CN)
    frames[:] = synth_batch(n, 0xB200 + 4000 + 1000 * rank)
    d_in = rip.DeviceBuffer(frames.nbytes, dev).upload(frames)
    d_out = rip.DeviceBuffer(n * H * W, dev)
    import ctypes as C
    stream = C.c_void_p()
    rip.check(L.rip_stream_create(dev, C.byref(stream)))

    def step():
        rip.fused_dev(d_in.ptr, d_out.ptr, W, H, n, rip.FMT_RGB8, KSIZE, weights, device=dev, stream=stream)

    # ---- resident (kernel) leg ----
    # nvidia-smi samples every 100 ms and the timed region lasts milliseconds, so the sampler runs from
    # before the warm-up to the end of the end-to-end leg (the GPU is busy throughout that window)
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    rip.check(L.rip_stream_sync(dev, stream))
    e0, e1 = rip.Event(dev), rip.Event(dev)
    launches0 = rip.launch_count()
    barrier()
    rip.check(L.rip_device_sync(dev))
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.sync()
    rip.check(L.rip_device_sync(dev))
    barrier()
    launches = rip.launch_count() - launches0
    ms_local = e0.elapsed_ns(e1) / 1e6
    ms_total = max_over_ranks(ms_local)

    # correctness spot check outside the timed region: first frame of the resident output vs oracle
    out0 = d_out.download((H, W))

    # ---- end-to-end leg: public host-buffer API, pinned host frames, H2D + kernel + D2H per step ----
    ctx = rip.Context([dev])
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    out_host = pin_out.array.reshape(n, H, W)
    for _ in range(2):
        ctx.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=KSIZE, weights=weights, out=out_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=KSIZE, weights=weights, out=out_host)
    t_e2e_local = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e_local)
    e2e_ok = bool(np.array_equal(out_host[0], out0))
    ctx.close()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps + end-to-end leg (100 ms nvidia-smi samples)"

    if rank == 0:
        mpx = world * px_per_step_rank * args.steps / (ms_total / 1e3) / 1e6
        ms_per_step = ms_total / args.steps
        peak, peak_src = measured_peak_gbs()
        achieved = px_per_step_rank * ALGO_BYTES_PER_PX / (ms_local / args.steps / 1e3) / 1e9  # this rank's kernel
        e2e_mpx = world * px_per_step_rank * e2e_steps / t_e2e / 1e6
        line = {
            "metric": "fused_4k_throughput", "value": mpx, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": n, "width": W, "height": H, "in_format": "RGB8",
                       "ksize": KSIZE, "sigma": SIGMA, "parallelism": f"frame-sharded x{world}, no collective",
                       "l2_policy": "batch (1.06 GB per step) larger than the 126 MB L2; no flush needed",
                       "dtype_note": "u8 in/out; gray in u32 integer arithmetic, blur/Sobel in fp32"},
            "frames_per_s": mpx * 1e6 / (W * H),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic_bytes(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": px_per_step_rank * ALGO_BYTES_PER_PX,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": e2e_mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": int(frames.nbytes) * world,
                    "d2h_bytes_per_step": int(n * H * W) * world, "steps": e2e_steps,
                    "frames_per_s": e2e_mpx * 1e6 / (W * H), "matches_resident_output": e2e_ok,
                    "api": "rip_process_host (pinned host buffers, 3 chunk streams per device)",
                    "host_numa_binding": numa or None},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(frames, weights)
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as O
            line["parity_spot_check"] = bool(np.array_equal(out0, O.fused(frames[0], KSIZE, weights=weights, threads=0)))
        print(json.dumps(line), flush=True)

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
