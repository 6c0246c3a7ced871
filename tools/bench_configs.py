#!/usr/bin/env python
"""Kernel times of the five BASELINE.json configurations on device-resident data (CUDA events, median of
N launches after warm-up), with algorithmic GB/s (SURVEY.md 8d bytes per pixel) and the fraction of the
measured HBM peak.  Correctness of every configuration is the job of tests/; this tool only times.

    python tools/bench_configs.py [--gpus N] > profiles/rNN_configs.txt
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rip_b200 as rip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--launches", type=int, default=21)
a = ap.parse_args()
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
rng = np.random.default_rng(0xB200)


def timed(fn, n=a.launches, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = rip.Event(), rip.Event()
        e0.record(); fn(); e1.record(); e1.sync()
        ts.append(e0.elapsed_ns(e1) / 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def row(name, us, best, px, bpp):
    gbs = px * bpp / us / 1e3
    print(f"{name:62s} {us:9.1f} us (best {best:8.1f})  {px / us:10.0f} Mpx/s  {gbs:8.1f} GB/s  {100 * gbs / PEAK:5.1f} % of {PEAK:.0f}")


def dev_buf(arr):
    return rip.DeviceBuffer(arr.nbytes).upload(arr)


print(f"# {rip.device_info(0).name.decode()}, HBM peak {PEAK} GB/s (MEASURED_PEAKS.json); median of {a.launches} launches, resident data")
# config 1: gray 640x512 RGBA -> (g,g,g,255)
img = rng.integers(0, 256, (512, 640, 4), dtype=np.uint8)
d_in, d_out = dev_buf(img), rip.DeviceBuffer(img.nbytes)
us, best = timed(lambda: rip.gray_dev(d_in.ptr, d_out.ptr, 640, 512, 1, rip.FMT_RGBA8, rip.GRAY_OUT_RGBA))
row("config 1  gray 640x512 RGBA -> (g,g,g,255), 1 frame (launch-bound)", us, best, 640 * 512, 8)
big = rng.integers(0, 256, (64, 1080, 1920, 4), dtype=np.uint8)
d_in, d_out = dev_buf(big), rip.DeviceBuffer(big.nbytes)
us, best = timed(lambda: rip.gray_dev(d_in.ptr, d_out.ptr, 1920, 1080, 64, rip.FMT_RGBA8, rip.GRAY_OUT_RGBA))
row("          gray 1920x1080 RGBA -> (g,g,g,255), batch 64", us, best, 64 * 1920 * 1080, 8)
# config 2: Gaussian 5x5 sigma 1.0 on 683x1023 RGBA; and the reference default K=17 sigma 6
img = rng.integers(0, 256, (1023, 683, 4), dtype=np.uint8)
img = np.ascontiguousarray(img[:, :680])   # keep the row pitch a multiple of 16 bytes irrelevant: tightly packed
d_in, d_out = dev_buf(img), rip.DeviceBuffer(img.nbytes)
for k, s in ((5, 1.0), (5, 1.5), (17, 6.0)):
    w = rip.gauss_weights(k, s)
    us, best = timed(lambda: rip.gauss_dev(d_in.ptr, d_out.ptr, img.shape[1], img.shape[0], 1, 4, k, w))
    row(f"config 2  exact Gaussian {k}x{k} sigma {s} on {img.shape[1]}x{img.shape[0]} RGBA, 1 frame", us, best, img.shape[0] * img.shape[1], 8)
big4 = rng.integers(0, 256, (16, 1080, 1920, 4), dtype=np.uint8)
d_in, d_out = dev_buf(big4), rip.DeviceBuffer(big4.nbytes)
w = rip.gauss_weights(5, 1.0)
us, best = timed(lambda: rip.gauss_dev(d_in.ptr, d_out.ptr, 1920, 1080, 16, 4, 5, w), n=7)
row("          exact Gaussian 5x5 on 1920x1080 RGBA, batch 16, noise in all four channels", us, best, 16 * 1920 * 1080, 8)
big4[..., 3] = 255   # what the reference uploads: cv::COLOR_BGR2RGBA frames, alpha = 255 throughout
d_in.upload(big4)
for k, s in ((5, 1.0), (9, 2.5), (17, 6.0)):
    w = rip.gauss_weights(k, s)
    us, best = timed(lambda: rip.gauss_dev(d_in.ptr, d_out.ptr, 1920, 1080, 16, 4, k, w), n=7)
    row(f"          exact Gaussian {k}x{k} sigma {s} on 1920x1080 RGBA, batch 16, alpha = 255", us, best, 16 * 1920 * 1080, 8)
# config 3: Sobel on 1080p RGB, batch 64
big = rng.integers(0, 256, (64, 1080, 1920, 3), dtype=np.uint8)
d_in, d_out = dev_buf(big), rip.DeviceBuffer(64 * 1080 * 1920)
us, best = timed(lambda: rip.sobel_dev(d_in.ptr, d_out.ptr, 1920, 1080, 64, rip.FMT_RGB8))
row("config 3  gray->Sobel 1920x1080 RGB8, batch 64", us, best, 64 * 1920 * 1080, 4)
# config 4: fused 4K RGB, 32 frames (one GPU's shard of the 256-frame batch)
big = rng.integers(0, 256, (32, 2160, 3840, 3), dtype=np.uint8)
d_in, d_out = dev_buf(big), rip.DeviceBuffer(32 * 2160 * 3840)
for s in (1.0, 1.5):
    w = rip.gauss_weights(5, s)
    us, best = timed(lambda: rip.fused_dev(d_in.ptr, d_out.ptr, 3840, 2160, 32, rip.FMT_RGB8, 5, w))
    row(f"config 4  fused gray->5x5(sigma {s})->Sobel 3840x2160 RGB8, 32 frames", us, best, 32 * 3840 * 2160, 4)
# smooth (natural-image-like) content: same kernel, same guard-band rate
yy, xx = np.mgrid[0:2160, 0:3840]
sm = (128 + 60 * np.sin(xx / 97.0) + 50 * np.cos(yy / 61.0) + rng.integers(-2, 3, (2160, 3840))).clip(0, 255).astype(np.uint8)
smooth = np.ascontiguousarray(np.stack([sm, np.roll(sm, 7, 1), np.roll(sm, 13, 0)], -1)[None].repeat(32, 0))
d_in2 = dev_buf(smooth)
w = rip.gauss_weights(5, 1.0)
us, best = timed(lambda: rip.fused_dev(d_in2.ptr, d_out.ptr, 3840, 2160, 32, rip.FMT_RGB8, 5, w))
row("          same, smooth content (sinusoids + noise)", us, best, 32 * 3840 * 2160, 4)
del d_in2, smooth
# config 5: fused 8K single frame, whole frame on one GPU
one = rng.integers(0, 256, (4320, 7680, 3), dtype=np.uint8)
d_in, d_out = dev_buf(one), rip.DeviceBuffer(4320 * 7680)
us, best = timed(lambda: rip.fused_dev(d_in.ptr, d_out.ptr, 7680, 4320, 1, rip.FMT_RGB8, 5, w))
row("config 5  fused 7680x4320 RGB8, 1 frame, whole frame on 1 GPU (fits L2)", us, best, 7680 * 4320, 4)
del d_in, d_out
# config 5: throughput over a batch of 16 8K frames resident on one GPU (2.1 GB in)
d_in, d_out = rip.DeviceBuffer(16 * one.nbytes), rip.DeviceBuffer(16 * 4320 * 7680)
for i in range(16):
    d_in.upload(np.roll(one, 17 * i, axis=1), offset=i * one.nbytes)
us, best = timed(lambda: rip.fused_dev(d_in.ptr, d_out.ptr, 7680, 4320, 16, rip.FMT_RGB8, 5, w), n=9)
row("config 5  fused 7680x4320 RGB8, batch 16 on 1 GPU", us, best, 16 * 7680 * 4320, 4)
del d_in, d_out
# camera-format input: NV12 frames, the luma plane is the image (2 algorithmic bytes per pixel)
nv = rng.integers(0, 256, (32, 2160 * 3 // 2, 3840), dtype=np.uint8)
d_in, d_out = dev_buf(nv), rip.DeviceBuffer(32 * 3840 * 2160)
us, best = timed(lambda: rip.fused_dev(d_in.ptr, d_out.ptr, 3840, 2160, 32, rip.FMT_NV12, 5, w))
row("          fused 5x5->Sobel on NV12 3840x2160 (luma plane), 32 frames", us, best, 32 * 3840 * 2160, 2)
us, best = timed(lambda: rip.sobel_dev(d_in.ptr, d_out.ptr, 3840, 2160, 32, rip.FMT_NV12))
row("          Sobel on NV12 3840x2160 (luma plane), 32 frames", us, best, 32 * 3840 * 2160, 2)
del d_in, d_out, nv
# informational CPU rows for config 3 (one 1080p frame): the genuine OpenCV calls of EdgeDetection.cpp:231-240
# where Python cv2 is installed, and the oracle port, single thread
try:
    import cv2
    g1080 = rng.integers(0, 256, (1080, 1920), dtype=np.uint8)
    kx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
    cv2.setNumThreads(1)
    def cv_sobel():
        gx = cv2.filter2D(g1080, cv2.CV_32F, kx)
        gy = cv2.filter2D(g1080, cv2.CV_32F, kx.T.copy())
        m = cv2.magnitude(gx, gy)
        return m.round().clip(0, 255).astype(np.uint8)
    cv_sobel()
    t0 = time.perf_counter()
    for _ in range(5):
        cv_sobel()
    dt = (time.perf_counter() - t0) / 5
    print(f"config 3  (CPU, informational) cv2 {cv2.__version__} filter2D x2 + magnitude on one 1920x1080 gray frame, 1 thread: "
          f"{dt * 1e3:.2f} ms ({1920 * 1080 / dt / 1e6:.0f} Mpx/s)")
except Exception as e:  # noqa: BLE001
    print(f"config 3  (CPU, informational) cv2 not available: {e}")
# row bands through the host pipeline (H2D of band + halo, kernel, D2H), 1..N GPUs: wall-clock latency
pin = rip.PinnedBuffer(one.nbytes)
pin.array[:] = one.reshape(-1)
src = pin.array.reshape(4320, 7680, 3)
outp = rip.PinnedBuffer(4320 * 7680)
dst = outp.array.reshape(1, 4320, 7680)
n = 1
while n <= a.gpus:
    ctx = rip.Context(list(range(n)))
    for _ in range(3):
        ctx.process(src, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w, out=dst, banded=True)
    ts = []
    for _ in range(9):
        t0 = time.perf_counter()
        ctx.process(src, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w, out=dst, banded=True)
        ts.append((time.perf_counter() - t0) * 1e6)
    ts.sort()
    print(f"config 5  fused 7680x4320, row bands + 3-row halo over {n} GPU(s), host buffers end to end: "
          f"{ts[len(ts) // 2]:9.1f} us per frame ({7680 * 4320 / ts[len(ts) // 2]:.0f} Mpx/s)")
    ctx.close()
    n *= 2
