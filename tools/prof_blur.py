#!/usr/bin/env python
"""Minimal driver for profiling the stand-alone Gaussian: python tools/prof_blur.py [K=5] [sigma] [frames=16] [launches=4]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rip_b200 as rip  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else (1.0 if k == 5 else 6.0)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 4
rng = np.random.default_rng(1)
h, w = 1080, 1920
img = rng.integers(0, 256, (n, h, w, 4), dtype=np.uint8)
if len(sys.argv) > 5 and sys.argv[5] == "alpha255":   # what real RGBA frames look like: a constant alpha channel
    img[..., 3] = 255
if len(sys.argv) > 5 and sys.argv[5] == "sky":        # alpha 255 and the upper half black (the reference's Artemis_* images)
    img[..., 3] = 255
    img[:, : h // 2, :, :3] = 0
d_in = rip.DeviceBuffer(img.nbytes).upload(img)
d_out = rip.DeviceBuffer(img.nbytes)
wt = rip.gauss_weights(k, sigma)
ts = []
for i in range(launches):
    e0, e1 = rip.Event(), rip.Event()
    e0.record(); rip.gauss_dev(d_in.ptr, d_out.ptr, w, h, n, 4, k, wt); e1.record(); e1.sync()
    ts.append(e0.elapsed_ns(e1) / 1e3)
us = sorted(ts[1:] or ts)[len(ts[1:] or ts) // 2]
px = n * h * w
print(f"gauss {k}x{k} sigma {sigma} on {n} x {w}x{h} RGBA {sys.argv[5] if len(sys.argv) > 5 else 'noise'}: {us:.1f} us ({px / us:.0f} Mpx/s, {px * 8 / us / 1e3:.0f} GB/s algorithmic)")
