#!/usr/bin/env python
"""Replayed (pixel, channel) entries per pixel of the stand-alone Gaussian kernels on the reference's images: python tools/prof_blur_stats.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rip_b200 as rip  # noqa: E402

d = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "images.npz")))
rng = np.random.default_rng(5)
imgs = {n: d[n] for n in ("Artemis_large1024.bgr", "Tulips_medium640.bgr")}
imgs["noise"] = rng.integers(0, 256, (1023, 683, 3), dtype=np.uint8)
for name, bgr in imgs.items():
    img = np.ascontiguousarray(np.concatenate([bgr[..., ::-1], np.full(bgr.shape[:2] + (1,), 255, np.uint8)], -1))
    h, w = img.shape[:2]
    d_in = rip.DeviceBuffer(img.nbytes).upload(img)
    d_out = rip.DeviceBuffer(img.nbytes)
    for k, s in ((5, 1.0), (17, 6.0)):
        wt = rip.gauss_weights(k, s)
        for force in ("RIP_BLUR_TILED", "RIP_BLUR_STREAM"):
            rip.set_option(force, 1)
            rip.slow_path_stats(True)
            rip.gauss_dev(d_in.ptr, d_out.ptr, w, h, 1, 4, k, wt)
            n = rip.slow_path_stats(False)
            rip.set_option(force, 0)
            print(f"{name} {w}x{h} {k}x{k} {force}: {n} replay entries = {n / (h * w):.4f} per pixel")
