#!/usr/bin/env python
"""Stand-alone Gaussian on the reference's Artemis_large1024 image (decoded pixels from tests/golden), tiled against streaming kernel,
also tiled 2x2 and 4x4 to 1366x2046 / 2732x4092: python tools/prof_blur_artemis.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rip_b200 as rip  # noqa: E402

d = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "images.npz")))
for name in ("Artemis_large1024.bgr", "Tulips_medium640.bgr"):
    bgr = d[name]
    base = np.ascontiguousarray(np.concatenate([bgr[..., ::-1], np.full(bgr.shape[:2] + (1,), 255, np.uint8)], -1))
    for rep in (1, 2, 4):
        img = np.ascontiguousarray(np.tile(base, (rep, rep, 1)))
        h, w = img.shape[:2]
        d_in = rip.DeviceBuffer(img.nbytes).upload(img)
        d_out = rip.DeviceBuffer(img.nbytes)
        for k, s in ((5, 1.0), (17, 6.0)):
            wt = rip.gauss_weights(k, s)
            res = []
            for force in ("RIP_BLUR_TILED", "RIP_BLUR_STREAM"):
                rip.set_option(force, 1)
                ts = []
                for i in range(24):
                    e0, e1 = rip.Event(), rip.Event()
                    e0.record(); rip.gauss_dev(d_in.ptr, d_out.ptr, w, h, 1, 4, k, wt); e1.record(); e1.sync()
                    ts.append(e0.elapsed_ns(e1) / 1e3)
                rip.set_option(force, 0)
                ts = sorted(ts[4:])
                res.append(ts[len(ts) // 2])
            print(f"{name} x{rep} {w}x{h} {k}x{k}: tiled {res[0]:7.1f} us   streaming {res[1]:7.1f} us")
