#!/usr/bin/env python
"""Stand-alone Gaussian on ONE frame, tiled kernel against streaming kernel: python tools/prof_blur_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rip_b200 as rip  # noqa: E402

rng = np.random.default_rng(3)
for (h, w) in ((512, 640), (1023, 680), (1080, 1920), (2160, 3840)):
    img = rng.integers(0, 256, (1, h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    d_in = rip.DeviceBuffer(img.nbytes).upload(img)
    d_out = rip.DeviceBuffer(img.nbytes)
    for k, s in ((5, 1.0), (9, 2.5), (17, 6.0)):
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
            res.append((ts[len(ts) // 2], ts[0]))
        print(f"{w}x{h} {k}x{k}: tiled {res[0][0]:7.1f} us (best {res[0][1]:7.1f})   streaming {res[1][0]:7.1f} us (best {res[1][1]:7.1f})")
