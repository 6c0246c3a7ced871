#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, on shapes that take
the fast paths AND the cold paths (flat and grey regions force the guard-band replay and the gray fix in every warp).

    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rip_b200 as rip  # noqa: E402

rng = np.random.default_rng(3)
ctx = rip.Context([0])
w10, w15, w17 = rip.gauss_weights(5, 1.0), rip.gauss_weights(5, 1.5), rip.gauss_weights(17, 6.0)


def content(h, w, cn):
    a = rng.integers(0, 256, (h, w, cn), dtype=np.uint8)
    a[: h // 3] = 200                       # flat: every pixel in the guard band
    a[h // 3: h // 2, : w // 2] = 0         # black
    a[h // 2: h // 2 + 8] = (a[h // 2: h // 2 + 8, :, :1] // 8) * 8   # greys: multiples of 1000 in the gray stage
    return a


n = 0
for (h, w) in ((96, 256), (70, 248), (41, 100), (33, 75)):
    for fmt, cn in ((rip.FMT_RGB8, 3), (rip.FMT_RGBA8, 4)):
        f = np.stack([content(h, w, cn) for _ in range(2)])
        for wt in (w10, w15):
            ctx.process(f, rip.OP_FUSED, fmt, ksize=5, weights=wt); n += 1
        ctx.process(f, rip.OP_EDGE, fmt); n += 1
        ctx.process(f, rip.OP_GRAY, fmt); n += 1
    g = np.stack([content(h, w, 1)[..., 0] for _ in range(2)])
    ctx.process(g, rip.OP_FUSED, rip.FMT_GRAY8, ksize=5, weights=w10); n += 1
    rgba = np.stack([content(h, w, 4) for _ in range(2)])
    for sw in ("BLUR_STREAM", "BLUR_TILED"):
        rip.set_option(sw, 1)
        ctx.process(rgba, rip.OP_GAUSSIAN, rip.FMT_RGBA8, ksize=5, weights=w15); n += 1
        rip.set_option(sw, 0)
    ctx.process(rgba, rip.OP_GAUSSIAN, rip.FMT_RGBA8, ksize=17, weights=w17); n += 1
    ctx.process(g, rip.OP_GAUSSIAN, rip.FMT_GRAY8, ksize=17, weights=w17); n += 1
    ctx.process(rgba, rip.OP_FUSED, rip.FMT_RGBA8, ksize=17, weights=w17); n += 1
one = content(203, 368, 3)
ctx.process(one, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w10, banded=True); n += 1
ctx.close()
print(f"sanitize_run: {n} pipeline calls, {rip.launch_count()} kernel launches")
