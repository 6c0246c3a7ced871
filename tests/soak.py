#!/usr/bin/env python
"""Randomised parity soak of the CUDA kernels against the oracle (test infrastructure; not collected by pytest: run once per build on a GPU box).

    python tests/soak.py [seconds=60] [seed=1]

Random kernel sizes, sigmas, shapes (narrow, ragged, batches), formats and content (noise, plateaus of every small size, constant
channels, black, near-constant alpha), the streaming kernels forced or chosen by size; every output must equal the oracle's bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rip_b200 as rip  # noqa: E402
from oracle import oracle as O  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = rip.Context([0])
BIG = int(os.environ.get("SOAK_BIG", "1"))   # scale of the random shapes (1: up to 260 x 400; 5: up to 1300 x 2000 -- several segments and band groups)


def content(h, w, c):
    kind = rng.integers(0, 6)
    img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    if kind == 1:      # dark, few levels: plateaus everywhere
        img = rng.integers(0, 4, (h, w, c), dtype=np.uint8)
    elif kind == 2:    # blocky: constant patches of random small sizes
        img = np.kron(rng.integers(0, 256, (-(-h // 6), -(-w // 7), c), dtype=np.uint8), np.ones((6, 7, 1), np.uint8))[:h, :w]
        img = np.ascontiguousarray(img)
    elif kind == 3:    # smooth gradient + 1 LSB of noise
        yy, xx = np.mgrid[0:h, 0:w]
        base = (xx * 0.7 + yy * 0.3).astype(np.int64)
        img = ((base[..., None] + rng.integers(0, 2, (h, w, c))) % 256).astype(np.uint8)
    for _ in range(int(rng.integers(0, 5))):   # rectangles of one value per channel
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        sy, sx = int(rng.integers(1, max(2, h // 2))), int(rng.integers(1, max(2, w // 2)))
        img[y: y + sy, x: x + sx] = rng.integers(0, 256, c, dtype=np.uint8) if rng.integers(0, 3) else rng.choice([0, 255])
    if c == 4:
        mode = rng.integers(0, 4)
        if mode == 0:
            img[..., 3] = 255
        elif mode == 1:
            img[..., 3] = 255
            img[rng.random((h, w)) < 0.01, 3] = 254
        elif mode == 2:
            img[..., 3] = 0
    return np.ascontiguousarray(img)


t0 = time.time()
n = 0
fails = 0
while time.time() - t0 < budget:
    n += 1
    op = rng.choice(["gauss", "gauss", "fused", "sobel"]) if os.environ.get("SOAK_OP", "") == "" else os.environ["SOAK_OP"]
    for name in ("RIP_BLUR_STREAM", "RIP_BLUR_TILED"):
        rip.set_option(name, 0)
    if op == "gauss":
        k = int(rng.choice([5, 5, 9, 17, 3, 7]))
        sigma = float(rng.choice([0.6, 1.0, 1.5, 2.5, 4.0, 6.0]))
        nf = int(rng.integers(1, 4))
        h, w = int(rng.integers(k, 260 * BIG)), int(rng.integers(k, 400 * BIG))
        force = rng.choice(["", "RIP_BLUR_STREAM", "RIP_BLUR_TILED"])
        if force:
            rip.set_option(force, 1)
        imgs = np.stack([content(h, w, 4) for _ in range(nf)])
        wt = rip.gauss_weights(k, sigma)
        got = ctx.process(imgs, rip.OP_GAUSSIAN, rip.FMT_RGBA8, ksize=k, weights=wt)
        want = np.stack([O.blur(imgs[i], k, weights=wt, threads=0) for i in range(nf)])
        ok = np.array_equal(got, want)
        desc = f"gauss K={k} sigma={sigma} {nf}x{h}x{w} {force}"
        if not ok:
            bad = np.argwhere(got != want)
            desc += f": {len(bad)} bytes differ, first (frame, y, x, ch) = {bad[:6].tolist()}, got {[int(got[tuple(b)]) for b in bad[:6]]} want {[int(want[tuple(b)]) for b in bad[:6]]}"
            desc += f", input there {[imgs[b[0], b[1], b[2]].tolist() for b in bad[:3]]}, channels hit {sorted(set(bad[:, 3].tolist()))}, rows {sorted(set(bad[:, 1].tolist()))[:12]}"
    else:
        fmt, c = [(rip.FMT_RGB8, 3), (rip.FMT_RGBA8, 4), (rip.FMT_GRAY8, 1), (rip.FMT_BGR8, 3)][int(rng.integers(0, 4))]
        nf = int(rng.integers(1, 3))
        h, w = int(rng.integers(8, 300 * BIG)), int(rng.integers(2, 80 * BIG)) * int(rng.choice([4, 8, 8]))
        sigma = float(rng.choice([1.0, 1.5]))
        imgs = np.stack([content(h, w, c) for _ in range(nf)])
        if c == 1:
            imgs = imgs[..., 0]
        wt = rip.gauss_weights(5, sigma)

        def gray_of(im):
            if c == 1:
                return im
            rgb = im[..., :3] if fmt != rip.FMT_BGR8 else im[..., ::-1]
            return O.gray(np.ascontiguousarray(rgb), threads=0)
        if op == "fused":
            got = ctx.process(imgs, rip.OP_FUSED, fmt, ksize=5, weights=wt)
            ok = all(np.array_equal(got[i], O.sobel(O.blur(gray_of(imgs[i]), 5, weights=wt, threads=0))) for i in range(nf))
        else:
            got = ctx.process(imgs, rip.OP_EDGE, fmt)
            ok = all(np.array_equal(got[i], O.sobel(gray_of(imgs[i]))) for i in range(nf))
        desc = f"{op} fmt={fmt} sigma={sigma} {nf}x{h}x{w}"
    if not ok:
        fails += 1
        print("MISMATCH:", desc, flush=True)
print(f"soak: {n} random cases in {time.time() - t0:.0f} s, {fails} mismatches (seed {seed})")
sys.exit(1 if fails else 0)
