#!/usr/bin/env python
"""Generate tests/golden/ from the reference checkout (run in the build container only).

What it records, and why:
  * images.npz      -- the reference's own fixtures (`/root/reference/images/*.jpg`) decoded with
                       cv2.imread exactly like the reference does (IMREAD_COLOR -> BGR,
                       ProgramHandler.cpp:116 / Comparator.cpp:14; IMREAD_GRAYSCALE,
                       EdgeDetection.cpp:202).  The GPU box has no /root/reference, so the decoded
                       pixels travel instead of the JPEGs.  The two largest non-config images are
                       left out to keep the fixture small; their hashes are still recorded.
  * cv2_sobel.npz   -- outputs of the REAL OpenCV calls the reference's CPU Sobel makes
                       (filter2D x2, magnitude, convertTo(CV_8UC1)), on the fixture images and on
                       seeded synthetic inputs; pins oracle.sobel to OpenCV itself.
  * expected.json   -- (a) the published Error_MAE values with their file:line, (b) sha256 of the
                       oracle's output per image per stage at generation time (regression pin,
                       including the images not shipped), (c) cv2 version.

Usage:  python tools/make_golden.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

# Published Error_MAE (CPU path vs OpenCL buffer path); value, file:line in the reference.
PUBLISHED_MAE = {
    "sobel": {
        "Tulips_square75": [7.4864, "src/EdgeDetection/results/Linux_100_Tulips_sorted_results.csv:2"],
        "Tulips_small240": [3.99475, "src/EdgeDetection/results/Linux_100_Tulips_sorted_results.csv:3"],
        "Tulips_medium640": [2.66722, "src/EdgeDetection/results/Linux_100_Tulips_sorted_results.csv:4"],
        "Tulips_large1024": [2.04212, "src/EdgeDetection/results/Linux_100_Tulips_sorted_results.csv:5"],
        "Artemis_square75": [3.81956, "src/EdgeDetection/results/Linux_100_Artemis_sorted_results.csv:2"],
    },
    "gray": {
        "Artemis_square75": [0.0712889, "src/Grayscale/results/Linux_100_Artemis_sorted_results.csv:2"],
        "Artemis_small240": [0.0823438, "src/Grayscale/results/Linux_100_Artemis_sorted_results.csv:3"],
        "Artemis_medium640": [0.0798595, "src/Grayscale/results/Linux_100_Artemis_sorted_results.csv:4"],
        "Artemis_large1024": [0.0517197, "src/Grayscale/results/Linux_100_Artemis_sorted_results.csv:5"],
        "Tulips_square75": [0.000355556, "src/Grayscale/results/Windows_100_Tulips_sorted_results.csv:2"],
    },
    "blur_k5_s1.5": {
        "Tulips_square75": [0.0, "src/GaussianBlur/results/Linux_100_Tulips_sorted_results.csv:2"],
        "Tulips_small240": [0.0, "src/GaussianBlur/results/Linux_100_Tulips_sorted_results.csv:3"],
        "Tulips_medium640": [0.0, "src/GaussianBlur/results/Linux_100_Tulips_sorted_results.csv:4"],
    },
}

SHIP_BGR = ["Tulips_square75", "Artemis_square75", "Tulips_small240", "Artemis_small240",
            "Tulips_medium640", "Artemis_large1024"]
SHIP_GRAY = ["Tulips_square75", "Artemis_square75", "Tulips_small240", "Artemis_small240",
             "Tulips_medium640"]

SX = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
SY = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], np.float32)


def cv2_sobel(gray: np.ndarray) -> np.ndarray:
    """The reference's CPU Sobel, through OpenCV itself (EdgeDetection.cpp:219-240)."""
    gx = cv2.filter2D(gray, cv2.CV_32F, SX)
    gy = cv2.filter2D(gray, cv2.CV_32F, SY)
    mag = cv2.magnitude(gx, gy)
    # Mat::convertTo(CV_8UC1) == saturate_cast<uchar>(cvRound(v)); cv2.convertScaleAbs(alpha=1)
    # does the same for non-negative input.
    return cv2.convertScaleAbs(mag)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synthetic_sobel_inputs() -> dict[str, np.ndarray]:
    rng = np.random.default_rng(0xB200)
    out = {
        "rand_97x131": rng.integers(0, 256, (97, 131), dtype=np.uint8),
        "rand_2x2": rng.integers(0, 256, (2, 2), dtype=np.uint8),
        "rand_3x64": rng.integers(0, 256, (3, 64), dtype=np.uint8),
        "rand_64x2": rng.integers(0, 256, (64, 2), dtype=np.uint8),
        "hramp_40x300": np.tile((np.arange(300) % 256).astype(np.uint8), (40, 1)),
        "vramp_300x40": np.tile((np.arange(300) % 256).astype(np.uint8)[:, None], (1, 40)),
        "checker_33x35": ((np.indices((33, 35)).sum(0) & 1) * 255).astype(np.uint8),
        "flat7_16x16": np.full((16, 16), 7, np.uint8),
    }
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    cv2.ocl.setUseOpenCL(False)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)

    images, cvs, hashes = {}, {}, {}
    for p in sorted(glob.glob(os.path.join(args.ref, "images", "*.jpg"))):
        name = os.path.basename(p)[:-4]
        bgr = cv2.imread(p, cv2.IMREAD_COLOR)
        gim = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
        rgba = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGBA)
        rgb = np.ascontiguousarray(rgba[..., :3])
        if name in SHIP_BGR:
            images[name + ".bgr"] = bgr
        if name in SHIP_GRAY:
            images[name + ".imread_gray"] = gim
        cs = cv2_sobel(gim)
        assert np.array_equal(cs, O.sobel(gim)), f"oracle.sobel != cv2 on {name}"
        if name in SHIP_GRAY:
            cvs[name] = cs
        hashes[name] = {
            "shape_hw": [int(bgr.shape[0]), int(bgr.shape[1])],
            "bgr": sha(bgr),
            "imread_gray": sha(gim),
            "gray": sha(O.gray(bgr, O.BGR)),
            "blur_rgba_k5_s1.0": sha(O.blur(rgba, 5, 1.0)),
            "blur_rgba_k5_s1.5": sha(O.blur(rgba, 5, 1.5)),
            "sobel_of_imread_gray": sha(cs),
            "sobel_of_gray": sha(O.sobel(O.gray(rgb, O.RGB))),
            "fused_k5_s1.0": sha(O.fused(rgb, 5, 1.0, O.RGB)),
        }
    for k, g in synthetic_sobel_inputs().items():
        cs = cv2_sobel(g)
        assert np.array_equal(cs, O.sobel(g)), f"oracle.sobel != cv2 on {k}"
        cvs["syn." + k + ".in"] = g
        cvs["syn." + k + ".out"] = cs

    np.savez_compressed(os.path.join(gold, "images.npz"), **images)
    np.savez_compressed(os.path.join(gold, "cv2_sobel.npz"), **cvs)
    with open(os.path.join(gold, "expected.json"), "w") as f:
        json.dump({"cv2_version": cv2.__version__, "published_mae": PUBLISHED_MAE,
                   "oracle_sha256": hashes,
                   "gauss_weights_k5_s1.0": [float(x) for x in O.gauss_weights(5, 1.0).ravel()],
                   "gauss_weights_k5_s1.5": [float(x) for x in O.gauss_weights(5, 1.5).ravel()]},
                  f, indent=1, sort_keys=True)
    for fn in ("images.npz", "cv2_sobel.npz", "expected.json"):
        print(fn, os.path.getsize(os.path.join(gold, fn)), "bytes")


if __name__ == "__main__":
    main()
