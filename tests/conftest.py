"""Shared fixtures.  `gpu` marks tests that need a B200 (run by the driver with `-m gpu`)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/librip_oracle.so) -- the checker, never the thing under test."""
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_images():
    return dict(np.load(os.path.join(GOLDEN, "images.npz")))


@pytest.fixture(scope="session")
def golden_cv2_sobel():
    return dict(np.load(os.path.join(GOLDEN, "cv2_sobel.npz")))


@pytest.fixture(scope="session")
def expected():
    with open(os.path.join(GOLDEN, "expected.json")) as f:
        return json.load(f)


def bgr_to_rgba(bgr: np.ndarray) -> np.ndarray:
    """cv::cvtColor(COLOR_BGR2RGBA) (ProgramHandler.cpp:127): swap R/B, alpha = 255."""
    h, w, _ = bgr.shape
    out = np.empty((h, w, 4), np.uint8)
    out[..., 0] = bgr[..., 2]
    out[..., 1] = bgr[..., 1]
    out[..., 2] = bgr[..., 0]
    out[..., 3] = 255
    return out


def synth_frame(kind: str, h: int, w: int, seed: int, cn: int = 3) -> np.ndarray:
    """Synthetic frames of SURVEY.md 8(d): 'uniform' iid u8, 'smooth' sinusoids + noise."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, (h, w, cn), dtype=np.uint8)
    if kind == "smooth":
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.zeros((h, w, cn), np.float32)
        for c in range(cn):
            acc = np.full((h, w), 127.0, np.float32)
            for _ in range(4):
                fx, fy = rng.uniform(0.002, 0.05, 2)
                ph = rng.uniform(0, 6.28)
                acc += 30.0 * np.sin(fx * xx + fy * yy + ph).astype(np.float32)
            img[..., c] = acc
        img += rng.integers(-2, 3, img.shape)
        return np.clip(img, 0, 255).astype(np.uint8)
    raise ValueError(kind)


@pytest.fixture
def opt():
    """Set a librip_cuda experiment switch (rip_debug_set_option) for one test; every switch touched is put back to 0
    afterwards."""
    import rip_b200 as rip
    touched = []

    def _set(name, value):
        touched.append(name)
        rip.set_option(name, value)

    yield _set
    for name in touched:
        rip.set_option(name, 0)
