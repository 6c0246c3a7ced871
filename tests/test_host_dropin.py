"""The C++ host classes (Controller / ProgramHandler / Comparator / FileHandler, package host/) driven the
way the reference's applications drive them: tests/cpp/host_dropin.cpp.

CPU test: the Comparator's CPU paths (the reference's definition of correct output) against the oracle.
GPU test: every method through ProgramHandler -> Controller -> librip_cuda must equal the Comparator's CPU
result bit for bit, plus the batch API and the results CSV."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "opencl-development-real-time-image-processing_b200")
sys.path.insert(0, ROOT)

IMAGES = {"noise": (61, 88), "smooth": (40, 128), "flat": (9, 12)}


def _write_ppm(path, rgb):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (rgb.shape[1], rgb.shape[0]))
        f.write(rgb.tobytes())


def _make_images(workdir):
    os.makedirs(os.path.join(workdir, "images"), exist_ok=True)
    rng = np.random.default_rng(0xB200)
    imgs = {}
    for name, (h, w) in IMAGES.items():
        if name == "noise":
            rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif name == "smooth":
            yy, xx = np.mgrid[0:h, 0:w]
            rgb = np.stack([(xx * 2 + yy) % 256, (xx + yy * 3) % 256, (200 - xx) % 256], -1).astype(np.uint8)
        else:
            rgb = np.full((h, w, 3), 77, np.uint8)
        _write_ppm(os.path.join(workdir, "images", name + ".ppm"), rgb)
        imgs[name] = rgb
    return imgs


def _build(tmp_path):
    exe = str(tmp_path / "host_dropin")
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "include"),
           "-I", cuda_inc, os.path.join(ROOT, "tests", "cpp", "host_dropin.cpp"), "-o", exe, "-L", PKG, "-lrip_host", "-lrip_cuda",
           "-Wl,-rpath," + PKG, "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def _check_cpu_files(workdir, imgs, oracle):
    w = oracle.gauss_weights(5, 1.5)
    for name, rgb in imgs.items():
        h, wd = rgb.shape[:2]
        rgba = np.concatenate([rgb, np.full((h, wd, 1), 255, np.uint8)], -1)
        gray = oracle.gray(rgb)
        want = {"GRAYSCALE": gray, "EDGE": oracle.sobel(gray), "GAUSSIAN": oracle.blur(rgba, 5, weights=w),
                "FUSED": oracle.fused(rgb, 5, weights=w)}
        for method, ref in want.items():
            got = np.fromfile(os.path.join(workdir, f"cpu_{method}_{name}.raw"), np.uint8).reshape(ref.shape)
            assert np.array_equal(got, ref), f"Comparator CPU path {method} differs from the oracle on {name}"


def test_comparator_cpu_paths_match_the_oracle(tmp_path, oracle):
    imgs = _make_images(str(tmp_path))
    exe = _build(tmp_path)
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    _check_cpu_files(str(tmp_path), imgs, oracle)


@pytest.mark.gpu
def test_reference_style_driver_is_exact_on_the_gpu(tmp_path, oracle):
    imgs = _make_images(str(tmp_path))
    exe = _build(tmp_path)
    res = subprocess.run([exe, str(tmp_path), "--gpu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("OK"), res.stdout
    _check_cpu_files(str(tmp_path), imgs, oracle)
    # the reference's CSV schema (FileHandler.cpp:28), one row per (method, image)
    rows = open(os.path.join(str(tmp_path), "results.csv")).read().strip().split("\n")
    assert rows[0].startswith("Timestamp, Image, Resolution, Num_Iterations, avg_CPU_Time_ms, avg_OpenCL_Time_ms")
    assert len(rows) == 1 + 4 * len(IMAGES)
    # gray comes back in the reference's (g,g,g,255) container
    g = np.fromfile(os.path.join(str(tmp_path), "gpu_GRAYSCALE_noise.raw"), np.uint8).reshape(IMAGES["noise"] + (4,))
    assert np.array_equal(g[..., 0], oracle.gray(imgs["noise"])) and np.all(g[..., 3] == 255) and np.array_equal(g[..., 0], g[..., 2])
