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


# ---- results tooling: the headless program, the extended CSV, the plots -----------------------------
EXT_HEADER = ("Timestamp, Image, Resolution, Num_Iterations, avg_CPU_Time_ms, avg_OpenCL_Time_ms, avg_OpenCL_kernel_ms, "
              "avg_OpenCL_kernel_write_ms, avg_OpenCL_kernel_read_ms, avg_OpenCL_kernel_operation_ms, Error_MAE, "
              "Method, max_abs_err, Mpix_s, fps, GBps, pct_hbm_peak, n_gpus")


def _build_headless(tmp_path):
    exe = str(tmp_path / "rip_headless")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "include"),
           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tools", "rip_headless.cpp"), "-o", exe, "-L", PKG, "-lrip_host", "-lrip_cuda",
           "-Wl,-rpath," + PKG, "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def test_headless_program_builds_and_plots_are_written(tmp_path):
    """No GPU: the program links against the in-tree libraries and prints its usage; the plot script turns an extended CSV and
    bench lines into SVG files."""
    exe = _build_headless(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 2 and "rip_headless images" in res.stderr
    csv = tmp_path / "results_extended.csv"
    lines = [EXT_HEADER]
    for method in ("GRAYSCALE", "FUSED"):
        for (w, h, cpu, gpu) in ((75, 75, 0.2, 0.05), (640, 512, 9.0, 0.12), (1920, 1080, 60.0, 0.4)):
            lines.append(f"2026-01-01 00:00:00, img_{w}.ppm, {w}x{h}, 10, {cpu}, {gpu}, {gpu / 10}, {gpu / 4}, {gpu / 8}, {gpu / 2}, 0, "
                         f"{method}, 0, {w * h / (gpu / 2) / 1e3}, {1e3 / gpu}, 100.0, 1.5, 1")
    csv.write_text("\n".join(lines) + "\n")
    bench = tmp_path / "bench.jsonl"
    bench.write_text("\n".join('{"metric": "fused_4k_throughput", "value": %d, "n_gpus": %d, "e2e": {"value": %d}, "e2e_nv12": {"value": %d}}'
                               % (650000 * n, n, 17000 * min(n, 2.4), 30000 * min(n, 2)) for n in (1, 2, 4, 8)) + "\n")
    out = tmp_path / "plots"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "plot_results.py"), "--csv", str(csv), "--bench", str(bench), "--out", str(out)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    made = sorted(os.listdir(out))
    for want in ("fused_times.svg", "fused_speedup.svg", "fused_throughput.svg", "grayscale_times.svg", "error_mae.svg", "scaling_resident.svg",
                 "scaling_end_to_end.svg"):
        assert want in made, made
        assert (out / want).read_text().startswith("<svg")


@pytest.mark.gpu
def test_headless_program_on_the_gpu(tmp_path):
    """PerformOnImages without a window: every method on every image, the extended results table, then the streaming loop."""
    _make_images(str(tmp_path))
    exe = _build_headless(tmp_path)
    csv = str(tmp_path / "ext.csv")
    res = subprocess.run([exe, "images", os.path.join(str(tmp_path), "images"), "--iterations", "3", "--csv", csv, "--hbm-peak", "6550",
                          "--synthetic", "75x75", "--synthetic", "427x240"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    rows = open(csv).read().strip().split("\n")
    assert rows[0] == EXT_HEADER
    assert len(rows) == 1 + 4 * (len(IMAGES) + 2)
    for r in rows[1:]:
        c = [x.strip() for x in r.split(",")]
        assert len(c) == 18 and c[11] in ("GRAYSCALE", "EDGE", "GAUSSIAN", "FUSED")
        assert float(c[10]) == 0.0 and int(c[12]) == 0 and float(c[13]) > 0 and float(c[14]) > 0 and int(c[17]) == 1
    res = subprocess.run([exe, "stream", "640x480", "--frames", "24", "--inflight", "3"], capture_output=True, text=True)
    assert res.returncode == 0 and "every frame equals" in res.stdout, res.stdout + res.stderr
