"""ctypes wrapper over oracle/librip_oracle.so -- CPU ORACLE, test infrastructure only.

Importers allowed by the build contract: tests/, __graft_entry__.smoke(), and bench.py's
`cpu_baseline` / `--impl reference` legs.  The product package never imports this module.
Each function names the reference file:line it follows in oracle/rip_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librip_oracle.so")

RGB, BGR = 0, 1


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "rip_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librip_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.rip_oracle_gray.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.rip_oracle_gauss_weights.argtypes = [C.c_int, C.c_float, f32p]
        L.rip_oracle_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u8p, C.c_int]
        L.rip_oracle_sobel.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int]
        L.rip_oracle_fused.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, u8p, u8p, C.c_int]
        L.rip_oracle_mae_ch0.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        L.rip_oracle_mae_ch0.restype = C.c_double
        L.rip_oracle_max_abs.argtypes = [u8p, u8p, C.c_long, C.POINTER(C.c_long)]
        L.rip_oracle_ocl_gray_rgba.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.rip_oracle_ocl_sobel_rgba.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.rip_oracle_ocl_blur_rgba.argtypes = [u8p, C.c_int, C.c_int, C.c_int, f32p, u8p]
        _lib = L
    return _lib


def _u8(a: np.ndarray):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _chk(rc: int, what: str):
    if rc != 0:
        raise ValueError(f"oracle {what} rejected its arguments (rc={rc})")


def gray(img: np.ndarray, order: int = RGB, threads: int = 1) -> np.ndarray:
    """img: (H, W, 3|4) u8 -> (H, W) u8."""
    img = np.ascontiguousarray(img)
    h, w, cn = img.shape
    out = np.empty((h, w), np.uint8)
    _chk(lib().rip_oracle_gray(_u8(img), w, h, cn, order, _u8(out), threads), "gray")
    return out


def gauss_weights(ksize: int, sigma: float) -> np.ndarray:
    out = np.empty((ksize, ksize), np.float32)
    _chk(lib().rip_oracle_gauss_weights(ksize, C.c_float(sigma), _f32(out)), "gauss_weights")
    return out


def blur(img: np.ndarray, ksize: int, sigma: float | None = None, weights: np.ndarray | None = None,
         threads: int = 1) -> np.ndarray:
    """img: (H, W) or (H, W, C) u8 -> same shape."""
    img = np.ascontiguousarray(img)
    if weights is None:
        weights = gauss_weights(ksize, sigma)
    weights = np.ascontiguousarray(weights, np.float32)
    h, w = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    _chk(lib().rip_oracle_blur(_u8(img), w, h, cn, ksize, _f32(weights), _u8(out), threads), "blur")
    return out


def sobel(g: np.ndarray, threads: int = 1) -> np.ndarray:
    g = np.ascontiguousarray(g)
    h, w = g.shape
    out = np.empty((h, w), np.uint8)
    _chk(lib().rip_oracle_sobel(_u8(g), w, h, _u8(out), threads), "sobel")
    return out


def fused(img: np.ndarray, ksize: int = 5, sigma: float = 1.0, order: int = RGB,
          weights: np.ndarray | None = None, threads: int = 1) -> np.ndarray:
    img = np.ascontiguousarray(img)
    h, w, cn = img.shape
    if weights is None:
        weights = gauss_weights(ksize, sigma)
    weights = np.ascontiguousarray(weights, np.float32)
    out = np.empty((h, w), np.uint8)
    _chk(lib().rip_oracle_fused(_u8(img), w, h, cn, order, ksize, _f32(weights), _u8(out), None, threads), "fused")
    return out


def mae_ch0(a: np.ndarray, b: np.ndarray) -> float:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape
    h, w = a.shape[:2]
    cn = 1 if a.ndim == 2 else a.shape[2]
    return float(lib().rip_oracle_mae_ch0(_u8(a), _u8(b), w, h, cn))


def max_abs(a: np.ndarray, b: np.ndarray) -> tuple[int, int]:
    a, b = np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(b).reshape(-1)
    assert a.size == b.size
    nd = C.c_long(0)
    mx = lib().rip_oracle_max_abs(_u8(a), _u8(b), a.size, C.byref(nd))
    return int(mx), int(nd.value)


# ---- OpenCL buffer-path semantics: only for replaying the published Error_MAE KATs ----
def ocl_gray_rgba(rgba: np.ndarray) -> np.ndarray:
    rgba = np.ascontiguousarray(rgba)
    h, w, _ = rgba.shape
    out = np.empty((h, w, 4), np.uint8)
    _chk(lib().rip_oracle_ocl_gray_rgba(_u8(rgba), w, h, _u8(out)), "ocl_gray")
    return out


def ocl_sobel_rgba(rgba: np.ndarray) -> np.ndarray:
    rgba = np.ascontiguousarray(rgba)
    h, w, _ = rgba.shape
    out = np.empty((h, w), np.uint8)
    _chk(lib().rip_oracle_ocl_sobel_rgba(_u8(rgba), w, h, _u8(out)), "ocl_sobel")
    return out


def ocl_blur_rgba(rgba: np.ndarray, ksize: int, weights: np.ndarray) -> np.ndarray:
    rgba = np.ascontiguousarray(rgba)
    weights = np.ascontiguousarray(weights, np.float32)
    h, w, _ = rgba.shape
    out = np.empty((h, w, 4), np.uint8)
    _chk(lib().rip_oracle_ocl_blur_rgba(_u8(rgba), w, h, ksize, _f32(weights), _u8(out)), "ocl_blur")
    return out


# ---- oracle/_ref: the REFERENCE's own weight generator, compiled from /root/reference (oracle/Makefile `ref`) ----
_REF_PATH = os.path.join(_HERE, "_ref", "librip_ref_weights.so")
_ref = None


def build_ref() -> str | None:
    """(Re)build oracle/_ref where the reference tree exists; elsewhere keep the prebuilt file.  Returns its path
    or None when there is neither."""
    if os.path.isdir("/root/reference/src/GaussianBlur"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _REF_PATH if os.path.exists(_REF_PATH) else None


def ref_gauss_weights(ksize: int, sigma: float) -> np.ndarray | None:
    """Controller::_GenerateGausianKernel of the reference itself (src/GaussianBlur/src/Controller.cpp:342-362,395-417),
    or None when oracle/_ref has not been built."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_PATH):
            return None
        _ref = C.CDLL(_REF_PATH)
        _ref.rip_ref_gauss_weights.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_float)]
    out = np.empty((ksize, ksize), np.float32)
    _chk(_ref.rip_ref_gauss_weights(ksize, C.c_float(sigma), _f32(out)), "ref_gauss_weights")
    return out
