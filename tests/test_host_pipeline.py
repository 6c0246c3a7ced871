"""The host-buffer pipeline of librip_cuda (rip_submit / rip_collect / rip_process_host): persistent device workers,
pinned staging for pageable callers, several jobs in flight, several host threads on one context -- all bit-exact
against the oracle.  Mirrors how the reference's frame loop would drive it (one call per camera frame,
/root/reference/src/RealtimeImageProcessing/RealtimeImageProcessing.cpp:325-418)."""
import threading

import numpy as np
import pytest

import rip_b200 as rip
from conftest import synth_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if rip.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device: librip_cuda has no CPU fallback")
    c = rip.Context(list(range(rip.device_count())))
    yield c
    c.close()


def test_frames_in_flight_complete_in_order_and_are_exact(ctx, oracle):
    w = rip.gauss_weights(5, 1.0)
    frames = [synth_frame("uniform" if i % 2 else "smooth", 360, 640, 7000 + i, 4) for i in range(9)]
    want = [oracle.fused(f[..., :3].copy(), 5, weights=w, threads=0) for f in frames]
    tickets, got = [], []
    for i, f in enumerate(frames):
        tickets.append(ctx.submit(f, rip.OP_FUSED, rip.FMT_RGBA8, ksize=5, weights=w, prof=(i == 0)))
        if len(tickets) == 3:
            got.append(tickets.pop(0).collect())
    got += [t.collect() for t in tickets]
    first, prof = got[0]
    assert prof[1] > prof[0] and prof[3] >= prof[2] and prof[5] >= prof[4]
    got[0] = first
    for i in range(9):
        assert np.array_equal(got[i], want[i]), f"frame {i}"


def test_pageable_and_pinned_callers_agree(ctx, oracle):
    w = rip.gauss_weights(5, 1.5)
    n, h, wd = 5, 270, 480
    pin_in, pin_out = rip.PinnedBuffer(n * h * wd * 3), rip.PinnedBuffer(n * h * wd)
    frames = pin_in.array.reshape(n, h, wd, 3)
    frames[:] = np.stack([synth_frame("uniform", h, wd, 7100 + i) for i in range(n)])
    out_pinned = ctx.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w, out=pin_out.array.reshape(n, h, wd)).copy()
    pageable = np.array(frames)   # a plain malloc'ed copy
    out_pageable = ctx.process(pageable, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w)
    mixed = ctx.process(pageable, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w, out=pin_out.array.reshape(n, h, wd))
    assert np.array_equal(out_pinned, out_pageable) and np.array_equal(out_pinned, mixed)
    for i in range(n):
        assert np.array_equal(out_pinned[i], oracle.fused(frames[i], 5, weights=w, threads=0))


def test_registered_caller_memory(ctx, oracle):
    import ctypes as C
    a = synth_frame("uniform", 512, 640, 7200, 4)
    L = rip.lib()
    rip.check(L.rip_host_register(a.ctypes.data, a.nbytes), "rip_host_register")
    try:
        got = ctx.process(a, rip.OP_GRAY, rip.FMT_RGBA8, gray_out=rip.GRAY_OUT_RGBA)
    finally:
        rip.check(L.rip_host_unregister(a.ctypes.data), "rip_host_unregister")
    g = oracle.gray(a, threads=0)
    assert np.array_equal(got[..., 0], g) and np.array_equal(got[..., 1], g) and np.all(got[..., 3] == 255)


def test_many_chunks_cycle_through_the_buffer_sets(ctx, oracle):
    """More chunks than buffer sets, pageable in and out: every set is reused while earlier results are still staged."""
    n, h, wd = 14, 1080, 1920   # 6.2 MB per RGB frame, 48 MiB chunks -> 7-frame chunks; RGBA -> 6
    frames = np.stack([synth_frame("uniform", h, wd, 7300 + (i % 3), 4) for i in range(n)])
    got = ctx.process(frames, rip.OP_EDGE, rip.FMT_RGBA8)
    want = [oracle.sobel(oracle.gray(frames[i], threads=0), threads=0) for i in range(3)]
    for i in range(n):
        assert np.array_equal(got[i], want[i % 3]), f"frame {i}"


def test_two_host_threads_share_one_context(ctx, oracle):
    w = rip.gauss_weights(5, 1.0)
    imgs = [np.stack([synth_frame("uniform", 200 + 8 * t, 320, 7400 + 10 * t + i) for i in range(4)]) for t in range(2)]
    res, errs = [None, None], []

    def work(t):
        try:
            for _ in range(5):
                res[t] = ctx.process(imgs[t], rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=w)
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for t in range(2):
        for i in range(4):
            assert np.array_equal(res[t][i], oracle.fused(imgs[t][i], 5, weights=w, threads=0))


def test_errors_surface_at_collect_or_submit(ctx):
    bad = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(rip.RipError):
        ctx.process(bad, rip.OP_GAUSSIAN, rip.FMT_RGB8, ksize=5, weights=rip.gauss_weights(5, 1.0))   # blur needs 1 or 4 channels
    with pytest.raises(rip.RipError):   # even kernel size: rejected by rip_submit itself
        ctx.submit(np.zeros((2, 8, 8, 3), np.uint8), rip.OP_FUSED, rip.FMT_RGB8, ksize=4, weights=np.ones((4, 4), np.float32))
    # the context still works afterwards
    out = ctx.process(np.zeros((16, 16, 3), np.uint8), rip.OP_EDGE, rip.FMT_RGB8)
    assert out.shape == (16, 16) and not out.any()


def test_tiny_frames_in_large_batches(ctx, oracle):
    """More than 65535 frames per chunk (gridDim.z of the tile kernels): slabs."""
    n = 70000
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, (n, 3, 5, 3), dtype=np.uint8)
    got = ctx.process(frames, rip.OP_EDGE, rip.FMT_RGB8)
    for i in (0, 1, 65534, 65535, 65536, n - 1):
        assert np.array_equal(got[i], oracle.sobel(oracle.gray(frames[i])))
