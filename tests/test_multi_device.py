"""In-process multi-device paths of librip_cuda: one rip_ctx over EVERY visible device.

* frame-batch sharding: rip_process_host cuts a batch into contiguous blocks of frames, one per device, each with its
  own worker thread, pinned staging and streams; no device-to-device traffic.
* row bands: rip_process_host_banded splits the output rows of ONE large frame over the devices and uploads every band
  with its halo rows (config 5: 7680x4320).

On the single-GPU test box the context degrades to one device (the same code path with one worker); under
`gpurun --gpus N` the same tests exercise N devices -- logs of those runs are kept in profiles/.
Every result is compared bit for bit with the CPU oracle."""
import numpy as np
import pytest

import rip_b200 as rip
from conftest import synth_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def all_ctx():
    n = rip.device_count()
    if n < 1:
        pytest.fail("GPU tests need a CUDA device: librip_cuda has no CPU fallback")
    c = rip.Context(list(range(n)))
    yield c
    c.close()


def _eq(got, want, what):
    if not np.array_equal(got, want):
        d = np.argwhere(got != want)
        raise AssertionError(f"{what}: {len(d)} bytes differ, first at {d[0].tolist()}")


def test_context_spans_every_visible_device(all_ctx):
    assert all_ctx.devices == list(range(rip.device_count()))
    print(f"[multi-device] context over {len(all_ctx.devices)} device(s)")


@pytest.mark.parametrize("n_frames", [1, 3, 8, 19])
def test_frame_shards_over_all_devices_are_exact(all_ctx, oracle, n_frames):
    """Batches that do and do not divide by the device count, fewer frames than devices included."""
    h, w = 270, 480
    frames = np.stack([synth_frame("uniform" if i % 3 else "smooth", h, w, 900 + i) for i in range(n_frames)])
    wts = rip.gauss_weights(5, 1.0)
    fused = all_ctx.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=wts)
    edge = all_ctx.process(frames, rip.OP_EDGE, rip.FMT_RGB8)
    gray = all_ctx.process(frames, rip.OP_GRAY, rip.FMT_RGB8)
    for i in range(n_frames):
        g = oracle.gray(frames[i], threads=0)
        _eq(gray[i], g, f"gray frame {i}")
        _eq(edge[i], oracle.sobel(g, threads=0), f"edge frame {i}")
        _eq(fused[i], oracle.fused(frames[i], 5, weights=wts, threads=0), f"fused frame {i}")


def test_frame_shards_rgba_blur_over_all_devices(all_ctx, oracle):
    frames = np.stack([synth_frame("uniform", 96, 200, 950 + i, 4) for i in range(11)])
    wts = rip.gauss_weights(5, 1.5)
    got = all_ctx.process(frames, rip.OP_GAUSSIAN, rip.FMT_RGBA8, ksize=5, weights=wts)
    for i in range(frames.shape[0]):
        _eq(got[i], oracle.blur(frames[i], 5, weights=wts, threads=0), f"blur frame {i}")


def test_4k_batch_sharded_over_all_devices_equals_one_device(all_ctx, oracle):
    """BASELINE config 4 shape (fewer frames): the N-device result is the 1-device result, and frame 0 / the last frame
    match the oracle."""
    n = max(4, 2 * len(all_ctx.devices))
    frames = np.stack([synth_frame("uniform", 2160, 3840, 4000 + i) for i in range(n)])
    wts = rip.gauss_weights(5, 1.0)
    got = all_ctx.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=wts)
    one = rip.Context([0])
    ref = one.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=wts)
    one.close()
    _eq(got, ref, "N-device vs 1-device")
    for i in (0, n - 1):
        _eq(got[i], oracle.fused(frames[i], 5, weights=wts, threads=0), f"4K frame {i}")


@pytest.mark.parametrize("op", ["fused", "edge"])
def test_8k_row_bands_one_band_per_device(all_ctx, oracle, op):
    """BASELINE config 5: one 7680x4320 frame, output rows split over the devices, each band uploaded with its halo."""
    frame = synth_frame("uniform", 4320, 7680, 5000)
    wts = rip.gauss_weights(5, 1.0)
    if op == "fused":
        got = all_ctx.process(frame, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=wts, banded=True)
        want = oracle.fused(frame, 5, weights=wts, threads=0)
    else:
        got = all_ctx.process(frame, rip.OP_EDGE, rip.FMT_RGB8, banded=True)
        want = oracle.sobel(oracle.gray(frame, threads=0), threads=0)
    _eq(got, want, f"8K banded {op} over {len(all_ctx.devices)} device(s)")


def test_row_bands_with_more_devices_than_make_sense(oracle):
    """A context that lists the same device several times behaves like that many devices (bands and shards)."""
    n = rip.device_count()
    devs = [i % n for i in range(4)]
    c = rip.Context(devs)
    frame = synth_frame("smooth", 203, 368, 5100)
    wts = rip.gauss_weights(5, 1.5)
    _eq(c.process(frame, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=wts, banded=True),
        oracle.fused(frame, 5, weights=wts, threads=0), "banded, 4 logical devices")
    frames = np.stack([synth_frame("uniform", 64, 128, 5200 + i) for i in range(6)])
    got = c.process(frames, rip.OP_FUSED, rip.FMT_RGB8, ksize=5, weights=wts)
    for i in range(6):
        _eq(got[i], oracle.fused(frames[i], 5, weights=wts, threads=0), f"sharded, 4 logical devices, frame {i}")
    c.close()
