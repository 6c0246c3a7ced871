"""N > 1 host logic on CPU: two `gloo` ranks shard a batch with the library's own partition
(rip_shard_frames / rip_band_rows -- the split rip_process_host and bench.py use), each rank computes its
part, the parts are gathered and must reassemble to the single-rank result; the per-rank timings reduce
with MAX as in bench.py.  No GPU is involved, so the oracle stands in for the device (this is a test)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, h, w, out_q):
    import torch
    import torch.distributed as dist
    import oracle as O
    import rip_b200 as rip

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = np.stack([np.random.default_rng(0xB200 + i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i in range(n_frames)])
        wts = O.gauss_weights(5, 1.0)
        # --- frame sharding (BASELINE config 4): contiguous blocks of whole frames, no exchange step ---
        first, count = rip.shard_frames(n_frames, world, rank)
        mine = np.stack([O.fused(frames[i], 5, weights=wts) for i in range(first, first + count)]) if count else np.zeros((0, h, w), np.uint8)
        padded = np.zeros((n_frames, h, w), np.uint8)
        padded[first:first + count] = mine
        t = torch.from_numpy(padded.astype(np.int32))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)          # test-only gather (blocks are disjoint)
        # --- row bands of one frame with a 3-row halo (BASELINE config 5) ---
        i0, ni, o0, no = rip.band_rows(h, world, rank, 3)
        band = O.fused_band(frames[0], i0, ni, o0, no, 5, wts) if hasattr(O, "fused_band") else O.fused(frames[0], 5, weights=wts)[o0:o0 + no]
        rows = np.zeros((h, w), np.int32)
        rows[o0:o0 + no] = band
        tb = torch.from_numpy(rows)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        # --- timing reduction as in bench.py: max over ranks ---
        ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
        if rank == 0:
            out_q.put((t.numpy().astype(np.uint8), tb.numpy().astype(np.uint8), float(ms.item()), (first, count), (i0, ni, o0, no)))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_reassemble_the_single_rank_result(oracle):
    import torch.multiprocessing as mp

    n_frames, h, w, world = 5, 37, 48, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got_band, ms, shard0, band0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = np.stack([np.random.default_rng(0xB200 + i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i in range(n_frames)])
    wts = oracle.gauss_weights(5, 1.0)
    want = np.stack([oracle.fused(f, 5, weights=wts) for f in frames])
    assert np.array_equal(got, want)
    assert np.array_equal(got_band, want[0])
    assert ms == 15.0
    assert shard0 == (0, 2) and band0 == (0, 21, 0, 18)


def test_partitions_cover_everything_exactly_once():
    import rip_b200 as rip

    for n, parts in [(256, 8), (32, 1), (5, 2), (3, 8), (7, 3), (0, 4)]:
        seen = []
        for i in range(parts):
            first, count = rip.shard_frames(n, parts, i)
            seen += list(range(first, first + count))
        assert seen == list(range(n)), (n, parts)
    for h, parts, halo in [(4320, 8, 3), (2160, 4, 1), (10, 3, 3), (8, 8, 3)]:
        rows = []
        for i in range(parts):
            i0, ni, o0, no = rip.band_rows(h, parts, i, halo)
            rows += list(range(o0, o0 + no))
            assert i0 == max(0, o0 - halo) and i0 + ni == min(h, o0 + no + halo)
        assert rows == list(range(h))
    with pytest.raises(rip.RipError):
        rip.band_rows(4, 8, 0, 3)


def test_reference_arm_prints_on_rank_zero_only():
    """bench.py --impl reference under a 2-rank launch: rank 0 alone runs and prints, rank 1 exits 0 silently."""
    import subprocess
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == ""
