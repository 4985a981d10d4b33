"""N > 1 host logic on CPU: shard ranges, per-shard streaming state, and the final gather over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from camkifu_b200 import sharding


def test_shard_ranges_cover_and_align():
    for n in (0, 1, 2, 3, 10, 100, 1000, 100000):
        for world in (1, 2, 4, 8):
            got = [sharding.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            for (a, b), (c, d) in zip(got, got[1:]):
                assert b == c and a <= b
            assert all(a % 3 == 0 for a, b in got if a < n)
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 3 or n < 3 * world
    assert sharding.halo_start(0) == 0 and sharding.halo_start(300) == 180 and sharding.halo_start(300) % 3 == 0


def test_rng_state_replay(oracle):
    """The RNG state at frame f equals the state after the k-means calls of frames 0, 3, ..., < f (39 draws each)."""
    px = np.random.default_rng(0).integers(0, 256, (400, 3)).astype(np.float32)
    st = oracle.rng_seed_state(5)
    states = {}
    s = st
    for f in range(0, 13, 3):
        states[f] = s
        s = oracle.c_kmeans(px, s)[3]
    for f, want in states.items():
        assert sharding.rng_state_at(st, f) == want
    assert sharding.kmeans_calls_before(4) == 2 and sharding.kmeans_calls_before(6) == 2


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = sharding.shard_range(n, rank, world)
        frames = torch.arange(a, b, dtype=torch.int64)
        local = ((frames[:, None] * 7 + torch.arange(361)[None, :]) % 3).to(torch.uint8)   # fake board states
        full = sharding.gather_board_states(local, n)
        want = ((torch.arange(n)[:, None] * 7 + torch.arange(361)[None, :]) % 3).to(torch.uint8)
        out[rank] = bool(torch.equal(full, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 64])
def test_gather_board_states_gloo_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    assert out[0] and out[1]


class _FakePipeline:
    """Stands in for DetectPipeline on a host without a GPU: a deterministic function of each frame, so that the ingest,
    sharding and gather logic of process_video can run under gloo."""

    class eng:
        device = torch.device("cpu")

    def detect_stream(self, batches, depth=2):
        for frames, mtx, st in batches:
            f = frames.numpy().astype(np.int64)
            s = (f.reshape(f.shape[0], -1).sum(1)[:, None, None] + np.arange(361).reshape(19, 19)[None]) % 3
            yield {"stones": s.astype(np.uint8), "keep": (s > 0).astype(np.uint8),
                   "conf": (s / 2.0).astype(np.float32)}


def _video_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from camkifu_b200.video import process_video
        frames = np.random.default_rng(0).integers(0, 256, (29, 24, 32, 3), dtype=np.uint8)
        res = process_video(frames, np.eye(3), mode="neural", batch=4, pipeline=_FakePipeline())
        ref = next(_FakePipeline().detect_stream([(torch.from_numpy(frames), None, 0)]))
        out[rank] = all(np.array_equal(res[k], ref[k]) for k in ref) and res["stones"].shape == (29, 19, 19)
    finally:
        dist.destroy_process_group()


def test_process_video_shards_and_gathers_gloo_world2():
    """process_video under torch.distributed (gloo, 2 ranks, CPU): each rank decodes and processes its own frame range
    (starts aligned to 3) and both end up with the whole video's board states in frame order."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_video_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


class _ShortClip:
    """An indexable that claims more frames than it has (what FFmpeg's CAP_PROP_FRAME_COUNT does for many containers)."""

    def __init__(self, frames, claimed):
        self.frames, self.claimed = frames, claimed

    def __len__(self):
        return self.claimed

    def __getitem__(self, i):
        if i >= len(self.frames):
            raise IndexError(i)
        return self.frames[i]


def _short_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from camkifu_b200.video import process_video
        frames = np.random.default_rng(1).integers(0, 256, (22, 24, 32, 3), dtype=np.uint8)
        stats = {}
        res = process_video(_ShortClip(frames, 30), np.eye(3), mode="neural", batch=4, pipeline=_FakePipeline(), stats=stats)
        ref = next(_FakePipeline().detect_stream([(torch.from_numpy(frames), None, 0)]))
        out[rank] = (all(np.array_equal(res[k], ref[k]) for k in ref), stats["frames"])
    finally:
        dist.destroy_process_group()


def test_short_video_does_not_stall_the_gather_gloo_world2():
    """The header promises 30 frames, the stream ends after 22: the second rank gets 7 of its 15 frames, no rank raises
    or blocks, and both return the 22 frames that exist, in order."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_short_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] + out[1][1] == 22


def test_process_video_without_process_group_ignores_gather():
    """rank / world passed explicitly but torch.distributed not initialised: the rank's own range comes back, no error."""
    from camkifu_b200.video import process_video
    frames = np.random.default_rng(2).integers(0, 256, (20, 24, 32, 3), dtype=np.uint8)
    res = process_video(frames, np.eye(3), mode="neural", batch=4, pipeline=_FakePipeline(), rank=1, world=2)
    a, b = sharding.shard_range(20, 1, 2)
    ref = next(_FakePipeline().detect_stream([(torch.from_numpy(frames[a:b]), None, 0)]))
    assert all(np.array_equal(res[k], ref[k]) for k in ref)


def test_running_average_halo_reproduces_the_stream(oracle):
    """sharding.halo_start: a shard that feeds its running average from ACCU_HALO frames before its start ends up with the
    accumulator of the uninterrupted stream, bit for bit, on noisy frames (a convergence property: see sharding.py)."""
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (260, 16, 16, 3), dtype=np.uint8)
    start = 201
    stream = np.zeros((16, 16, 3), np.float32)
    for i in range(start + 1):
        oracle.c_accumulate(frames[i], stream, first=(i == 0))
    h0 = sharding.halo_start(start)
    assert h0 % 3 == 0 and start - h0 >= sharding.ACCU_HALO
    shard = np.zeros((16, 16, 3), np.float32)
    for i in range(h0, start + 1):
        oracle.c_accumulate(frames[i], shard, first=(i == h0))
    assert np.array_equal(shard, stream)
