"""K1 parity on the GPU: ckb_warp / ckb_accumulate against the oracle (bit-exact) and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from camkifu_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from camkifu_b200.engine import StoneEngine
    return StoneEngine(19)


def test_warp_golden_stream(engine, golden):
    g = golden("clustering_stream.npz")
    frames = torch.from_numpy(g["frames"]).cuda()
    out = engine.warp(frames, g["mtx"]).cpu().numpy()
    assert np.array_equal(out, g["goban"])  # what the reference's StonesFinder._doframe produced, bit for bit


@pytest.mark.parametrize("H,W", [(480, 640), (1080, 1920), (2160, 3840), (481, 643)])
def test_warp_bit_exact_vs_oracle(engine, oracle, H, W):
    rng = np.random.default_rng(H)
    n = 3
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    mats = []
    for k in range(n):
        c = synth.random_corners(rng, H, W, jitter=0.10)
        if k == 2:
            c -= np.float32([0.35 * W, 0.3 * H])   # quad partly outside the frame: constant-0 border taps
        mats.append(synth.board_homography(c, 380))
    out = engine.warp(torch.from_numpy(frames).cuda(), np.array(mats)).cpu().numpy()
    for k in range(n):
        assert np.array_equal(out[k], oracle.c_warp(frames[k], mats[k], 380)), "frame %d" % k
    # single shared homography and a non-contiguous (cropped) view with pitches
    out1 = engine.warp(torch.from_numpy(frames).cuda(), mats[0]).cpu().numpy()
    for k in range(n):
        assert np.array_equal(out1[k], oracle.c_warp(frames[k], mats[0], 380))


def test_warp_wild_homographies(engine, oracle):
    """Horizon inside the canonical square, singular matrices, saturating scales, random dense matrices: still the
    oracle's bits (which are cv2's: tests/test_oracle_vs_cv2.py::test_warp_wild_homographies_bit_exact)."""
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    mats = synth.wild_homographies(rng, 24)
    frames = torch.from_numpy(np.stack([src] * len(mats))).cuda()
    out = engine.warp(frames, np.array(mats)).cpu().numpy()
    for k, M in enumerate(mats):
        assert np.array_equal(out[k], oracle.c_warp(src, M, 380)), "homography %d" % k


def test_warp_pitched_view_and_many_frames(engine, oracle):
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (40, 300, 420, 3), dtype=np.uint8)
    view = torch.from_numpy(big).cuda()[:, 10:250, 20:340, :]   # 40 frames of 240x320 inside a pitched buffer
    c = synth.random_corners(rng, 240, 320)
    M = synth.board_homography(c, 380)
    out = engine.warp(view, M).cpu().numpy()
    for k in (0, 17, 31, 32, 39):                                # crosses the 32-frame launch chunk
        assert np.array_equal(out[k], oracle.c_warp(np.ascontiguousarray(big[k, 10:250, 20:340]), M, 380))


@pytest.mark.parametrize("gsize", [9, 13])
def test_warp_other_board_sizes(oracle, gsize):
    from camkifu_b200.engine import StoneEngine
    eng = StoneEngine(gsize)
    rng = np.random.default_rng(gsize)
    frame = rng.integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
    M = synth.board_homography(synth.random_corners(rng, 2160, 3840), 20 * gsize)
    out = eng.warp(torch.from_numpy(frame).cuda(), M).cpu().numpy()[0]
    assert np.array_equal(out, oracle.c_warp(frame, M, 20 * gsize))


def test_accumulate_stream_and_snapshots(engine, oracle, golden):
    g = golden("clustering_stream.npz")
    goban = torch.from_numpy(g["goban"]).cuda()
    accu = torch.empty((380, 380, 3), dtype=torch.float32, device="cuda")
    snaps = engine.accumulate(goban, accu, first=True, snap_every=3, snap_phase=0)
    assert np.array_equal(accu.cpu().numpy(), g["accu_last"])
    assert snaps.shape[0] == 3
    ref = np.empty((380, 380, 3), np.float32)
    k = 0
    for i in range(7):
        oracle.c_accumulate(g["goban"][i], ref, 0.2, first=(i == 0))
        if i == 1:
            assert np.array_equal(ref, g["accu_1"])
        if i % 3 == 0:
            assert np.array_equal(snaps[k].cpu().numpy(), ref)
            k += 1
    # continuing an existing state in two calls gives the same bits
    accu2 = torch.empty_like(accu)
    engine.accumulate(goban[:4], accu2, first=True)
    engine.accumulate(goban[4:], accu2, first=False)
    assert torch.equal(accu, accu2)
