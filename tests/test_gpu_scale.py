"""Size-independent properties at BASELINE.json's full sizes (1080p frames, 64-frame batches and beyond), where the
oracle would take minutes: batches equal their parts, results do not depend on batch composition or order, ragged
batch sizes cross the library's internal 64-frame passes, empty inputs are no-ops."""
import numpy as np
import pytest
import torch

from camkifu_b200 import synth, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from camkifu_b200.engine import StoneEngine
    e = StoneEngine(19)
    e.set_cnn_weights(weights.glorot_params(seed=0))
    return e


@pytest.fixture(scope="module")
def clip():
    frames, mtx, truth, _ = synth.make_clip_parallel(77, 24, 1080, 1920)
    return frames, mtx, truth


def test_warp_batch_equals_singles_and_per_frame_homographies(engine, clip, oracle):
    frames, mtx, _ = clip
    d = torch.from_numpy(frames).cuda()
    whole = engine.warp(d, mtx)
    assert np.array_equal(whole[5].cpu().numpy(), oracle.c_warp(frames[5], mtx, 380))       # one full-size oracle check
    for i in (0, 11, 23):
        assert torch.equal(engine.warp(d[i], mtx)[0], whole[i])
    # one homography per frame (n_mtx = n): frame i under a slightly different board position
    rng = np.random.default_rng(0)
    mtxs = np.stack([mtx @ np.array([[1, 0, rng.normal(0, 3)], [0, 1, rng.normal(0, 3)], [0, 0, 1]]) for _ in range(24)])
    per = engine.warp(d, mtxs)
    for i in (3, 17):
        assert torch.equal(engine.warp(d[i], mtxs[i])[0], per[i])
    assert engine.warp(d[:0], mtx).shape[0] == 0


def test_cnn_ragged_batches_cross_the_64_frame_pass(engine, clip):
    """130 frames = two internal passes of 64 + a tail of 2: identical to the frames evaluated in other groupings, and
    equivariant under a permutation of the batch."""
    frames, mtx, _ = clip
    goban = engine.warp(torch.from_numpy(frames).cuda(), mtx)
    big = goban.repeat(6, 1, 1, 1)[:130].contiguous()
    out = engine.cnn_forward(big)
    ref = engine.cnn_forward(goban)
    for k in ("stones", "conf", "keep", "softmax"):
        for j in range(130):
            assert torch.equal(out[k][j], ref[k][j % 24]), (k, j)
    perm = torch.randperm(130, generator=torch.Generator().manual_seed(1)).cuda()
    outp = engine.cnn_forward(big[perm].contiguous())
    assert torch.equal(outp["softmax"], out["softmax"][perm])
    one = engine.cnn_forward(goban[7:8])
    assert torch.equal(one["softmax"][0], ref["softmax"][7])
    assert engine.cnn_forward(goban[:0])["stones"].shape[0] == 0


def test_find_stones_batch_equals_singles_and_truth(engine, clip):
    from camkifu_b200.engine import rng_seed, rng_advance
    frames, mtx, truth = clip
    goban = engine.warp(torch.from_numpy(frames).cuda(), mtx)
    st0 = rng_seed(5)
    states = [rng_advance(st0, i) for i in range(24)]
    res = engine.find_stones(goban, states, want=("stones", "trusted", "centers", "compactness", "labels"))
    assert np.array_equal(res["stones"].cpu().numpy(), truth)       # the synthetic boards are read perfectly
    assert bool(res["trusted"].all())
    for i in (0, 9, 23):
        one = engine.find_stones(goban[i:i + 1], [states[i]], want=("stones", "centers", "compactness", "labels"))
        for k in ("stones", "centers", "compactness", "labels"):
            assert torch.equal(one[k][0], res[k][i]), (k, i)
    # 200 frames take the wide-batch launch configuration: same bits
    big = goban.repeat(9, 1, 1, 1)[:200].contiguous()
    res2 = engine.find_stones(big, [states[i % 24] for i in range(200)], want=("stones", "centers"))
    for j in range(200):
        assert torch.equal(res2["centers"][j], res["centers"][j % 24]) and torch.equal(res2["stones"][j], res["stones"][j % 24])


def test_pipeline_both_modes_1080p(engine, clip):
    """DetectPipeline ("both": k-means + CNN) on 1080p host frames, ragged sub-batches, against the engine calls."""
    from camkifu_b200.engine import rng_seed, rng_advance
    from camkifu_b200.pipeline import DetectPipeline, pinned_frames
    frames, mtx, truth = clip
    pipe = DetectPipeline(1080, 1920, mode="both", sub_batch=10, engine=engine)
    host = pinned_frames(24, 1080, 1920)
    host.copy_(torch.from_numpy(frames))
    st0 = rng_seed(2)
    res = {k: v.copy() for k, v in pipe.detect(host, mtx, rng_state=st0).items()}
    goban = engine.warp(torch.from_numpy(frames).cuda(), mtx)
    nn = engine.cnn_forward(goban)
    km = engine.find_stones(goban, [rng_advance(st0, i) for i in range(24)])
    assert np.array_equal(res["stones"], nn["stones"].cpu().numpy())
    assert np.array_equal(res["conf"], nn["conf"].cpu().numpy())
    assert np.array_equal(res["km_stones"], km["stones"].cpu().numpy())
    assert np.array_equal(res["km_stones"], truth)
    assert pipe.h2d_bytes < 0.6 * frames.nbytes          # only the board's bounding box crossed PCIe


def test_error_codes(engine):
    """The C ABI reports misuse through status codes and messages, never by crashing."""
    from camkifu_b200._lib import CkbError
    from camkifu_b200.engine import StoneEngine
    g = torch.zeros((1, 380, 380, 3), dtype=torch.uint8, device="cuda")
    fresh = StoneEngine(19)
    with pytest.raises(CkbError, match="ckb_set_cnn_weights"):
        fresh.cnn_forward(g)                             # no weights yet
    with pytest.raises(CkbError):
        fresh.set_cnn_weights(np.zeros(10, np.float32))  # wrong parameter count
    bad = weights.glorot_params(seed=0).copy()
    bad[5] = np.nan
    with pytest.raises(CkbError, match="not finite"):
        fresh.set_cnn_weights(bad)
    small = StoneEngine(9)
    with pytest.raises(CkbError, match="19x19"):
        small.set_cnn_weights(weights.glorot_params(seed=0))     # the network is defined for 19x19 only
    # a singular homography is not an error: cv::invert yields the zero matrix and every tap lands on pixel (0, 0)
    src = torch.full((1, 8, 8, 3), 9, dtype=torch.uint8, device="cuda")
    assert int(engine.warp(src, np.zeros((3, 3))).min()) == 9
    L = engine.L
    assert L.ckb_cnn_forward(engine._h, None, 1, None, 0, None, None, None, None, None) != 0
    assert b"bad argument" in L.ckb_last_error(engine._h)


@pytest.mark.parametrize("gsize", [9, 13, 19])
def test_config4_4k_mixed_board_sizes(oracle, gsize):
    """BASELINE config 4: 3840x2160 frames, boards of 9 / 13 / 19 lines under perspective jitter — the warped pixels, the
    k-means labels and the board state are bit-identical to the oracle's (one oracle run per board size, as the
    reference needs one process per gsize)."""
    from camkifu_b200.engine import StoneEngine, rng_seed, rng_advance
    S = 20 * gsize
    eng = StoneEngine(gsize)
    frames, mtx, truth, _ = synth.make_clip_parallel(40 + gsize, 3, 2160, 3840, gsize=gsize)
    goban = eng.warp(torch.from_numpy(frames).cuda(), mtx)
    st0 = rng_seed(gsize)
    states = [rng_advance(st0, i) for i in range(3)]
    res = eng.find_stones(goban, states, want=("stones", "trusted", "labels", "centers"))
    for i in range(3):
        g_ref = oracle.c_warp(frames[i], mtx, S)
        assert np.array_equal(goban[i].cpu().numpy(), g_ref)
        ref = oracle.c_find_stones(g_ref, states[i], gsize, 0, gsize, 0, gsize)
        assert np.array_equal(res["labels"][i].cpu().numpy(), ref["labels"])
        assert np.array_equal(res["centers"][i].cpu().numpy(), ref["centers"])
        assert np.array_equal(res["stones"][i].cpu().numpy(), ref["stones"])
    assert np.array_equal(res["stones"].cpu().numpy(), truth)


def test_pipeline_full_mode_stream_matches_oracle(engine, oracle):
    """DetectPipeline(mode="full") — everything a finder runs per frame, statistics branch and CNN branch on two streams —
    over three batches of a small 'game' clip: per-zone foreground counts equal the oracle's MOG2 (streamed across the
    batches, the reference's learning-rate schedule), k-means board states equal the oracle's find_stones per frame, CNN
    board states equal the single-stream engine call."""
    from camkifu_b200 import synth, weights
    from camkifu_b200.engine import rng_seed, rng_advance
    from camkifu_b200.pipeline import DetectPipeline
    import cv2
    H, W, n = 240, 320, 60
    frames, mtx, truth, _ = synth.make_game_clip(9, n, H, W, events=[(20, 1, 3, 4), (40, 2, 10, 12)])
    engine.set_cnn_weights(weights.glorot_params(seed=0))
    pipe = DetectPipeline(H, W, mode="full", sub_batch=8, engine=engine)
    st0 = rng_seed(4)
    got = {}
    for b0 in range(0, n, 20):
        res = pipe.detect(torch.from_numpy(frames[b0:b0 + 20]), mtx, rng_state=rng_advance(st0, b0))
        for k, v in res.items():
            got.setdefault(k, []).append(v.copy())
    got = {k: np.concatenate(v) for k, v in got.items()}
    model = oracle.CMog2((380, 380))
    rects = oracle.c_zone_rects(19)
    for i in range(n):
        g = cv2.warpPerspective(frames[i], mtx, (380, 380))
        fg = model.apply(g, 0.01 if i < 50 else 0.005)
        counts = np.array([[int(fg[a0:a1, b0:b1].sum()) // 255 for (a0, b0, a1, b1) in rects[r]] for r in range(19)])
        assert np.array_equal(got["fg_counts"][i], counts), "frame %d" % i
        if i % 7 == 0:
            ref = oracle.c_find_stones(g, rng_advance(st0, i))
            assert np.array_equal(got["km_stones"][i], ref["stones"]) and bool(got["km_trusted"][i]) == ref["trusted"]
    gob = engine.warp(torch.from_numpy(frames[:20]).cuda(), mtx)
    nn = engine.cnn_forward(gob, want_softmax=False)
    assert np.array_equal(got["stones"][:20], nn["stones"].cpu().numpy())
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0 and pipe.frames_seen == n
