"""Pin the C restatement of the OpenCV arithmetic (oracle/ck_oracle.c) bit-exactly against cv2 in this image."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from camkifu_b200 import synth  # noqa: E402


def test_invert_and_warp_bit_exact(oracle):
    rng = np.random.default_rng(0)
    for t in range(9):
        H, W = [(480, 640), (1080, 1920), (2160, 3840)][t % 3]
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        c = synth.random_corners(rng, H, W, jitter=0.12 if t % 2 else 0.03)
        if t % 4 == 3:
            c += np.float32([[-300, -200]] * 4)                    # quad partly outside the frame
        S = (380, 260, 180)[t % 3] if t >= 6 else 380
        M = synth.board_homography(c, S)
        assert np.array_equal(cv2.invert(M)[1], oracle.c_invert3x3(M))
        assert np.array_equal(cv2.warpPerspective(src, M, (S, S)), oracle.c_warp(src, M, S))


def test_warp_wild_homographies_bit_exact(oracle):
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    for M in synth.wild_homographies(rng, 24):
        assert np.array_equal(cv2.warpPerspective(src, M, (380, 380)), oracle.c_warp(src, M, 380))


def test_accumulate_weighted_bit_exact(oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (380, 380, 3), dtype=np.uint8).astype(np.float32)
    b = a.copy()
    for _ in range(6):
        g = rng.integers(0, 256, (380, 380, 3), dtype=np.uint8)
        cv2.accumulateWeighted(g, a, 0.2)
        oracle.c_accumulate(g, b, 0.2)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("kind", ["u8_board", "f32_accu_cols", "u8_noise", "f32_noise"])
def test_kmeans_bit_exact_and_rng_carry(oracle, kind):
    rng = np.random.default_rng(2)
    if kind in ("u8_board", "f32_accu_cols"):
        fr, M, _, _ = synth.make_clip(3, 2, 240, 320)
        g = cv2.warpPerspective(fr[0], M, (380, 380))
        img = g.astype(np.float32)
        if kind == "f32_accu_cols":
            cv2.accumulateWeighted(cv2.warpPerspective(fr[1], M, (380, 380)), img, 0.2)
            img = img[:, 120:260]
    elif kind == "u8_noise":
        img = rng.integers(0, 256, (200, 200, 3), dtype=np.uint8).astype(np.float32)
    else:
        img = (rng.random((200, 200, 3)) * 255).astype(np.float32)
    px = np.ascontiguousarray(img).reshape(-1, 3)
    crit = (cv2.TERM_CRITERIA_EPS, 15, 3)
    cv2.setRNGSeed(9)
    st = oracle.rng_seed_state(9)
    for _ in range(2):                                             # the second call continues the RNG stream
        r, l, c = cv2.kmeans(px, 3, None, crit, 3, cv2.KMEANS_PP_CENTERS)
        comp, lab, cen, st, _ = oracle.c_kmeans(px, st)
        assert np.array_equal(l.ravel(), lab) and np.array_equal(c, cen) and r == comp


def test_refpath_matches_c_oracle(oracle):
    """The cv2-calling restatement (cpu_baseline) and the C restatement agree end to end."""
    fr, M, truth, _ = synth.make_clip(5, 1, 240, 320)
    ref = oracle.RefPath(19)
    g = ref.warp(fr[0], M)
    cv2.setRNGSeed(4)
    stones, trusted, centers, comp = ref.find_stones(g)
    res = oracle.c_find_stones(oracle.c_warp(fr[0], M, 380), oracle.rng_seed_state(4))
    assert trusted and res["trusted"]
    assert np.array_equal(stones, res["stones"]) and np.array_equal(centers, res["centers"])
    assert np.array_equal(stones, truth[0])


def test_cnn_oracle_vs_torch_and_f64(oracle):
    """No reference implementation of the net can run here (Keras/Theano absent) -> the forward arithmetic is pinned
    against an independent torch-CPU fp32 implementation and an fp64-accumulate run, within float tolerance."""
    from camkifu_b200 import weights
    fr, M, _, _ = synth.make_clip(6, 1, 240, 320)
    xs = oracle.c_nn_gather(cv2.warpPerspective(fr[0], M, (380, 380)))[:24]
    params = weights.glorot_params(seed=0, bias_scale=0.5)
    y32, z32 = oracle.c_cnn_forward(xs, params, want_logits=True)
    y64, z64 = oracle.c_cnn_forward(xs, params, acc64=True, want_logits=True)
    yt = oracle.torch_cnn(params)(xs)
    scale = np.abs(z64).max()
    assert np.abs(z32 - z64).max() <= 1e-4 * scale
    assert np.abs(y32 - y64).max() <= 1e-3 and np.abs(yt - y64).max() <= 1e-3
    assert np.array_equal(y32.argmax(1), y64.argmax(1)) and np.array_equal(yt.argmax(1), y64.argmax(1))


@pytest.mark.parametrize("rates", ["reference", "auto", "fast", "mixed"])
def test_mog2_bit_exact(oracle, rates):
    """cko_mog2_apply against cv2.createBackgroundSubtractorMOG2(detectShadows=False).apply — the call of
    StonesFinder._learn_bg (stonesfinder.py:171-176) — on a noisy scene with moving and appearing objects: every mask
    of the stream is bit-identical (mode creation, matching, re-sorting, pruning, the first-frame rate of 1/2)."""
    import cv2
    rng = np.random.default_rng(3)
    S, n = 96, 90
    bg = rng.integers(0, 256, (S, S, 3)).astype(np.int16)
    ref = cv2.createBackgroundSubtractorMOG2(detectShadows=False)
    mine = oracle.CMog2((S, S))
    fg_seen = 0
    for i in range(n):
        f = bg + rng.integers(-12, 13, (S, S, 3))
        if i > 20:
            x = (i * 3) % (S - 20)
            f[30:50, x:x + 20] = rng.integers(0, 256, 3)
        if 50 < i < 75:
            f[60:80, 10:40] += 60
        f = np.clip(f, 0, 255).astype(np.uint8)
        lr = {"reference": 0.01 if i < 50 else 0.005, "auto": -1.0, "fast": 0.2, "mixed": (-1.0, 0.01, 0.3, 0.0)[i % 4]}[rates]
        a, b = ref.apply(f, learningRate=lr), mine.apply(f, lr)
        assert np.array_equal(a, b), "frame %d: %d pixels differ" % (i, int((a != b).sum()))
        fg_seen += int((a > 0).sum()) if i > 0 else 0
    assert fg_seen > 10000      # the scene did produce foreground
