"""K3 + K2 parity on the GPU: ckb_find_stones against the oracle (= cv2.kmeans semantics) and the reference's goldens."""
import numpy as np
import pytest
import torch

from camkifu_b200 import synth

pytestmark = pytest.mark.gpu

ALL = ("stones", "trusted", "ratios", "centers", "compactness", "labels")


@pytest.fixture(scope="module")
def engine():
    from camkifu_b200.engine import StoneEngine
    return StoneEngine(19)


def compare(res, k, ref, exact_centers=True):
    assert np.array_equal(res["labels"][k].cpu().numpy(), ref["labels"]), "k-means labels differ"
    c = res["centers"][k].cpu().numpy()
    if exact_centers:
        assert np.array_equal(c, ref["centers"])
    assert np.allclose(c, ref["centers"], rtol=1e-3, atol=0)            # the stated float tolerance
    assert abs(float(res["compactness"][k]) - ref["compactness"]) <= 1e-9 * ref["compactness"]
    assert np.array_equal(res["ratios"][k].cpu().numpy(), ref["ratios"])
    assert np.array_equal(res["stones"][k].cpu().numpy(), ref["stones"])
    assert bool(res["trusted"][k]) == ref["trusted"]


def test_golden_full_board(engine, golden, oracle):
    g = golden("clustering_full.npz")
    seeds = g["seeds"]
    imgs = torch.from_numpy(np.stack([g["goban_0"], g["goban_1"], g["goban_2"], g["goban_sparse"]])).cuda()
    states = [engine.L.ckb_rng_seed(int(s)) for s in seeds[:4]]
    res = engine.find_stones(imgs, states, want=ALL)
    for k in range(3):
        assert np.array_equal(res["ratios"][k].cpu().numpy(), g["ratios_%d" % k])     # reference's cluster_colors
        assert np.array_equal(res["centers"][k].cpu().numpy(), g["centers_%d" % k])
        assert np.array_equal(res["stones"][k].cpu().numpy(), g["stones_%d" % k])      # reference's find_stones
        assert bool(res["trusted"][k])
    assert bool(g["sparse_is_none"]) and not bool(res["trusted"][3])                    # reference returned None
    sub = engine.find_stones(imgs[:1], [engine.L.ckb_rng_seed(int(seeds[4]))], rs=6, re=13, cs=12, ce=19)
    if g["stones_region"][0, 0] != 255:
        assert np.array_equal(sub["stones"][0].cpu().numpy(), g["stones_region"])


def test_golden_stream_accu_path(engine, golden, oracle):
    """The reference's standalone SfClustering._find: float32 accu, rows 0..19, columns 6..13, every third frame."""
    g = golden("clustering_stream.npz")
    goban = torch.from_numpy(g["goban"]).cuda()
    accu = torch.empty((380, 380, 3), dtype=torch.float32, device="cuda")
    snaps = engine.accumulate(goban, accu, first=True, snap_every=3, snap_phase=0)
    states = [engine.L.ckb_rng_seed(int(g["seeds"][i])) for i in (0, 3, 6)]
    res = engine.find_stones(snaps, states, rs=0, re=19, cs=6, ce=13, want=ALL)
    for k in range(3):
        ref = oracle.c_find_stones(snaps[k].cpu().numpy(), states[k], 19, 0, 19, 6, 13)
        compare(res, k, ref)
    # replay bulk_update on the device stones and compare with the board the reference's controller ended with
    board = np.zeros((19, 19), np.uint8)
    for k in range(3):
        assert bool(res["trusted"][k])
        board = res["stones"][k].cpu().numpy().copy()   # all 361 are submitted; E removes, B/W (re)places
    assert np.array_equal(board, g["board"])


@pytest.mark.parametrize("kind", ["u8", "f32"])
def test_batch_vs_oracle_rng_carry(engine, oracle, kind):
    import cv2
    n = 6
    frames, M, truth, _ = synth.make_clip(77, n, 240, 320)
    goban = engine.warp(torch.from_numpy(frames).cuda(), M)
    if kind == "f32":
        accu = torch.empty((380, 380, 3), dtype=torch.float32, device="cuda")
        imgs = engine.accumulate(goban, accu, first=True, snap_every=1)
    else:
        imgs = goban
    # one RNG stream carried across the n calls, as the reference's process-global cv::theRNG() would be
    from camkifu_b200.engine import rng_seed, rng_advance
    st0 = rng_seed(5)
    states = [rng_advance(st0, k) for k in range(n)]
    res = engine.find_stones(imgs, states, want=ALL)
    host = imgs.cpu().numpy()
    st = st0
    cv2.setRNGSeed(5)
    for k in range(n):
        ref = oracle.c_find_stones(host[k], st)
        assert st == states[k]
        st = ref["rng_state"]
        compare(res, k, ref)
        if kind == "u8":
            assert np.array_equal(ref["stones"], truth[k])


def test_noise_and_degenerate_images(engine, oracle):
    rng = np.random.default_rng(9)
    imgs = np.zeros((4, 380, 380, 3), np.uint8)
    imgs[0] = rng.integers(0, 256, (380, 380, 3), dtype=np.uint8)            # many Lloyd iterations
    imgs[1][:, :190] = 200                                                    # two colours only: a k-means++ centre
    imgs[1][:, 190:] = 30                                                     #   duplicates -> empty-cluster repair
    imgs[2][:] = 128                                                          # constant image
    imgs[3] = rng.integers(0, 2, (380, 380, 1), dtype=np.uint8) * 255         # black / white salt and pepper
    states = [engine.L.ckb_rng_seed(100 + k) for k in range(4)]
    res = engine.find_stones(torch.from_numpy(imgs).cuda(), states, want=ALL)
    for k in range(4):
        ref = oracle.c_find_stones(imgs[k], states[k])
        compare(res, k, ref)


def test_float32_sums_past_2_24(engine, oracle):
    """uint8 images whose cluster sums leave the range where float32 holds integers exactly. OpenCV adds the members one
    by one in float32; the kernel takes exact integer sums up to 2^24, evaluates the rounding of sums in [2^24, 2^25)
    as a parity automaton in parallel, and walks serially only across 2^24 and towards 2^25. Bright images put most of
    the board into one cluster: (0) sums cross 2^24 a third of the way in, (1) they get within reach of 2^25, (2) nearly
    all pixels are 255: beyond 2^25. Centres, labels and compactness must still be OpenCV's bits."""
    rng = np.random.default_rng(21)
    imgs = np.zeros((3, 380, 380, 3), np.uint8)
    for k, (frac, lo) in enumerate(((0.62, 180), (0.86, 215), (0.985, 255))):
        bright = rng.random((380, 380)) < frac
        imgs[k] = np.where(bright[:, :, None], rng.integers(lo, 256, (380, 380, 3)),
                           np.where(rng.random((380, 380, 1)) < 0.5, rng.integers(0, 40, (380, 380, 3)),
                                    rng.integers(90, 130, (380, 380, 3)))).astype(np.uint8)
    states = [engine.L.ckb_rng_seed(300 + k) for k in range(3)]
    res = engine.find_stones(torch.from_numpy(imgs).cuda(), states, want=ALL)
    for k in range(3):
        ref = oracle.c_find_stones(imgs[k], states[k])
        compare(res, k, ref)
    assert res["centers"][0].max() > 200


@pytest.mark.parametrize("gsize", [9, 13])
def test_other_board_sizes(oracle, gsize):
    from camkifu_b200.engine import StoneEngine
    eng = StoneEngine(gsize)
    rng = np.random.default_rng(gsize)
    stones = synth.random_stones(rng, gsize)
    corners = synth.random_corners(rng, 480, 640)
    M = synth.board_homography(corners, 20 * gsize)
    frame = synth.render_frame(rng, 480, 640, stones, corners)
    goban = eng.warp(torch.from_numpy(frame).cuda(), M)
    st = eng.L.ckb_rng_seed(3)
    res = eng.find_stones(goban, [st], want=ALL)
    ref = oracle.c_find_stones(goban[0].cpu().numpy(), st, gsize)
    compare(res, 0, ref)
    assert np.array_equal(ref["stones"], stones)


def test_random_subregions_vs_oracle(engine, oracle):
    """SfMeta calls find_stones on sub-regions of 6-7 rows / columns (sf_meta.py:253): random regions, uint8 and float32
    images, against the oracle (labels, centres, ratios, stones, density verdict)."""
    rng = np.random.default_rng(33)
    frames, M, truth, _ = synth.make_clip(17, 2, 360, 480)
    import cv2
    gob = np.stack([cv2.warpPerspective(f, M, (380, 380)) for f in frames])
    f32 = (gob[0].astype(np.float32) * 0.8 + gob[1].astype(np.float32) * 0.2)
    for t in range(8):
        rs, cs = int(rng.integers(0, 13)), int(rng.integers(0, 13))
        re, ce = rs + int(rng.integers(4, 8)), cs + int(rng.integers(4, 8))
        re, ce = min(re, 19), min(ce, 19)
        st = engine.L.ckb_rng_seed(500 + t)
        if t % 2:
            img = torch.from_numpy(f32).cuda()[None]
            ref = oracle.c_find_stones(f32, st, 19, rs, re, cs, ce)
        else:
            img = torch.from_numpy(gob[t % 2:t % 2 + 1]).cuda()
            ref = oracle.c_find_stones(gob[t % 2], st, 19, rs, re, cs, ce)
        res = engine.find_stones(img, [st], rs, re, cs, ce, want=ALL)
        compare(res, 0, ref)


def test_tiny_regions_and_odd_batches(engine, oracle):
    """Clusters of 1 / 2 / 4 CTAs (regions of one to a few zones: the cluster size follows the pixel count), partial last
    chunks, and batch sizes around the launch geometry (1, 3, 130 frames)."""
    rng = np.random.default_rng(55)
    frames, M, truth, _ = synth.make_clip(19, 3, 360, 480)
    import cv2
    gob = np.stack([cv2.warpPerspective(f, M, (380, 380)) for f in frames])
    for t, (rs, re, cs, ce) in enumerate(((0, 1, 0, 1), (18, 19, 18, 19), (3, 4, 5, 8), (7, 9, 7, 9), (0, 3, 16, 19),
                                          (10, 19, 0, 2), (0, 19, 18, 19))):
        st = engine.L.ckb_rng_seed(700 + t)
        res = engine.find_stones(torch.from_numpy(gob[t % 3:t % 3 + 1]).cuda(), [st], rs, re, cs, ce, want=ALL)
        compare(res, 0, oracle.c_find_stones(gob[t % 3], st, 19, rs, re, cs, ce))
    # batch sizes: the same three images repeated; every frame must equal the oracle's result for its own RNG state
    for n in (1, 3, 130):
        imgs = torch.from_numpy(gob[np.arange(n) % 3]).cuda()
        states = [engine.L.ckb_rng_seed(900 + k % 5) for k in range(n)]
        res = engine.find_stones(imgs, states, want=("stones", "trusted", "centers"))
        refs = {}
        for k in range(n):
            key = (k % 3, k % 5)
            if key not in refs:
                refs[key] = oracle.c_find_stones(gob[k % 3], states[k])
            assert np.array_equal(res["stones"][k].cpu().numpy(), refs[key]["stones"])
            assert np.array_equal(res["centers"][k].cpu().numpy(), refs[key]["centers"])
            assert bool(res["trusted"][k]) == refs[key]["trusted"]
