"""SfMeta / SfContours per-zone statistics: oracle restatements and the host mirror against the
golden vectors recorded from the unmodified reference (oracle/gen_golden.py::gen_meta), and the CUDA kernels against both."""
import numpy as np
import pytest

from camkifu_b200 import meta


def test_oracle_zone_means_match_reference(golden, oracle):
    g = golden("meta.npz")
    for k in range(2):
        rs, re, cs, ce = (int(v) for v in g["zm_region_%d" % k])
        z = oracle.meta_zone_means(g["zm_img_%d" % k], g["zm_mask_%d" % k], 19, rs, re, cs, ce)
        assert np.array_equal(z, g["zm_zones_%d" % k])
    assert 0 < g["zm_zones_0"][:, :, 0].sum() < 361            # both kinds of zones are exercised


def test_oracle_vote_and_foreground_match_reference(golden, oracle):
    g = golden("meta.npz")
    assert [tuple(r) for r in g["regions"]] == oracle.meta_subregions() == meta.subregions()
    for i in range(6):
        mv = oracle.meta_vote(g["vote_hist_%d" % i], g["vote_empty_%d" % i].astype(bool))
        assert np.array_equal(mv, g["vote_moves_%d" % i])
    rects = oracle.c_zone_rects(19)
    for t in range(g["fg_masks"].shape[0]):
        fg = np.unpackbits(g["fg_masks"][t])[:380 * 380].reshape(380, 380) * np.uint8(255)
        counts = np.array([[fg[a0:a1, b0:b1].sum() // 255 for (a0, b0, a1, b1) in rects[r]] for r in range(19)])
        for k, (rs, re, cs, ce) in enumerate(meta.subregions()):
            assert oracle.meta_check_foreground(fg, rs, re, cs, ce) == bool(g["fg_calm"][t, k])
            # the product's form: from the per-zone counts only
            assert meta.check_foreground(counts, rects, rs, re, cs, ce) == bool(g["fg_calm"][t, k])
    assert 0.2 < g["fg_calm"].mean() < 0.95


@pytest.fixture(scope="module")
def engine():
    from camkifu_b200.engine import StoneEngine
    return StoneEngine(19)


@pytest.mark.gpu
def test_gpu_zone_means(engine, golden, oracle):
    import torch
    g = golden("meta.npz")
    for k in range(2):
        rs, re, cs, ce = (int(v) for v in g["zm_region_%d" % k])
        z = engine.zone_means(torch.from_numpy(g["zm_img_%d" % k]).cuda(), torch.from_numpy(g["zm_mask_%d" % k]).cuda(),
                              rs, re, cs, ce)
        assert np.array_equal(z[0].cpu().numpy(), g["zm_zones_%d" % k])          # the unmodified reference's table
    # a batch with random masks, against the oracle
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (3, 380, 380, 3), dtype=np.uint8)
    masks = (rng.random((3, 380, 380)) < rng.random((3, 1, 1))).astype(np.uint8)
    masks[2, :200] = 0
    z = engine.zone_means(torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda(), 2, 17, 5, 19).cpu().numpy()
    for k in range(3):
        assert np.array_equal(z[k], oracle.meta_zone_means(imgs[k], masks[k], 19, 2, 17, 5, 19))


@pytest.mark.gpu
def test_gpu_history_vote_and_foreground(engine, golden, oracle):
    import torch
    g = golden("meta.npz")
    for i in range(6):
        mv = engine.history_vote(torch.from_numpy(g["vote_hist_%d" % i]).cuda(), torch.from_numpy(g["vote_empty_%d" % i]).cuda())
        assert np.array_equal(mv.cpu().numpy(), g["vote_moves_%d" % i])
    rects = oracle.c_zone_rects(19)
    fgs = np.stack([np.unpackbits(m)[:380 * 380].reshape(380, 380) * np.uint8(255) for m in g["fg_masks"]])
    counts = engine.zone_fg_counts(torch.from_numpy(fgs).cuda()).cpu().numpy()
    for t in range(fgs.shape[0]):
        for k, (rs, re, cs, ce) in enumerate(meta.subregions()):
            assert meta.check_foreground(counts[t], rects, rs, re, cs, ce) == bool(g["fg_calm"][t, k])


@pytest.mark.gpu
def test_gpu_find_stones_regions_equals_serial_calls(engine, oracle):
    """SfMeta's nine regions x n frames in one set of launches = nine find_stones calls per frame (labels are not part of
    the batched outputs; stones, trust, ratios, centres and compactness are), and = the oracle."""
    import cv2
    import torch
    from camkifu_b200 import synth
    from camkifu_b200.engine import rng_seed, rng_advance
    frames, M, truth, _ = synth.make_clip(41, 3, 360, 480)
    gob = np.stack([cv2.warpPerspective(f, M, (380, 380)) for f in frames])
    gob[2, 100:, :] = 250                                   # a bright frame: float32 sums past 2^24 in some regions
    regions = meta.subregions()
    st0 = rng_seed(11)
    states = [[rng_advance(st0, 9 * f + r) for r in range(9)] for f in range(3)]     # one RNG stream, region after region
    want = ("stones", "trusted", "ratios", "centers", "compactness")
    res = engine.find_stones_regions(torch.from_numpy(gob).cuda(), regions, states, want=want)
    for f in range(3):
        for r, (rs, re, cs, ce) in enumerate(regions):
            ref = oracle.c_find_stones(gob[f], states[f][r], 19, rs, re, cs, ce)
            assert np.array_equal(res["stones"][f, r].cpu().numpy(), ref["stones"])
            assert bool(res["trusted"][f, r]) == ref["trusted"]
            assert np.array_equal(res["ratios"][f, r].cpu().numpy(), ref["ratios"])
            assert np.array_equal(res["centers"][f, r].cpu().numpy(), ref["centers"])
            assert abs(float(res["compactness"][f, r]) - ref["compactness"]) <= 1e-9 * ref["compactness"]
    one = engine.find_stones(torch.from_numpy(gob[:1]).cuda(), [states[0][4]], *regions[4])
    assert np.array_equal(one["stones"][0].cpu().numpy(), res["stones"][0, 4].cpu().numpy())
