"""The oracle (oracle/ck_oracle.c + oracle/oracle.py) against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py) and the reference's own known-answer tests (test/camkifu/stone/test_tmanager.py:18-27)."""
import numpy as np
import pytest


@pytest.mark.parametrize("gsize", [9, 13, 19])
def test_geometry_tables(golden, oracle, gsize):
    g = golden("geometry_g%d.npz" % gsize)
    rects = oracle.c_zone_rects(gsize)
    assert np.array_equal(rects, g["rects"])                       # StonesFinder.getrect, stonesfinder.py:412-450
    mask, area = oracle.c_zone_mask(gsize)
    assert np.array_equal(mask * g["cover"], g["mask"])            # getmask, stonesfinder.py:452-493
    assert area == int(g["zone_area"]) == 315
    if gsize == 19:                                                # SURVEY A.5 probes
        assert tuple(rects[18, 18]) == (360, 360, 379, 379) and tuple(rects[18, 3]) == (360, 60, 379, 80)


def test_stream_warp_accu_and_moves(golden, oracle):
    """StonesFinder._doframe -> SfClustering._find over 7 frames (stonesfinder.py:123-152, sf_clustering.py:23-46)."""
    g = golden("clustering_stream.npz")
    frames, mtx = g["frames"], g["mtx"]
    accu = np.empty((380, 380, 3), np.float32)
    board = np.zeros((19, 19), np.uint8)
    for i in range(frames.shape[0]):
        goban = oracle.c_warp(frames[i], mtx, 380)
        assert np.array_equal(goban, g["goban"][i]), "warp differs at frame %d" % i
        oracle.c_accumulate(goban, accu, 0.2, first=(i == 0))
        if i == 1:
            assert np.array_equal(accu, g["accu_1"])
        if i % 3 == 0:
            res = oracle.c_find_stones(accu, oracle.rng_seed_state(int(g["seeds"][i])), 19, 0, 19, 6, 13)
            assert res["trusted"]
            # bulk_update semantics (stonesfinder.py:286-321): E on a non-empty spot removes, B/W on a different
            # colour emits a delete then the new colour
            moves = []
            for r in range(19):
                for c in range(19):
                    s = res["stones"][r, c]
                    if s == 0 and board[r, c] != 0:
                        moves.append((0, r, c))
                    elif s != 0:
                        if board[r, c] != 0:
                            if board[r, c] == s:
                                continue
                            moves.append((0, r, c))
                        moves.append((s, r, c))
            for s, r, c in moves:
                board[r, c] = s
            assert np.array_equal(np.array(moves, np.int32).reshape(-1, 3), g["moves_%d" % i])
        else:
            assert g["moves_%d" % i].shape[0] == 0
    assert np.array_equal(accu, g["accu_last"])
    assert np.array_equal(board, g["board"])


def test_full_board_find_stones(golden, oracle):
    g = golden("clustering_full.npz")
    seeds = g["seeds"]
    for k in range(3):
        res = oracle.c_find_stones(g["goban_%d" % k], oracle.rng_seed_state(int(seeds[k])))
        assert np.array_equal(res["ratios"], g["ratios_%d" % k])
        assert np.array_equal(res["centers"], g["centers_%d" % k])
        assert res["trusted"] and np.array_equal(res["stones"], g["stones_%d" % k])
        assert np.array_equal(res["stones"], g["truth_%d" % k])    # and the synthetic position is recovered
    res = oracle.c_find_stones(g["goban_sparse"], oracle.rng_seed_state(int(seeds[3])))
    assert bool(g["sparse_is_none"]) and not res["trusted"]       # check_density -> None
    res = oracle.c_find_stones(g["goban_0"], oracle.rng_seed_state(int(seeds[4])), 19, 6, 13, 12, 19)
    expect = g["stones_region"]
    if expect[0, 0] == 255:
        assert not res["trusted"]
    else:
        assert res["trusted"] and np.array_equal(res["stones"], expect)


def test_neural_geometry_and_codec(golden, oracle):
    import ctypes as C
    g = golden("neural_geometry.npz")
    assert (int(g["split"]), int(g["step"]), int(g["nb_classes"]), int(g["r_width"]), int(g["c_width"])) == (10, 2, 81, 40, 40)
    for i in range(10):
        for j in range(10):
            x0, y0 = C.c_int(), C.c_int()
            oracle.lib().cko_nn_patch_origin(i, j, C.byref(x0), C.byref(y0))
            assert (x0.value, x0.value + 40, y0.value, y0.value + 40) == tuple(g["rect_nn"][i, j])
    assert np.array_equal(oracle.c_nn_gather(g["goban"]), g["xs"])  # NNManager.generate_xs
    for k in range(81):
        assert np.array_equal(oracle.c_compute_stones(k), g["stones_of_label"][k])
    # the reference's own known-answer tests, test/camkifu/stone/test_tmanager.py:18-27
    E, B, W = 0, 1, 2
    assert list(oracle.c_compute_stones(27)) == [E, E, E, B]
    assert list(oracle.c_compute_stones(36)) == [E, E, B, B]
    assert list(oracle.c_compute_stones(64)) == [B, E, B, W]
    sol = np.array([oracle.c_compute_stones(k) for k in range(81)])
    assert list(np.where(sol[:, 3] == W)[0]) == list(range(54, 81))
    assert list(np.where(sol[:, 0] == W)[0]) == list(range(2, 81, 3))
    assert np.array_equal(np.array([[np.where(sol[:, d] == col)[0] for col in range(3)] for d in range(4)]),
                          g["class_indices"])
    assert int(g["label_allB"]) == 40


def test_neural_decode(golden, oracle):
    """NNCache.predict_all_stones + SfNeural.predict_all (nn_cache.py:25-41, sf_neural.py:57-70)."""
    g = golden("neural_decode.npz")
    stones, conf, keep = oracle.c_nn_decode(g["y"])
    assert np.array_equal(stones, g["stones"])
    assert np.array_equal(conf, g["conf"])
    moves = np.array([(stones[r, c], r, c) for r in range(19) for c in range(19) if keep[r, c]], np.int32)
    assert np.array_equal(moves.reshape(-1, 3), g["moves"])


def test_background_stream(golden, oracle):
    """background_stream.npz: the reference's StonesFinder (learn_bg=True) driven through _doframe. The oracle's warp +
    MOG2 restatement reproduces every foreground mask and the per-zone sums SfNeural.is_agitated thresholds."""
    g = golden("background_stream.npz")
    frames, mtx, init = g["frames"], g["mtx"], int(g["bg_init_frames"])
    model = oracle.CMog2((380, 380))
    rects = oracle.c_zone_rects(19)
    for i in range(frames.shape[0]):
        goban = oracle.c_warp(frames[i], mtx, 380)
        fg = model.apply(goban, 0.01 if i < init else 0.005)
        assert np.array_equal(np.packbits(fg > 0), g["masks"][i]), "frame %d" % i
        zones = np.array([[int((fg[a0:a1, b0:b1] > 0).sum()) for (a0, b0, a1, b1) in rects[r]] for r in range(19)])
        assert np.array_equal(zones, g["zone_fg"][i])
    assert g["zone_fg"][12].max() > 200     # the hand covers whole zones mid-clip
