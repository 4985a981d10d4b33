"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE from /root/reference (authoring container only).

TEST INFRASTRUCTURE ONLY. Usage:  python -m oracle.gen_golden            (from the repo root)

The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs are committed as small fixtures:

  geometry_g{9,13,19}.npz  StonesFinder.getrect / getmask / zone_area / PosGrid.mtx   (stonesfinder.py:412-493,964-981)
                           one process per board size: gsize is an import-time constant of the reference
  clustering_stream.npz    SfClustering driven through StonesFinder._doframe for 7 frames of a 320x240 clip:
                           goban_img, accu, piped bulk moves (stonesfinder.py:123-152, sf_clustering.py:23-46)
  clustering_full.npz      SfClustering.find_stones / cluster_colors on whole-board uint8 images (sf_clustering.py:48-129)
  neural_geometry.npz      NNManager geometry + codec + class_indices (nn_manager.py:92-126,216-275,360-382) and the
                           reference's own known-answer tests (test/camkifu/stone/test_tmanager.py:18-27)
  neural_decode.npz        NNCache.predict_all_stones + SfNeural.predict_all with a scripted net (nn_cache.py:25-52,
                           sf_neural.py:57-70)
  background_stream.npz    a StonesFinder with learn_bg=True driven through _doframe for 30 frames of a 160x120 "game"
                           clip (a hand places two stones): the MOG2 foreground masks of _learn_bg
                           (stonesfinder.py:113-115,171-176) and the per-zone foreground sums is_agitated reduces them to
                           (sf_neural.py:178-180)
  meta.npz                 SfMeta / SfContours per-zone statistics: the zone table of the unmodified
                           SfContours.find_stones with the mask it was computed from, Region.commit's votes on random
                           histories, Region.check_foreground of the 3 x 3 regions on random masks (see gen_meta)
  neural_stream.npz        the unmodified SfNeural._find over 72 frames of such a clip (background sampling, initial
                           assessment, mark_targets / select_targets / process_targets / lookback, sf_neural.py:36-176)
                           with `net.predict` = the oracle's float32 forward on the trained fixture weights
                           (sfneural_trained.npz, oracle/train_fixture.py): every instruction piped to the controller,
                           per frame, plus the targets / heat-map state
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
CODE = {'E': 0, 'B': 1, 'W': 2}


def codes(obj_arr):
    return np.vectorize(CODE.get)(obj_arr).astype(np.uint8)


def gen_geometry():
    from oracle import refimport
    refimport.load()
    from golib.config.golib_conf import gsize
    from camkifu.stone.stonesfinder import StonesFinder
    sf = StonesFinder(refimport.FakeVManager(None, gsize), learn_bg=False)
    rects = np.array([[sf.getrect(r, c) for c in range(gsize)] for r in range(gsize)], dtype=np.int32)
    mask = sf.getmask().astype(np.uint8)
    # pixels outside every zone are np.empty garbage in the reference (stonesfinder.py:468): blank them for comparison
    cover = np.zeros_like(mask)
    for r in range(gsize):
        for c in range(gsize):
            x0, y0, x1, y1 = rects[r, c]
            cover[x0:x1, y0:y1] = 1
    mask *= cover
    np.savez_compressed(os.path.join(GOLD, "geometry_g%d.npz" % gsize), rects=rects, mask=mask, cover=cover,
                        zone_area=np.int64(sf.zone_area), posgrid=sf._posgrid.mtx,
                        canonical_size=np.int64(sf.canonical_shape[0]))
    print("geometry", gsize, "zone_area", sf.zone_area)


def gen_clustering():
    import cv2
    from oracle import refimport
    from camkifu_b200 import synth
    refimport.load()
    from camkifu.stone.sf_clustering import SfClustering

    # --- streaming: the real _doframe -> _find path, 7 frames (find_stones fires at frames 0, 3, 6)
    frames, mtx, truth, corners = synth.make_clip(11, 7, 240, 320, new_board_every=3)
    vm = refimport.FakeVManager(mtx)
    sf = SfClustering(vm)
    del sf.bg_model  # learn_bg has no effect on the outputs recorded here (MOG2 only feeds get_foreground())
    gobans, accus, moves, seeds = [], [], [], []
    for i in range(frames.shape[0]):
        seed = 1000 + i
        cv2.setRNGSeed(seed)
        n0 = len(vm.controller.piped)
        sf._doframe(frames[i].copy())
        sf.total_f_processed += 1  # what VidProcessor.execute does after _doframe (video.py:106-107)
        gobans.append(sf.goban_img.copy())
        accus.append(sf.accu.copy())
        bulk = [a for (ins, a) in vm.controller.piped[n0:] if ins == "bulk"]
        mv = np.array([(CODE[m.color], m.y, m.x) for m in bulk[0][0]], dtype=np.int32) if bulk else np.zeros((0, 3), np.int32)
        moves.append(mv)
        seeds.append(seed)
    np.savez_compressed(os.path.join(GOLD, "clustering_stream.npz"), frames=frames, mtx=mtx, truth=truth,
                        goban=np.array(gobans), accu_last=accus[-1], accu_1=accus[1], seeds=np.array(seeds),
                        board=codes(vm.controller.stones),
                        **{"moves_%d" % i: m for i, m in enumerate(moves)})
    print("stream: board matches truth at", (codes(vm.controller.stones)[:, 6:13] == truth[-1][:, 6:13]).mean())

    # --- whole-board find_stones on uint8 canonical images
    out = {}
    for k, seed in enumerate((21, 22, 23)):
        fr, M, tr, _ = synth.make_clip(seed, 1, 240, 320)
        g = cv2.warpPerspective(fr[0], M, (380, 380))
        cv2.setRNGSeed(seed)
        ratios, centers = sf.cluster_colors(g.astype(np.float32))
        cv2.setRNGSeed(seed)
        stones = sf.find_stones(g)
        out["goban_%d" % k] = g
        out["ratios_%d" % k] = ratios
        out["centers_%d" % k] = centers
        out["stones_%d" % k] = codes(stones) if stones is not None else np.full((19, 19), 255, np.uint8)
        out["truth_%d" % k] = tr[0]
    # a low-density board: find_stones must return None (check_density, sf_clustering.py:170-178)
    rng = np.random.default_rng(5)
    st = np.zeros((19, 19), np.uint8)
    st[3, 3] = 1
    corners = synth.random_corners(rng, 240, 320)
    M = synth.board_homography(corners, 380)
    g = cv2.warpPerspective(synth.render_frame(rng, 240, 320, st, corners), M, (380, 380))
    cv2.setRNGSeed(77)
    res = sf.find_stones(g)
    out["goban_sparse"] = g
    out["sparse_is_none"] = np.bool_(res is None)
    # a sub-region call as SfMeta makes them (sf_meta.py:253)
    cv2.setRNGSeed(31)
    res = sf.find_stones(out["goban_0"], rs=6, re=13, cs=12, ce=19)
    out["stones_region"] = codes(res) if res is not None else np.full((19, 19), 255, np.uint8)
    out["seeds"] = np.array([21, 22, 23, 77, 31])
    np.savez_compressed(os.path.join(GOLD, "clustering_full.npz"), **out)
    print("full-board: accuracy", [(out["stones_%d" % k] == out["truth_%d" % k]).mean() for k in range(3)],
          "sparse none:", out["sparse_is_none"])


def gen_neural():
    import cv2
    from oracle import refimport
    from oracle import oracle as O
    from camkifu_b200 import synth, weights
    refimport.load()
    from camkifu.stone.nn_manager import NNManager
    from camkifu.stone.nn_cache import NNCache
    import camkifu.stone.sf_neural as sfn

    m = NNManager()
    origins = np.array([[m._get_rect_nn(*m._subregion(i, j)) for j in range(10)] for i in range(10)], np.int32)
    subregions = np.array([[m._subregion(i, j) for j in range(10)] for i in range(10)], np.int32)
    stones_of_label = np.array([codes(NNManager.compute_stones(k)) for k in range(81)], np.uint8)
    kat = {27: "EEEB", 36: "EEBB", 64: "BEBW"}  # test_tmanager.py:19-21
    for k, s in kat.items():
        assert "".join(NNManager.compute_stones(k)) == s
    ci = m.class_indices()
    assert list(ci[3, 2]) == list(range(54, 81)) and list(ci[0, 2]) == list(range(2, 81, 3))  # test_tmanager.py:23-27
    allB = np.full((2, 2), 'B', dtype=object)
    fr, M, tr, _ = synth.make_clip(41, 1, 240, 320)
    g = cv2.warpPerspective(fr[0], M, (380, 380))
    xs = m.generate_xs(g)
    np.savez_compressed(os.path.join(GOLD, "neural_geometry.npz"), rect_nn=origins, subregions=subregions,
                        stones_of_label=stones_of_label, class_indices=ci, label_allB=np.int64(
                            NNManager.compute_label(0, 2, 0, 2, allB)), goban=g, xs=xs,
                        split=np.int64(m.split), step=np.int64(m.step), nb_classes=np.int64(m.nb_classes),
                        r_width=np.int64(m.r_width), c_width=np.int64(m.c_width))

    # decode path with a scripted net: y comes from the C oracle's forward on seeded Glorot weights
    params = weights.glorot_params(seed=0)
    y = O.c_cnn_forward(xs, params)

    class ScriptedNet:
        def __init__(self):
            self.k = 0

        def predict(self, x):
            assert x.shape == (1, 40, 40, 3)
            idx = [i for i in range(100) if np.array_equal(xs[i], x[0])][0]
            return y[idx][None]

    NNManager._network = ScriptedNet()
    cache = NNCache(m, g)
    st = cache.predict_all_stones()
    vm = refimport.FakeVManager(M)
    sf = sfn.SfNeural(vm)
    sf.cache = NNCache(sf.manager, g)
    sf.predict_all()
    bulk = [a for (ins, a) in vm.controller.piped if ins == "bulk"]
    mv = np.array([(CODE[q.color], q.y, q.x) for q in bulk[0][0]], dtype=np.int32) if bulk else np.zeros((0, 3), np.int32)
    np.savez_compressed(os.path.join(GOLD, "neural_decode.npz"), y=y, stones=codes(st[:, :, 0]),
                        conf=st[:, :, 1].astype(np.float32), moves=mv, min_confidence=np.float64(sfn.MIN_CONFIDENCE))
    print("neural: kept", len(mv), "moves; conf range", float(st[:, :, 1].min()), float(st[:, :, 1].max()))


GAME_EVENTS = [(8, 1, 9, 9), (18, 2, 3, 14)]          # (frame, colour code, r, c)
NEURAL_EVENTS = [(14, 1, 9, 9), (26, 2, 3, 14), (36, 1, 15, 4)]


def gen_background():
    from oracle import refimport
    from camkifu_b200 import synth
    refimport.load()
    from camkifu.stone.stonesfinder import StonesFinder

    class Plain(StonesFinder):          # the base class with its background model, no detection
        def _find(self, goban_img):
            pass

        def _learn(self):
            pass

    frames, mtx, truth, _ = synth.make_game_clip(5, 30, 120, 160, events=GAME_EVENTS)
    vm = refimport.FakeVManager(mtx)
    sf = Plain(vm, learn_bg=True)
    sf.bg_init_frames = 10              # 50 for videos (stonesfinder.py:115): shortened so that both rates are exercised
    masks, zones = [], []
    for i in range(frames.shape[0]):
        sf._doframe(frames[i].copy())
        sf.total_f_processed += 1
        fg = sf.get_foreground() if i >= sf.bg_init_frames else sf._fg
        masks.append(np.packbits(fg > 0))
        zones.append([[np.sum(fg[a0:a1, b0:b1]) / 255 for (a0, b0, a1, b1) in [sf.getrect(r, c)]][0]
                      for r in range(19) for c in range(19)])
        assert set(np.unique(fg)) <= {0, 255}
    zones = np.array(zones, dtype=np.int32).reshape(-1, 19, 19)
    np.savez_compressed(os.path.join(GOLD, "background_stream.npz"), frames=frames, mtx=mtx, masks=np.array(masks),
                        zone_fg=zones, bg_init_frames=np.int64(sf.bg_init_frames), events=np.array(GAME_EVENTS))
    print("background: fg pixels per frame", [int(z.sum()) for z in zones])


def gen_meta():
    """SfMeta / SfContours per-zone statistics, from the UNMODIFIED reference:
      * SfContours.find_stones (sf_contours.py:48-111) on a synthetic board + foreground blob: the `zones` table it builds
        (visible flag + int16 mean B, G, R per zone, _norm_channels :113-126) and the `mask` of filled convex hulls it was
        computed from. The reference targets OpenCV 3 (`_, contours, hierarchy = cv2.findContours(..)`): under cv2 4.x
        the call is wrapped to return three values; `zones` / `mask` are locals, captured through wrappers around
        SfContours.find_color and cv2.drawContours. Nothing in the reference is edited.
      * Region.commit (sf_meta.py:305-340) on random detection histories: the moves it submits.
      * Region.check_foreground (sf_meta.py:342-381) for the 3 x 3 regions on random foreground masks."""
    import cv2
    from oracle import refimport
    from camkifu_b200 import synth
    refimport.load()
    from camkifu.core import imgutil
    from camkifu.stone.sf_contours import SfContours
    from camkifu.stone.sf_meta import Region, SfMeta
    out = {}
    # ---- zone means
    _fc, _dc = cv2.findContours, cv2.drawContours

    def fc3(*a, **k):
        r = _fc(*a, **k)
        return (None,) + tuple(r) if len(r) == 2 else r

    cap = {}

    def dc(img, conts, idx, color=None, **kw):
        if isinstance(color, tuple) and tuple(color) == (1, 1, 1) and kw.get("thickness") == -1:
            cap["mask"] = img
        return _dc(img, conts, idx, color, **kw) if color is not None else _dc(img, conts, idx, **kw)

    orig_find_color = SfContours.find_color

    def rec(r, c, zones, stones):
        cap["zones"] = zones.copy()
        return orig_find_color(r, c, zones, stones)

    cv2.findContours, cv2.drawContours = fc3, dc
    SfContours.find_color = staticmethod(rec)
    try:
        frames, mtx, truth, _ = synth.make_clip(31, 2, 360, 480)
        for k, (rs, re, cs, ce) in enumerate(((0, 19, 0, 19), (6, 13, 0, 7))):
            g = cv2.warpPerspective(frames[k], mtx, (380, 380))
            sf = SfContours(refimport.FakeVManager(mtx))
            sf.total_f_processed = 100
            fg = np.zeros((380, 380), np.uint8)
            cv2.circle(fg, (150, 90), 9, 255, -1)
            sf._fg = fg
            cap.clear()
            stones = sf.find_stones(g, rs=rs, re=re, cs=cs, ce=ce)
            x0, y0 = sf.getrect(rs, cs)[:2]
            x1, y1 = sf.getrect(re - 1, ce - 1)[2:]
            full = np.zeros((380, 380), np.uint8)
            full[x0:x1, y0:y1] = cap["mask"][:, :, 0]
            out["zm_img_%d" % k], out["zm_mask_%d" % k] = g, full
            out["zm_region_%d" % k] = np.array([rs, re, cs, ce])
            out["zm_zones_%d" % k] = cap["zones"]
            out["zm_stones_%d" % k] = codes(stones)
    finally:
        cv2.findContours, cv2.drawContours = _fc, _dc
        SfContours.find_color = staticmethod(orig_find_color)

    # ---- Region.commit / check_foreground with a scripted finder
    class FakeSf:
        contour, cluster = "contour", "cluster"

        def __init__(self, base):
            self.base, self.empty, self.fg, self.sent = base, None, None, []

        def is_empty(self, r, c):
            return bool(self.empty[r, c])

        def bulk_update(self, moves):
            self.sent.extend(moves)

        def suggest(self, color, r, c, doprint=True):
            self.sent.append((color, r, c))

        def get_foreground(self):
            return self.fg

        def getrect(self, r, c, cursor=1.0):
            return self.base.getrect(r, c, cursor)

        def stone_radius(self):
            return self.base.stone_radius()

    base = SfContours(refimport.FakeVManager(None))
    fake = FakeSf(base)
    bounds = [SfMeta.subregion(type("S", (), {"split": 3})(), r, c) for r in range(3) for c in range(3)]
    out["regions"] = np.array(bounds)
    rng = np.random.default_rng(7)
    hist_all, empty_all, moves_all = [], [], []
    colors = np.array(['E', 'B', 'W'], dtype=object)
    for trial in range(6):
        histo = 3 if trial < 4 else 5
        hist = np.zeros((19, 19, histo), np.uint8)
        moves = np.zeros((19, 19), np.uint8)
        fake.empty = rng.random((19, 19)) < 0.8
        for (rs, re, cs, ce) in bounds:
            reg = Region(fake, (rs, re, cs, ce), histo, finder=None, state="search")
            cb = imgutil.CyclicBuffer((re - rs, ce - cs), histo, dtype=object, init='E')
            h = rng.choice(3, size=(re - rs, ce - cs, histo), p=(0.5, 0.3, 0.2)).astype(np.uint8)
            stable = rng.random((re - rs, ce - cs)) < 0.4           # many intersections agree across the history
            h[stable] = h[stable][:, :1]
            from golib.config.golib_conf import E, B, W
            lut = np.array([E, B, W], dtype=object)
            cb.buffer[:] = lut[h]
            fake.sent = []
            reg.commit(cb)
            for col, r, c in fake.sent:
                moves[r, c] = CODE[col]
            hist[rs:re, cs:ce] = h
        hist_all.append(hist if histo == 3 else hist)
        empty_all.append(fake.empty.astype(np.uint8))
        moves_all.append(moves)
    for i in range(6):
        out["vote_hist_%d" % i], out["vote_empty_%d" % i], out["vote_moves_%d" % i] = hist_all[i], empty_all[i], moves_all[i]
    calm = []
    fgs = []
    for trial in range(12):
        fg = np.zeros((380, 380), np.uint8)
        for _ in range(int(rng.integers(0, 5))):
            ctr = (int(rng.integers(0, 380)), int(rng.integers(0, 380)))
            cv2.circle(fg, ctr, int(rng.integers(4, 40)), 255, -1)
        if trial % 4 == 3:
            cv2.rectangle(fg, (0, 0), (25, 25), 255, -1)              # a corner zone
        fake.fg = fg
        row = []
        for (rs, re, cs, ce) in bounds:
            reg = Region(fake, (rs, re, cs, ce), 3, finder=None, state="search")
            row.append(bool(reg.check_foreground()))
        calm.append(row)
        fgs.append(np.packbits(fg > 0))
    out["fg_masks"], out["fg_calm"] = np.array(fgs), np.array(calm)
    np.savez_compressed(os.path.join(GOLD, "meta.npz"), **out)
    print("meta: zones visible", int(out["zm_zones_0"][:, :, 0].sum()), "votes", [int((m > 0).sum()) for m in moves_all],
          "calm", np.array(calm).mean())


def gen_neural_stream():
    import cv2
    from oracle import refimport
    from oracle import oracle as O
    from camkifu_b200 import synth, weights
    refimport.load()
    from camkifu.stone.nn_manager import NNManager
    import camkifu.stone.sf_neural as sfn

    # realistic weights (oracle/train_fixture.py): the stream then carries real moves and peaked softmax outputs
    params = np.load(os.path.join(GOLD, "sfneural_trained.npz"))["params"]
    margins = []

    class OracleNet:                    # stands in for the Keras model: float32 forward of the same architecture
        def predict(self, x):
            y = O.c_cnn_forward(np.ascontiguousarray(x, dtype=np.uint8), params)
            top = np.sort(y[0])[-2:]
            margins.append((float(top[1] - top[0]), float(abs(y[0].max() / y[0].sum() - sfn.MIN_CONFIDENCE))))
            return y

    NNManager._network = OracleNet()
    n = 72
    frames, mtx, truth, _ = synth.make_game_clip(int(os.environ.get("CKB_NS_SEED", "11")), n, 120, 160, events=NEURAL_EVENTS)
    vm = refimport.FakeVManager(mtx)
    sf = sfn.SfNeural(vm)
    sf.bg_init_frames = 10
    log, targets, heat = [], [], []
    for i in range(n):
        n0 = len(vm.controller.piped)
        sf._doframe(frames[i].copy())
        sf.total_f_processed += 1
        for ins, a in vm.controller.piped[n0:]:
            if ins == "bulk":
                for m in a[0]:
                    log.append((i, 0, CODE[m.color], m.y, m.x))
            elif ins == "append":
                log.append((i, 1, CODE[a[0].color], a[0].y, a[0].x))
            elif ins == "delete":
                log.append((i, 2, 0, a[1], a[0]))
        targets.append(sf.targets.copy())
        heat.append(np.array([[0 if h is None else h.energy + 100 for h in row] for row in sf.heatmap], np.int32))
    np.savez_compressed(os.path.join(GOLD, "neural_stream.npz"), frames=frames, mtx=mtx, truth=truth,
                        log=np.array(log, dtype=np.int32).reshape(-1, 5), targets=np.array(targets), heat=np.array(heat),
                        board=codes(vm.controller.stones), bg_init_frames=np.int64(10), events=np.array(NEURAL_EVENTS),
                        min_label_margin=np.float64(min(m[0] for m in margins)),
                        min_conf_margin=np.float64(min(m[1] for m in margins)))
    print("neural stream:", len(log), "piped stone updates over", n, "frames;", len(margins), "net calls; min label margin",
          min(m[0] for m in margins), "min |conf - 0.6|", min(m[1] for m in margins))
    NNManager._network = None


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "geometry":
        gen_geometry()
    elif what == "background":
        gen_background()
    elif what == "neural_stream":
        gen_neural_stream()
    elif what == "meta":
        gen_meta()
    elif what == "all":
        for g in (9, 13, 19):
            subprocess.run([sys.executable, "-m", "oracle.gen_golden", "geometry"], check=True, cwd=ROOT,
                           env=dict(os.environ, CKB_GSIZE=str(g)))
        gen_clustering()
        gen_neural()
        gen_background()
        gen_neural_stream()
        gen_meta()
