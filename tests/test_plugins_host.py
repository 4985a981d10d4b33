"""Host-side (no GPU) checks of the drop-in boundary: label codec against the reference's known-answer tests and golden
vectors, the interface mirror against the reference's geometry, and — when the reference tree is present — registration
of the plugins with the real camkifu.core.VManager reflection."""
import numpy as np
import pytest

from camkifu_b200 import hostapi, plugins
from camkifu_b200.harness import HeadlessVManager


def test_codec_known_answers(golden):
    # the reference's own known-answer tests (test/camkifu/stone/test_tmanager.py:18-27)
    assert "".join(plugins.compute_stones(27)) == "EEEB"
    assert "".join(plugins.compute_stones(36)) == "EEBB"
    assert "".join(plugins.compute_stones(64)) == "BEBW"
    ci = plugins.class_indices()
    assert list(ci[3, 2]) == list(range(54, 81)) and list(ci[0, 2]) == list(range(2, 81, 3))
    g = golden("neural_geometry.npz")
    assert np.array_equal(ci, g["class_indices"])
    code = {'E': 0, 'B': 1, 'W': 2}
    for k in range(81):
        assert [code[s] for s in plugins.compute_stones(k)] == list(g["stones_of_label"][k])
    allB = np.full((2, 2), 'B', dtype=object)
    assert plugins.compute_label(0, 2, 0, 2, allB) == int(g["label_allB"]) == 40
    for i in range(10):
        for j in range(10):
            assert plugins.subregion(i, j) == tuple(g["subregions"][i, j])


def test_cache_decode_matches_reference_golden(golden):
    """NNCacheB200 on the golden softmax reproduces NNCache.predict_all_stones of the reference."""
    g = golden("neural_decode.npz")
    cache = plugins.NNCacheB200(g["y"])
    st = cache.predict_all_stones()
    code = {'E': 0, 'B': 1, 'W': 2}
    assert np.array_equal(np.vectorize(code.get)(st[:, :, 0]).astype(np.uint8), g["stones"])
    assert np.array_equal(st[:, :, 1].astype(np.float32), g["conf"])
    s, c = cache.predict_stone(4, 7)
    assert code[s] == g["stones"][4, 7] and np.float32(c) == g["conf"][4, 7]
    # reference quirk kept: on row / column 18 predict_stone indexes region 9 with r % 2 == 0, i.e. reads row 17
    s, c = cache.predict_stone(18, 18)
    assert code[s] == g["stones"][17, 17]


@pytest.mark.parametrize("base", ["mirror"])
def test_mirror_geometry_and_bulk_update(golden, base):
    g = golden("geometry_g19.npz")
    sf = hostapi.StonesFinderBase(HeadlessVManager(), learn_bg=False)
    rects = np.array([[sf.getrect(r, c) for c in range(19)] for r in range(19)], dtype=np.int32)
    assert np.array_equal(rects, g["rects"])
    assert np.array_equal(sf.getmask().astype(np.uint8) * g["cover"], g["mask"])
    assert sf.zone_area == int(g["zone_area"]) == 315
    # bulk_update semantics: add, skip unchanged, recolour = remove + add, E removes
    ctl = sf.vmanager.controller
    sf.bulk_update([('B', 3, 4), ('W', 5, 6), ('E', 7, 7)])
    assert ctl.stones[3, 4] == 'B' and ctl.stones[5, 6] == 'W'
    n = len(ctl.piped)
    sf.bulk_update([('B', 3, 4)])
    assert len(ctl.piped) == n                       # nothing new: no command
    sf.bulk_update([('W', 3, 4), ('E', 5, 6)])
    last = ctl.bulk_moves()[-1]
    assert last == [('E', 3, 4), ('W', 3, 4), ('E', 5, 6)]
    assert ctl.stones[3, 4] == 'W' and ctl.stones[5, 6] == 'E'
    assert ctl.piped[-1][0] == "auto_save"


def test_plugins_construct_without_gpu_and_tolerate_no_vmanager():
    # SfMeta builds its delegates with vmanager=None (sf_meta.py:52): construction must not dereference it
    for cls in (plugins.SfClusteringB200, plugins.SfNeuralB200):
        sf = cls(None)
        assert sf.total_f_processed == 0 and sf.goban_img is None
        assert callable(sf.find_stones) and callable(sf._find) and callable(sf._doframe)
    sf = plugins.SfNeuralB200(HeadlessVManager(video="photo.png"))
    assert sf.bg_init_frames == 0
    assert plugins.SfNeuralB200(HeadlessVManager()).bg_init_frames == 50


def test_registration_with_reference_vmanager():
    """The reference looks finders up BY NAME in cvconf.sfinders (vmanager.py:163-198)."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip("reference tree not present")
    refimport.load()
    from camkifu.config import cvconf
    from camkifu.core.vmanager import VManagerBase
    from camkifu.stone import StonesFinder
    plugins.register(cvconf)
    assert ("camkifu_b200.plugins", "SfNeuralB200") in cvconf.sfinders
    Sc = plugins.SfClusteringB200
    found = VManagerBase._reflect("SfClusteringB200", cvconf.sfinders)
    assert found is plugins.SfClusteringB200 and issubclass(found, StonesFinder)
    assert VManagerBase._reflect("SfNeuralB200", cvconf.sfinders) is plugins.SfNeuralB200
    # instantiation the way check_sf does it (vmanager.py:367): sf_class(vmanager), no GPU needed yet
    sf = Sc(refimport.FakeVManager(None))
    assert sf.canonical_shape == (380, 380) and sf.getrect(18, 18) == (360, 360, 379, 379)


def test_heat_point_mirrors_reference():
    """HeatPointB200 against the reference's HeatPoint (sf_neural.py:198-244) under random check / render sequences,
    including the energy decrement that rendering the heat map performs on spent points."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip("reference tree not present")
    refimport.load()
    import camkifu.stone.sf_neural as sfn
    from camkifu_b200 import plugins
    rng = np.random.default_rng(0)
    for trial in range(200):
        color = ('B', 'W')[trial % 2]
        a = sfn.HeatPoint(color, 0.9, 5)
        b = plugins.HeatPointB200(color, 0.9, 5)
        for step in range(14):
            if (a == sfn.HeatPoint) and rng.random() < 0.6:
                c, conf = ('B', 'W', 'E')[rng.integers(3)], float(rng.random())
                a.check(c, conf)
                b.check(c, conf)
                assert a.is_valid() == b.is_valid()
            str(a)                      # what _drawvalues does every steady-state frame
            b.cool()
            assert (a.energy, a.nb_checks, a.nb_passed) == (b.energy, b.nb_checks, b.nb_passed)
            assert a.confidence == b.confidence
            assert (a == sfn.HeatPoint) == b.live and (a == sfn.COLD) == b.is_cold()


def test_posgrid_mirror_matches_reference(golden):
    """hostapi.PosGridMirror against PosGrid.mtx recorded from the reference (geometry_g19.npz)."""
    from camkifu_b200 import hostapi
    g = golden("geometry_g19.npz")
    pg = hostapi.PosGridMirror(380)
    assert pg.mtx.dtype == np.int16 and np.array_equal(pg.mtx, g["posgrid"])
