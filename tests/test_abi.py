"""CPU-side checks of the C ABI: the library loads, exports every symbol include/camkifu_b200.h declares, and its
host-only helpers (geometry tables, cv::RNG stepping, 3x3 inverse) agree with the oracle / golden vectors."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from camkifu_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "camkifu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ckb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 15
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libcamkifu_b200.so does not export " + n
    assert set(names) == set(_lib.SIGNATURES)      # header and ctypes binding declare the same entry points
    assert _lib.lib().ckb_version() == 100


@pytest.mark.parametrize("gsize", [9, 13, 19])
def test_geometry_tables_match_reference(golden, gsize):
    g = golden("geometry_g%d.npz" % gsize)
    L = _lib.lib()
    rects = np.zeros((gsize, gsize, 4), np.int32)
    mask = np.zeros((20 * gsize, 20 * gsize), np.uint8)
    assert L.ckb_zone_rects(gsize, rects.ctypes.data_as(C.c_void_p)) == 0
    assert L.ckb_zone_mask(gsize, mask.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(rects, g["rects"])
    assert np.array_equal(mask * g["cover"], g["mask"])


def test_rng_and_inverse_helpers(oracle):
    L = _lib.lib()
    assert L.ckb_rng_seed(0) == oracle.rng_seed_state(0) == 0xffffffff
    assert L.ckb_rng_seed(1234) == oracle.rng_seed_state(1234)
    # one cv2.kmeans call consumes 39 draws: the oracle's state after a call equals 39 steps
    px = np.random.default_rng(0).integers(0, 256, (500, 3)).astype(np.float32)
    st0 = oracle.rng_seed_state(7)
    _, _, _, st1, _ = oracle.c_kmeans(px, st0)
    assert L.ckb_rng_advance(st0, 39) == st1
    # the host-side jump-ahead (one modular power) against stepping draw by draw
    from camkifu_b200.engine import rng_advance, rng_states
    for n in (0, 1, 2, 3, 50, 1234):
        assert rng_advance(st0, n) == L.ckb_rng_advance(st0, 39 * n)
    assert rng_states(st0, 5, 4) == [L.ckb_rng_advance(st0, 39 * (5 + i)) for i in range(4)]
    rng = np.random.default_rng(1)
    for _ in range(200):
        M = rng.normal(size=(3, 3)) * np.array([[1, 1, 500], [1, 1, 500], [1e-3, 1e-3, 1]])
        out = np.zeros((3, 3))
        assert L.ckb_invert_homography(M.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(out, oracle.c_invert3x3(M))
    assert L.ckb_invert_homography(np.zeros((3, 3)).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) != 0
