"""ctypes front-end of the C oracle + a cv2-calling restatement of the reference's Python control flow.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

Two layers:
  * `c_*` functions: oracle/ck_oracle.c — plain-C restatement of the OpenCV / Keras arithmetic (bit-exact model).
  * `RefPath`: what the reference's Python does per frame, calling cv2 exactly where the reference calls it
    (stonesfinder.py:140, sf_clustering.py:33-36,99-129) — this is the CPU implementation a CamKifu user runs today
    and is what the cpu_baseline times. The CNN forward stands in for Keras `predict` with torch-CPU fp32.
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.cko_kmeans3.restype = C.c_double
        _lib.cko_rng_seed_state.restype = C.c_uint64
        _lib.cko_rng_seed_state.argtypes = [C.c_uint32]
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def c_invert3x3(M):
    M = np.ascontiguousarray(M, dtype=np.float64)
    out = np.empty((3, 3), np.float64)
    lib().cko_invert3x3(_p(M), _p(out))
    return out


def c_warp(frame, M, size):
    frame = np.ascontiguousarray(frame, dtype=np.uint8)
    M = np.ascontiguousarray(M, dtype=np.float64)
    out = np.empty((size, size, 3), np.uint8)
    lib().cko_warp_perspective_u8c3(_p(frame), C.c_int(frame.shape[0]), C.c_int(frame.shape[1]),
                                    C.c_size_t(frame.strides[0]), _p(M), _p(out), C.c_int(size), C.c_int(size))
    return out


def c_accumulate(src, acc, alpha=0.2, first=False):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    assert acc.dtype == np.float32 and acc.flags.c_contiguous
    lib().cko_accumulate_weighted(_p(src), _p(acc), C.c_size_t(src.size), C.c_float(alpha), C.c_int(int(first)))
    return acc


class CMog2:
    """cv2.createBackgroundSubtractorMOG2(detectShadows=False) restated (ck_oracle.c: cko_mog2_apply)."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        n = self.shape[0] * self.shape[1]
        self.state = np.zeros((n, 25), np.float32)
        self.nmodes = np.zeros(n, np.uint8)
        self.nframes = 0

    def apply(self, img, learningRate=-1.0):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        assert img.shape == self.shape + (3,)
        self.nframes += 1
        mask = np.empty(self.shape, np.uint8)
        lib().cko_mog2_apply(_p(img), C.c_int(self.nmodes.size), _p(self.state), _p(self.nmodes), C.c_int(self.nframes),
                             C.c_double(learningRate), _p(mask))
        return mask


def rng_seed_state(seed: int) -> int:
    return int(lib().cko_rng_seed_state(seed))


def c_kmeans(pixels, rng_state: int, K=3, eps=3.0, attempts=3):
    """returns (compactness, labels[N] i32, centers[K,3] f32, new_rng_state, iters[attempts])"""
    pixels = np.ascontiguousarray(pixels, dtype=np.float32).reshape(-1, 3)
    N = pixels.shape[0]
    labels = np.empty(N, np.int32)
    centers = np.empty((K, 3), np.float32)
    iters = np.zeros(attempts, np.int32)
    st = C.c_uint64(rng_state)
    comp = lib().cko_kmeans3(_p(pixels), C.c_int(N), C.c_int(K), C.c_double(eps), C.c_int(0), C.c_int(attempts),
                             C.byref(st), _p(labels), _p(centers), _p(iters))
    return comp, labels, centers, int(st.value), iters


def c_zone_rects(gsize=19):
    r = np.empty((gsize, gsize, 4), np.int32)
    lib().cko_zone_rects(C.c_int(gsize), _p(r))
    return r


def c_zone_mask(gsize=19):
    m = np.empty((20 * gsize, 20 * gsize), np.uint8)
    area = lib().cko_zone_mask(C.c_int(gsize), _p(m))
    return m, int(area)


def c_zone_classify(labels, centers, gsize=19, rs=0, re=None, cs=0, ce=None):
    re = gsize if re is None else re
    ce = gsize if ce is None else ce
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    ratios = np.empty((gsize, gsize, 3), np.uint8)
    stones = np.empty((gsize, gsize), np.uint8)
    ok = lib().cko_zone_classify(_p(labels), _p(centers), C.c_int(gsize), C.c_int(rs), C.c_int(re), C.c_int(cs),
                                 C.c_int(ce), _p(ratios), _p(stones))
    return ratios, stones, bool(ok)


def region_bbox(gsize, rs, re, cs, ce):
    """(x0, y0, x1, y1) of the sub-image cluster_colors() slices (sf_clustering.py:99-101)."""
    r = c_zone_rects(gsize)
    return int(r[rs, cs, 0]), int(r[rs, cs, 1]), int(r[re - 1, ce - 1, 2]), int(r[re - 1, ce - 1, 3])


def c_find_stones(img, rng_state: int, gsize=19, rs=0, re=None, cs=0, ce=None):
    """SfClustering.find_stones (sf_clustering.py:48-75) on the C oracle.
    returns dict(stones u8[g,g], trusted, ratios, centers, labels, compactness, rng_state)"""
    re = gsize if re is None else re
    ce = gsize if ce is None else ce
    x0, y0, x1, y1 = region_bbox(gsize, rs, re, cs, ce)
    sub = np.ascontiguousarray(img[x0:x1, y0:y1].astype(np.float32))
    comp, labels, centers, st, iters = c_kmeans(sub.reshape(-1, 3), rng_state)
    ratios, stones, ok = c_zone_classify(labels, centers, gsize, rs, re, cs, ce)
    return dict(stones=stones, trusted=ok, ratios=ratios, centers=centers, labels=labels.reshape(x1 - x0, y1 - y0),
                compactness=comp, rng_state=st, iters=iters)


# ---------------------------------------------------------------------------------------------------------------- CNN
CNN_SHAPES = [("w1", (5, 5, 3, 32)), ("b1", (32,)), ("w2", (5, 5, 32, 32)), ("b2", (32,)),
              ("w3", (3, 3, 32, 90)), ("b3", (90,)), ("w4", (3, 3, 90, 90)), ("b4", (90,)),
              ("w5", (3240, 160)), ("b5", (160,)), ("w6", (160, 81)), ("b6", (81,))]
CNN_NPARAM = 658665


def c_cnn_forward(xs, params, acc64=False, want_logits=False, want_acts=False):
    xs = np.ascontiguousarray(xs, dtype=np.uint8).reshape(-1, 40, 40, 3)
    params = np.ascontiguousarray(params, dtype=np.float32)
    assert params.size == CNN_NPARAM
    n = xs.shape[0]
    y = np.empty((n, 81), np.float32)
    logits = np.empty((n, 81), np.float32) if want_logits else None
    acts = np.empty((n, lib().cko_acts_per_patch()), np.float32) if want_acts else None
    lib().cko_cnn_forward(_p(xs), C.c_int(n), _p(params), C.c_int(int(acc64)), _p(y),
                          _p(logits) if want_logits else None, _p(acts) if want_acts else None)
    out = [y]
    if want_logits:
        out.append(logits)
    if want_acts:
        out.append(acts)
    return out[0] if len(out) == 1 else tuple(out)


def c_nn_gather(goban):
    goban = np.ascontiguousarray(goban, dtype=np.uint8)
    assert goban.shape == (380, 380, 3)
    xs = np.empty((100, 40, 40, 3), np.uint8)
    lib().cko_nn_gather(_p(goban), _p(xs))
    return xs


def c_nn_decode(y):
    y = np.ascontiguousarray(y, dtype=np.float32).reshape(100, 81)
    stones = np.empty((19, 19), np.uint8)
    conf = np.empty((19, 19), np.float32)
    keep = np.empty((19, 19), np.uint8)
    lib().cko_nn_decode(_p(y), _p(stones), _p(conf), _p(keep))
    return stones, conf, keep.astype(bool)


def c_compute_stones(label):
    four = np.empty(4, np.uint8)
    lib().cko_nn_compute_stones(C.c_int(int(label)), _p(four))
    return four


# ------------------------------------------------------------------------------------------- reference control flow
class RefPath:
    """The reference's per-frame CPU path restated with the same third-party calls (cv2 / numpy) it makes.

    warp       : stonesfinder.py:140            cv2.warpPerspective(frame, mtx, (S, S))
    accumulate : sf_clustering.py:33-36         astype(float32) / cv2.accumulateWeighted(., ., 0.2)
    find_stones: sf_clustering.py:48-178        cv2.kmeans + the 361-zone np.unique loop + interpret + density
    """

    def __init__(self, gsize=19):
        import cv2
        self.cv2 = cv2
        self.gsize = gsize
        self.S = 20 * gsize
        self.rects = c_zone_rects(gsize)  # geometry tables are constants (pinned against the reference's getrect)
        self.mask, self.zone_area = c_zone_mask(gsize)
        self.accu = None

    def warp(self, frame, mtx):
        return self.cv2.warpPerspective(frame, mtx, (self.S, self.S))

    def accumulate(self, goban_img):
        if self.accu is None:
            self.accu = goban_img.astype(np.float32)
        else:
            self.cv2.accumulateWeighted(goban_img, self.accu, 0.2)
        return self.accu

    def find_stones(self, img, rs=0, re=None, cs=0, ce=None):
        cv2 = self.cv2
        g = self.gsize
        re = g if re is None else re
        ce = g if ce is None else ce
        if img.dtype != np.float32:
            img = img.astype(np.float32)
        x0, y0 = self.rects[rs, cs, 0], self.rects[rs, cs, 1]
        x1, y1 = self.rects[re - 1, ce - 1, 2], self.rects[re - 1, ce - 1, 3]
        sub = img[x0:x1, y0:y1]
        pixels = np.reshape(sub, (sub.shape[0] * sub.shape[1], 3))
        crit = (cv2.TERM_CRITERIA_EPS, 15, 3)
        retval, labels, centers = cv2.kmeans(pixels, 3, None, crit, 3, cv2.KMEANS_PP_CENTERS)
        cvals = [int(sum(c) / 3) for c in centers]
        labels = np.reshape(labels, sub.shape[:2])
        labels += 1
        labels *= self.mask[x0:x1, y0:y1].astype(labels.dtype)
        ratios = np.zeros((g, g, 3), dtype=np.uint8)
        ratios[:, :, cvals.index(sorted(cvals)[1])] = 1
        for x in range(rs, re):
            for y in range(cs, ce):
                a0, b0, a1, b1 = self.rects[x, y]
                vals, counts = np.unique(labels[a0 - x0:a1 - x0, b0 - y0:b1 - y0], return_counts=True)
                for i in range(len(vals)):
                    if 0 < vals[i]:
                        ratios[x][y][vals[i] - 1] = 100 * counts[i] / sum(counts)
        colors = [1 if v == min(cvals) else (2 if v == max(cvals) else 0) for v in cvals]
        stones = np.zeros((g, g), np.uint8)
        for i in range(rs, re):
            for j in range(cs, ce):
                stones[i, j] = colors[int(np.argmax(ratios[i][j]))]
        vals, counts = np.unique(stones, return_counts=True)
        trusted = not (len(vals) < 3 or min(counts) < 2)
        return stones, trusted, centers, retval


_torch_net = None


def torch_cnn(params):
    """torch-CPU fp32 stand-in for `keras Model.predict` (nn_manager.py:277-298); weights from the flat param blob."""
    import torch
    import torch.nn as nn
    off = 0
    t = {}
    for name, shp in CNN_SHAPES:
        n = int(np.prod(shp))
        t[name] = torch.from_numpy(np.asarray(params[off:off + n], dtype=np.float32).reshape(shp).copy())
        off += n
    net = nn.Sequential(nn.Conv2d(3, 32, 5), nn.ReLU(), nn.Conv2d(32, 32, 5), nn.ReLU(), nn.MaxPool2d(2),
                        nn.Conv2d(32, 90, 3), nn.ReLU(), nn.Conv2d(90, 90, 3), nn.ReLU(), nn.MaxPool2d(2))
    convs = [net[0], net[2], net[5], net[7]]
    with torch.no_grad():
        for k, cv in enumerate(convs, start=1):
            cv.weight.copy_(t["w%d" % k].permute(3, 2, 0, 1))  # (kh,kw,cin,cout) -> (cout,cin,kh,kw)
            cv.bias.copy_(t["b%d" % k])
    w5, b5, w6, b6 = t["w5"], t["b5"], t["w6"], t["b6"]

    def predict(x_u8):
        x = torch.from_numpy(np.ascontiguousarray(x_u8)).to(torch.float32).permute(0, 3, 1, 2)
        with torch.no_grad():
            h = net(x)                                   # (n, 90, 6, 6)
            h = h.permute(0, 2, 3, 1).reshape(x.shape[0], 3240)  # Flatten in H, W, C order
            h = torch.relu(h @ w5 + b5)
            z = h @ w6 + b6
            return torch.softmax(z, dim=1).numpy()
    return predict


# --------------------------------------------------------------------------- SfMeta / SfContours zone statistics (f4)
def meta_subregions(gsize=19, split=3):
    """SfMeta.subregion for every (row, col) (sf_meta.py:101-125): [(rs, re, cs, ce)] row-major."""
    step = int(gsize / split)
    out = []
    for row in range(split):
        for col in range(split):
            re, ce = (row + 1) * step, (col + 1) * step
            if gsize - re < step:
                re = gsize
            if gsize - ce < step:
                ce = gsize
            out.append((row * step, re, col * step, ce))
    return out


def meta_zone_means(img, mask, gsize=19, rs=0, re=None, cs=0, ce=None):
    """The zone table of SfContours.find_stones (sf_contours.py:87-102) with _norm_channels (:113-126).
    img (S, S, 3) uint8; mask (S, S) 0/1 in canonical coordinates. Returns int16 (re-rs, ce-cs, 4)."""
    re = gsize if re is None else re
    ce = gsize if ce is None else ce
    rects = c_zone_rects(gsize)
    m3 = (mask != 0).astype(np.uint8)[:, :, None]
    visible_sub, masked_sub = img * m3, img * (1 - m3)
    zones = np.zeros((re - rs, ce - cs, 4), dtype=np.int16)
    for r in range(re - rs):
        for c in range(ce - cs):
            a0, b0, a1, b1 = (int(v) for v in rects[r + rs, c + cs])
            area = (a1 - a0) * (b1 - b0)
            visible_area = int(np.sum(m3[a0:a1, b0:b1, 0]))
            if 0.4 * area < visible_area:
                zones[r, c, 0] = 1
                src, norm = visible_sub[a0:a1, b0:b1], visible_area
            else:
                src, norm = masked_sub[a0:a1, b0:b1], area - visible_area
            for k in range(3):
                zones[r, c, k + 1] = int(int(np.sum(src[:, :, k])) / norm)
    return zones


def meta_vote(hist, empty):
    """Region.commit's per-intersection rule (sf_meta.py:318-331). hist (..., histo) uint8 codes, empty (...) bool.
    Returns uint8 move codes (0 = none, 1 = B, 2 = W)."""
    hist = np.asarray(hist)
    out = np.zeros(hist.shape[:-1], np.uint8)
    size = hist.shape[-1]
    for idx in np.ndindex(*hist.shape[:-1]):
        if not empty[idx]:
            continue
        vals, counts = np.unique(hist[idx], return_counts=True)
        if len(vals) == 2 and 0 in vals:
            k = 0 if vals[0] == 0 else 1
            if counts[k] / size < 0.4:
                out[idx] = vals[1 - k]
        elif len(vals) == 1 and 0 not in vals:
            out[idx] = vals[0]
    return out


def meta_outer_border(rs, re, cs, ce, gsize=19):
    """Region.outer_border (sf_meta.py:444-463): the (row, col) of the zones around the region, in the reference's order."""
    out = []
    x = max(0, cs - 1)
    for y in range(max(0, rs - 1), min(gsize, re + 1)):
        out.append((y, x))
    y = min(gsize - 1, re)
    for x in range(max(1, cs), min(gsize, ce + 1)):
        out.append((y, x))
    x = min(gsize - 1, ce)
    for y in range(min(gsize - 2, re - 1), max(-1, rs - 2), -1):
        out.append((y, x))
    y = max(0, rs - 1)
    for x in range(min(gsize - 2, ce - 1), max(0, cs - 1), -1):
        out.append((y, x))
    return out


def meta_check_foreground(fg, rs, re, cs, ce, gsize=19):
    """Region.check_foreground (sf_meta.py:342-381) on a foreground mask (0 / 255): True = calm."""
    import math
    rects = c_zone_rects(gsize)
    moving, border_threshold = 0, 2
    for (r, c) in meta_outer_border(rs, re, cs, ce, gsize):
        a0, b0, a1, b1 = (int(v) for v in rects[r, c])
        if (a1 - a0) * (b1 - b0) * 0.7 < np.sum(fg[a0:a1, b0:b1]) / 255:
            if (a0 == 0 or a1 == fg.shape[0] - 1) and (b0 == 0 or b1 == fg.shape[1] - 1):
                moving = border_threshold
            moving += 1
            if border_threshold <= moving:
                return False
    x0, y0 = rects[rs, cs, 0], rects[rs, cs, 1]
    x1, y1 = rects[re - 1, ce - 1, 2], rects[re - 1, ce - 1, 3]
    threshold = 3 * ((fg.shape[0] / gsize / 2) ** 2) * math.pi
    return not (threshold < np.sum(fg[x0:x1, y0:y1]) / 255)
