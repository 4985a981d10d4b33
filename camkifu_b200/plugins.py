"""Drop-in StonesFinder plugins backed by the B200 CUDA library.

    SfClusteringB200   replaces camkifu.stone.sf_clustering.SfClustering   (src/camkifu/stone/sf_clustering.py)
    SfNeuralB200       replaces camkifu.stone.sf_neural.SfNeural + NNCache (sf_neural.py, nn_cache.py, nn_manager.py)

Both keep the reference's plugin contract: constructor `(vmanager)`, frame hook `_doframe(frame)`
-> `_find(goban_img)`, results through `suggest` / `bulk_update`, the SfMeta delegate `find_stones(img, rs, re, cs, ce,
**kwargs) -> (19, 19) object array of E/B/W or None`, registration by `(module, class name)` in `cvconf.sfinders`
(cvconf.py:31-40, looked up by VManagerBase._reflect, vmanager.py:163-198). The base class is the reference's own
`camkifu.stone.StonesFinder` when that package is importable and camkifu_b200.hostapi.StonesFinderBase otherwise.

The warp lives in the base `_doframe` (stonesfinder.py:140), so `_doframe` is overridden: the frame goes to the device,
`ckb_warp` produces the canonical image there, `self.goban_img` is kept as the host copy other code reads
(vmanager.py:310-321), and detection runs on the device-resident image. Finders that ask for a background model
(`learn_bg`, stonesfinder.py:113-115) get the MOG2 model on the device as well (`ckb_mog2_apply` in `_learn_bg`,
bit-identical masks; `get_foreground()` returns the host copy, `is_agitated` reads per-zone counts reduced on the
device). There is no CPU fallback: without the CUDA library or a GPU the first frame raises.
"""
import numpy as np

from . import hostapi
from .hostapi import gsize, E, B, W, CODE_TO_COLOR

MIN_CONFIDENCE = 0.6          # sf_neural.py:18
TARGET_THRESH = 15            # sf_neural.py:19-21
TARGET_INCR = 5
NB_LOOKBACK = 3
_COLOR_INDEX = {E: 0, B: 1, W: 2}   # nn_manager.py:29-30


# ------------------------------------------------------------------------------------------------- label codec (host)
def subregion(i: int, j: int, split: int = 10, step: int = 2):
    """NNManager._subregion (nn_manager.py:92-126): rows / columns of region (i, j); the last region is shifted back
    so that every region holds step x step intersections (region 9 = rows 17..18)."""
    assert 0 <= i < split and 0 <= j < split
    rs, cs = min(i * step, gsize - step), min(j * step, gsize - step)
    return rs, rs + step, cs, cs + step


def compute_stones(label: int, dimension: int = 4) -> np.ndarray:
    """NNManager.compute_stones (nn_manager.py:246-254): base-3 digits of the class, least significant first."""
    out = np.ndarray(dimension, dtype=object)
    k = int(label)
    for d in range(dimension):
        out[d] = CODE_TO_COLOR[k % 3]
        k //= 3
    return out


def compute_label(rs, re, cs, ce, stones) -> int:
    """NNManager.compute_label (nn_manager.py:236-244)."""
    val = 0
    for r in range(rs, re):
        for c in range(cs, ce):
            val += _COLOR_INDEX[stones[r, c]] * 3 ** ((r - rs) * (ce - cs) + (c - cs))
    return val


def class_indices(nb_classes: int = 81) -> np.ndarray:
    """NNManager.class_indices (nn_manager.py:360-382): [intersection, colour] -> the classes coding that colour there."""
    dim = 4
    digits = np.array([[(k // 3 ** d) % 3 for d in range(dim)] for k in range(nb_classes)])
    out = np.empty((dim, 3, nb_classes // 3), dtype=np.uint8)
    for d in range(dim):
        for col in range(3):
            out[d, col] = np.where(digits[:, d] == col)[0]
    return out


def stones_from_codes(codes: np.ndarray) -> np.ndarray:
    """uint8 codes {0, 1, 2} -> object array of the E / B / W constants."""
    lut = np.empty(3, dtype=object)
    lut[0], lut[1], lut[2] = E, B, W
    return lut[np.asarray(codes, dtype=np.intp)]


class NNCacheB200:
    """NNCache (nn_cache.py) for one canonical image, with every region's softmax computed by one device call."""

    def __init__(self, y: np.ndarray):
        self.y = np.asarray(y, dtype=np.float32).reshape(10, 10, 81)

    def predict_y(self, i, j):
        return self.y[i, j]

    def predict_4_stones(self, i, j):
        y = self.y[i, j]
        stones = compute_stones(int(np.argmax(y))).reshape(2, 2)
        return stones, max(y) / sum(y)

    def predict_stone(self, r, c):
        y = self.y[r // 2, c // 2]
        return compute_stones(int(np.argmax(y)))[2 * (r % 2) + c % 2], max(y) / sum(y)

    def predict_all_stones(self):
        out = np.ndarray((gsize, gsize, 2), dtype=object)
        for i in range(10):
            for j in range(10):
                rs, re, cs, ce = subregion(i, j)
                out[rs:re, cs:ce, 0], out[rs:re, cs:ce, 1] = self.predict_4_stones(i, j)
        return out


class HeatPointB200:
    """A recent prediction awaiting confirmation (HeatPoint, sf_neural.py:198-244): it is re-checked NB_LOOKBACK times,
    must pass two thirds of the checks, and lingers a few frames after its energy is spent ("cooling") before the
    location becomes a target candidate again."""

    def __init__(self, color, confidence, stamp, energy=NB_LOOKBACK):
        self.target = self.energy = energy
        self.color, self.confidence, self.stamp = color, confidence, stamp
        self.nb_checks = self.nb_passed = 0

    @property
    def live(self):
        return self.energy > 0

    def check(self, color, confidence):
        self.nb_checks += 1
        self.energy -= 1
        agreed = color == self.color
        self.nb_passed += int(agreed)
        # running mean in which the initial confidence counts as the first sample
        self.confidence = (self.confidence * self.nb_checks + (confidence if agreed else 0)) / (self.nb_checks + 1)

    def is_valid(self):
        reachable = 2 * self.target / 3 <= self.nb_passed + self.energy
        if not reachable:
            self.energy = 0
            self.confidence = 0.0
        return reachable

    def is_cold(self):
        return self.energy < -5

    def cool(self):
        """What drawing the heat map does to a spent point in the reference (HeatPoint.__repr__ decrements the energy
        of points with energy <= 0 every time the map is rendered, i.e. once per steady-state frame)."""
        if self.energy <= 0:
            self.energy -= 1


# ----------------------------------------------------------------------------------------------------- device plumbing
class _DeviceFrames:
    """Mixin: engine, device frame / canonical buffers, and the `_doframe` that warps on the GPU."""

    _engine_obj = None

    def _engine(self):
        if self._engine_obj is None:
            from .engine import StoneEngine
            self._engine_obj = StoneEngine(gsize)
            import torch
            S = 20 * gsize
            self._torch = torch
            self._d_goban = torch.empty((1, S, S, 3), dtype=torch.uint8, device=self._engine_obj.device)
            self._d_frame = None
        return self._engine_obj

    def _upload_frame(self, frame: np.ndarray, mtx):
        torch = self._torch
        eng = self._engine_obj
        frame = np.ascontiguousarray(frame)
        if frame.ndim != 3 or frame.shape[2] != 3 or frame.dtype != np.uint8:
            raise ValueError("expected a BGR uint8 frame, got %s %s" % (frame.dtype, frame.shape))
        if self._d_frame is None or tuple(self._d_frame.shape[1:3]) != frame.shape[:2]:
            self._d_frame = torch.empty((1,) + frame.shape, dtype=torch.uint8, device=eng.device)
        eng.upload_frames(torch.from_numpy(frame)[None], self._d_frame, eng.frame_roi(mtx, frame.shape[0], frame.shape[1]))
        return self._d_frame

    def _fetch(self, **tensors):
        """Device tensors -> numpy arrays with ONE synchronisation: every copy goes into a pinned staging buffer of its
        own on the current stream, then the stream is waited for once (a `.cpu()` per tensor would block per tensor).
        The arrays are copies: they stay valid after the next frame."""
        torch = self._torch
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        for k, t in tensors.items():
            buf = self._pinned.get(k)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = self._pinned[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self._engine_obj.device).synchronize()
        return {k: self._pinned[k].numpy().copy() for k in tensors}

    def _device_image(self, img: np.ndarray):
        """The device copy of a canonical image: the one just warped when `img` is that very array, else an upload."""
        if img is self.goban_img and getattr(self, "_goban_on_device", False):
            return self._d_goban
        t = self._torch.from_numpy(np.ascontiguousarray(img))
        return t.to(self._engine_obj.device)[None]

    # ---- background model on the device (StonesFinder.__init__ learn_bg / _learn_bg / get_foreground)
    _bg_on = False

    def _enable_bg(self):
        """Call from __init__ after the base constructor ran with learn_bg=False (no cv2 model is created)."""
        self._bg_on = True
        self._bg_state = None
        self._bg_frames = 0
        self._d_fg = None
        self._fg_host = None
        self._zone_fg = None
        if not hasattr(self, "bg_init_frames"):
            video = getattr(self.vmanager, "current_video", None)
            still = isinstance(video, str) and video.lower().endswith((".png", ".jpg", ".jpeg"))
            self.bg_init_frames = 0 if still else 50          # stonesfinder.py:115

    def _learn_bg(self, fetch: bool = True):
        """Enqueue the background-model update of the frame just warped; returns the device tensor of per-zone foreground
        counts (None when the model is off). fetch=False leaves the read-back to the caller (_doframe reads it together
        with the canonical image: one synchronisation per frame)."""
        if not self._bg_on:
            return None
        if not getattr(self, "_goban_on_device", False):
            raise RuntimeError("the background model runs on the device image produced by _doframe")
        eng = self._engine()
        if self._bg_state is None:
            self._bg_state = eng.mog2_new_state()
            self._d_fg = self._torch.empty((1, 20 * gsize, 20 * gsize), dtype=self._torch.uint8, device=eng.device)
        learning = 0.01 if self.total_f_processed < self.bg_init_frames else 0.005      # stonesfinder.py:174
        eng.mog2_apply(self._d_goban, self._bg_state, self._bg_frames, [learning], out=self._d_fg)
        self._bg_frames += 1
        d_counts = eng.zone_fg_counts(self._d_fg)
        self._fg_host = None
        if fetch:
            self._zone_fg = self._fetch(zone_fg=d_counts)["zone_fg"][0]
        return d_counts

    def get_foreground(self):
        """The foreground mask of the last frame, (S, S) uint8 0 / 255 (stonesfinder.py:502-515)."""
        if not self._bg_on or self._d_fg is None:
            raise ValueError("This StonesFinder doesn't seem to be segmenting background. See self.__init__()")
        if self._fg_host is None:
            self._fg_host = self._d_fg[0].cpu().numpy()
        return self._fg_host

    def is_agitated(self, r, c, fg=None, ratio=0.7):
        """SfNeural.is_agitated (sf_neural.py:178-180): more than `ratio` of the zone rectangle is foreground. The
        per-zone foreground counts come from the device (ckb_zone_fg_counts); `fg` is accepted for compatibility."""
        a0, b0, a1, b1 = self.getrect(r, c)
        return (a1 - a0) * (b1 - b0) * ratio < int(self._zone_fg[r, c])

    def _doframe(self, frame):
        self.intersections = None
        bf = getattr(self.vmanager, "board_finder", None)
        transform = getattr(bf, "mtx", None) if bf is not None else None
        if transform is None:
            self._goban_on_device = False
            sup = getattr(super(), "_doframe", None)
            if sup is not None and not isinstance(self, hostapi.StonesFinderBase):
                sup(frame)   # the reference's "NO BOARD LOCATION AVAILABLE" branch (stonesfinder.py:148-152)
            return
        eng = self._engine()
        eng.warp(self._upload_frame(frame, transform), transform, out=self._d_goban)
        self._goban_on_device = True
        d_counts = self._learn_bg(fetch=False)       # background model enqueued behind the warp
        if d_counts is not None:
            got = self._fetch(goban=self._d_goban, zone_fg=d_counts)
            self._zone_fg = got["zone_fg"][0]
        else:
            got = self._fetch(goban=self._d_goban)
        self.goban_img = got["goban"][0]
        self._learn()
        self._find(self.goban_img)


def build_classes(Base):
    """The two plugin classes on top of `Base` (the reference's StonesFinder or its mirror)."""

    class SfClusteringB200(_DeviceFrames, Base):
        """K-means stones finder (see SfClustering): running average of canonical frames, 3-means colour clustering of
        the region's pixels (cv2.kmeans semantics incl. its RNG stream), per-intersection label histogram, B / E / W by
        centre brightness, density sanity check. `rng_state` is the cv::RNG state the next k-means call starts from
        (the reference draws from cv2's process-global generator; here the finder owns it — see set_rng_seed)."""

        def __init__(self, vmanager):
            super().__init__(vmanager, learn_bg=False)   # no cv2 model: like every finder of the reference this one
            self._enable_bg()                            # keeps a background model (get_foreground), here on the device
            self._d_accu = None
            self._has_accu = False
            self.rng_state = None

        def set_rng_seed(self, seed: int):
            """Equivalent of cv2.setRNGSeed(seed) for this finder's k-means calls."""
            from .engine import rng_seed
            self.rng_state = rng_seed(seed)

        @property
        def accu(self):
            return self._d_accu[0].cpu().numpy() if self._has_accu else None

        def _learn(self):
            pass

        def _find(self, goban_img):
            eng = self._engine()
            torch = self._torch
            if self._d_accu is None:
                self._d_accu = torch.empty((1, 20 * gsize, 20 * gsize, 3), dtype=torch.float32, device=eng.device)
            eng.accumulate(self._device_image(goban_img), self._d_accu[0], first=not self._has_accu)
            self._has_accu = True
            if not self.total_f_processed % 3:
                stones = self._find_stones_device(self._d_accu, 0, 19, 6, 13)
                if stones is not None:
                    self.bulk_update([(stones[i][j], i, j) for i in range(gsize) for j in range(gsize)])

        def _find_stones_device(self, d_img, rs, re, cs, ce):
            from .engine import rng_seed, rng_advance
            eng = self._engine()
            if self.rng_state is None:
                self.rng_state = rng_seed(0)
            res = eng.find_stones(d_img, [self.rng_state], rs, re, cs, ce)
            self.rng_state = rng_advance(self.rng_state, 1)
            got = self._fetch(km_stones=res["stones"], km_trusted=res["trusted"])
            if not bool(got["km_trusted"][0]):
                return None
            return stones_from_codes(got["km_stones"][0])

        def find_stones_regions(self, img, regions):
            """find_stones(img, rs, re, cs, ce) for every (rs, re, cs, ce) of `regions` (at most 16) in one set of
            launches: what SfMeta's 3 x 3 Regions ask of their `cluster` delegate on one frame (sf_meta.py:245-262),
            without nine serial calls. The RNG stream advances region after region, as the serial calls would. Returns
            a list with, per region, the (19, 19) object array of 'E' / 'B' / 'W' or None (density check failed)."""
            from .engine import rng_seed, rng_advance, rng_states
            eng = self._engine()
            if img.dtype != np.uint8:
                return [self.find_stones(img, *r) for r in regions]      # float32 images: one region per call
            if self.rng_state is None:
                self.rng_state = rng_seed(0)
            states = rng_states(self.rng_state, 0, len(regions))
            res = eng.find_stones_regions(self._device_image(img), regions, [states])
            self.rng_state = rng_advance(self.rng_state, len(regions))
            got = self._fetch(kmr_stones=res["stones"], kmr_trusted=res["trusted"])
            return [stones_from_codes(got["kmr_stones"][0, k]) if got["kmr_trusted"][0, k] else None
                    for k in range(len(regions))]

        def find_stones(self, img, rs=0, re=gsize, cs=0, ce=gsize, **kwargs):
            """SfClustering.find_stones (sf_clustering.py:48-75). img: (S, S, 3) uint8 or float32 canonical image."""
            self._engine()
            if img.dtype not in (np.uint8, np.float32):
                img = img.astype(np.float32)
            return self._find_stones_device(self._device_image(img), rs, re, cs, ce)

        def _window_name(self):
            return "camkifu_b200.SfClusteringB200"

    class SfNeuralB200(_DeviceFrames, Base):
        """CNN stones finder (see SfNeural): every frame the 100 overlapping 2x2-intersection patches of the canonical
        image go through the network in one tensor-core pass and the MOG2 background model is updated on the device.
        The control flow is the reference's (sf_neural.py:36-176): load the net on frame 0, wait `bg_init_frames`,
        one `predict_all`, then per frame `mark_targets` (zones whose foreground is agitated accumulate heat),
        `select_targets` (hot regions that have calmed down), `process_targets` (submit what the network sees there)
        and `lookback` (re-check recent predictions, cancel the inconsistent ones). The reference evaluates the network
        lazily per region; here all 100 softmax vectors of the frame are already in the cache."""

        cnn_params = None   # class-level default: flat float32 blob (camkifu_b200.weights); set before the first frame

        def __init__(self, vmanager):
            super().__init__(vmanager, learn_bg=False)       # no cv2 model: the background model lives on the device
            self._enable_bg()
            self.cache = None
            self.has_sampled = False
            self.indices = class_indices()
            self.targets = np.zeros((gsize, gsize), dtype=np.uint8)     # per-location agitation (wraps like the reference's)
            self.heatmap = np.ndarray((gsize, gsize), dtype=object)     # recent predictions awaiting confirmation
            self.heatmap[:] = None
            self._weights_loaded = False

        def _load_net(self):
            from . import weights
            eng = self._engine()
            params = self.cnn_params if self.cnn_params is not None else weights.glorot_params(seed=0)
            eng.set_cnn_weights(params)
            self._weights_loaded = True

        def _predict(self, goban_img):
            if not self._weights_loaded:
                self._load_net()
            out = self._engine().cnn_forward(self._device_image(goban_img))
            got = self._fetch(softmax=out["softmax"], nn_stones=out["stones"], nn_conf=out["conf"], nn_keep=out["keep"])
            self.cache = NNCacheB200(got["softmax"][0])
            self._last = {"stones": got["nn_stones"][0], "conf": got["nn_conf"][0], "keep": got["nn_keep"][0]}
            return self.cache

        def _find(self, goban_img):
            if self.total_f_processed == 0:
                self._load_net()                       # SfNeural._find: the first frame only loads the net
            elif self.total_f_processed < self.bg_init_frames:
                pass                                   # "BACKGROUND SAMPLING": the reference waits bg_init_frames
            elif not self.has_sampled:
                self._predict(goban_img)
                self.predict_all()
                self.has_sampled = True
            else:
                self._predict(goban_img)
                self.mark_targets()
                self.process_targets()
                self.lookback()

        def predict_all(self):
            """SfNeural.predict_all (sf_neural.py:57-70): every non-empty intersection seen with confidence > 0.6."""
            stones, conf, keep = self._last["stones"], self._last["conf"], self._last["keep"]
            moves = []
            for r in range(gsize):
                for c in range(gsize):
                    if keep[r, c]:
                        color = CODE_TO_COLOR[stones[r, c]]
                        moves.append((color, r, c))
                        self.heatmap[r, c] = HeatPointB200(color, conf[r, c], self.total_f_processed)
            self.bulk_update(moves)

        def mark_targets(self, canvas=None):
            """SfNeural.mark_targets (sf_neural.py:72-84): locations without a pending prediction heat up while their
            zone is agitated, and every warm location cools by one per frame."""
            for r in range(gsize):
                for c in range(gsize):
                    if self.heatmap[r, c] is None and self.is_agitated(r, c):
                        self.targets[r, c] += TARGET_INCR
            self.targets[self.targets > 0] -= 1

        def select_targets(self, canvas=None):
            """SfNeural.select_targets (sf_neural.py:129-154): regions holding a location hotter than TARGET_THRESH, once
            none of their zones is agitated any more (ratio 0.5); selecting a region resets its heat."""
            chosen = []
            for i in range(10):
                for j in range(10):
                    rs, re, cs, ce = subregion(i, j)
                    if not (self.targets[rs:re, cs:ce] > TARGET_THRESH).any():
                        continue
                    if any(self.is_agitated(a, b, ratio=0.5) for a in range(rs, re) for b in range(cs, ce)):
                        continue
                    chosen.append((i, j))
                    self.targets[rs:re, cs:ce] = 0
            return chosen

        def predict_moves(self, targets):
            """SfNeural.predict_moves (sf_neural.py:101-127)."""
            moves = set()
            if not len(targets):
                return moves
            stones = self.get_stones()
            for i, j in targets:
                new_stones, confidence = self.cache.predict_4_stones(i, j)
                if confidence < MIN_CONFIDENCE:
                    continue
                rs, re, cs, ce = subregion(i, j)
                for a, b in np.transpose(np.where(new_stones != E)):
                    r, c = int(a + rs), int(b + cs)
                    if stones[r, c] == E:
                        moves.add((new_stones[a, b], r, c, confidence))
            return moves

        @staticmethod
        def get_color_ratio(moves):
            """SfNeural.get_color_ratio (sf_neural.py:185-194): |log3(#B / #W)| with both counts bumped if one is 0."""
            import math
            count = {B: 0, W: 0}
            for m in moves:
                if m[0] != E:
                    count[m[0]] += 1
            if 0 in count.values():
                count[B] += 1
                count[W] += 1
            return abs(math.log(count[B] / count[W], 3))

        def process_targets(self, targets=None):
            """SfNeural.process_targets (sf_neural.py:86-99): what the network sees in the selected regions is submitted
            (a single stone through `suggest`, several through `bulk_update`) unless the batch is lopsided in colour, and
            every submitted stone becomes a heat point."""
            if targets is None:
                targets = self.select_targets()
            moves = self.predict_moves(targets)
            if not len(moves) or not self.get_color_ratio(moves) < 1:
                return
            for color, r, c, confidence in moves:
                self.heatmap[r, c] = HeatPointB200(color, confidence, self.total_f_processed)
            if len(moves) == 1:
                try:
                    self.suggest(*moves.pop()[0:3], doprint=False)
                except Exception as de:  # DeletedError of whichever base is in use
                    if type(de).__name__ != "DeletedError":
                        raise
                    print(de)
            else:
                self.bulk_update([m[0:3] for m in moves])

        def lookback(self, canvas=None):
            """SfNeural.lookback (sf_neural.py:156-176): a live heat point whose stone is still on the goban is
            re-predicted every 11th frame; one that can no longer pass two thirds of its checks is cancelled."""
            stones = self.get_stones()
            cancelled = []
            for r in range(gsize):
                for c in range(gsize):
                    hp = self.heatmap[r, c]
                    if hp is None or not hp.live:
                        continue
                    if hp.color != stones[r, c]:
                        self.heatmap[r, c] = None        # someone else changed the location: leave it alone
                        continue
                    if 10 < self.total_f_processed - hp.stamp:
                        hp.stamp = self.total_f_processed
                        hp.check(*self.cache.predict_stone(r, c))
                        if not hp.is_valid():
                            cancelled.append((E, r, c))
            if cancelled:
                self.bulk_update(cancelled)
            for r in range(gsize):
                for c in range(gsize):
                    hp = self.heatmap[r, c]
                    if hp is not None and hp.is_cold():
                        self.heatmap[r, c] = None
            for r in range(gsize):                       # the reference renders the map here, which cools spent points
                for c in range(gsize):
                    if self.heatmap[r, c] is not None:
                        self.heatmap[r, c].cool()

        def find_stones(self, img, rs=0, re=gsize, cs=0, ce=gsize, **kwargs):
            """The SfMeta delegate contract for the CNN finder: board state of the intersections in [rs, re) x [cs, ce)
            that the network reports with confidence > 0.6 (E elsewhere)."""
            self._engine()
            self._predict(img)
            codes = np.where(self._last["keep"], self._last["stones"], 0)
            out = np.zeros((gsize, gsize), np.uint8)
            out[rs:re, cs:ce] = codes[rs:re, cs:ce]
            return stones_from_codes(out)

        def _window_name(self):
            return "camkifu_b200.SfNeuralB200"

    SfClusteringB200.__qualname__ = "SfClusteringB200"
    SfNeuralB200.__qualname__ = "SfNeuralB200"
    return SfClusteringB200, SfNeuralB200


def _default_base():
    try:
        from camkifu.stone import StonesFinder   # the reference package, when installed next to this one
        return StonesFinder
    except ImportError:
        return hostapi.StonesFinderBase


def bind(base=None):
    """(Re)create the module-level plugin classes on top of `base` (default: the reference's StonesFinder if it can be
    imported now, else the mirror). Called at import and again by register(), since an application may put the
    reference on sys.path after importing this module."""
    global SfClusteringB200, SfNeuralB200
    base = base or _default_base()
    cur = globals().get("SfClusteringB200")
    if cur is None or base not in cur.__mro__:
        SfClusteringB200, SfNeuralB200 = build_classes(base)
        SfClusteringB200.__module__ = SfNeuralB200.__module__ = __name__
    return SfClusteringB200, SfNeuralB200


bind()


def register(cvconf=None):
    """Append the plugins to the reference's finder registry (camkifu.config.cvconf.sfinders) so that
    `VManager(controller, imqueue, bf, sf="SfNeuralB200")` / `ckmain --sf SfClusteringB200` find them by name."""
    if cvconf is None:
        from camkifu.config import cvconf
    bind()
    for name in ("SfClusteringB200", "SfNeuralB200"):
        entry = (__name__, name)
        if entry not in cvconf.sfinders:
            cvconf.sfinders.append(entry)
    return cvconf.sfinders
