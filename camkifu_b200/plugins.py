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
(vmanager.py:310-321), and detection runs on the device-resident image. There is no CPU fallback: without the CUDA
library or a GPU the first frame raises.
"""
import numpy as np

from . import hostapi
from .hostapi import gsize, E, B, W, CODE_TO_COLOR

MIN_CONFIDENCE = 0.6          # sf_neural.py:18
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

    def _device_image(self, img: np.ndarray):
        """The device copy of a canonical image: the one just warped when `img` is that very array, else an upload."""
        if img is self.goban_img and getattr(self, "_goban_on_device", False):
            return self._d_goban
        t = self._torch.from_numpy(np.ascontiguousarray(img))
        return t.to(self._engine_obj.device)[None]

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
        self.goban_img = self._d_goban[0].cpu().numpy()
        self._goban_on_device = True
        self._learn_bg()
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
            super().__init__(vmanager, learn_bg=False)   # MOG2 is not used by this finder's detection
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
            if not bool(res["trusted"][0].item()):
                return None
            return stones_from_codes(res["stones"][0].cpu().numpy())

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
        image go through the network in one tensor-core pass. Start-up follows the reference (load the net on frame 0,
        wait `bg_init_frames`, then `predict_all`); afterwards every region is re-evaluated each frame and submitted
        with the reference's `predict_moves` rule — the foreground-driven choice of regions (mark_targets /
        select_targets / lookback, sf_neural.py:72-176) is SURVEY.md section 8(f2), outside the path built here."""

        cnn_params = None   # class-level default: flat float32 blob (camkifu_b200.weights); set before the first frame

        def __init__(self, vmanager):
            super().__init__(vmanager, learn_bg=False)
            if not hasattr(self, "bg_init_frames"):
                video = getattr(vmanager, "current_video", None)
                still = isinstance(video, str) and video.lower().endswith((".png", ".jpg", ".jpeg"))
                self.bg_init_frames = 0 if still else 50
            self.cache = None
            self.has_sampled = False
            self.indices = class_indices()
            self._weights_loaded = False

        def _learn(self):
            pass

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
            self.cache = NNCacheB200(out["softmax"][0].cpu().numpy())
            self._last = {k: out[k][0].cpu().numpy() for k in ("stones", "conf", "keep")}
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
                self.process_targets([(i, j) for i in range(10) for j in range(10)])

        def predict_all(self):
            """SfNeural.predict_all (sf_neural.py:57-70): every non-empty intersection seen with confidence > 0.6."""
            stones, conf, keep = self._last["stones"], self._last["conf"], self._last["keep"]
            moves = [(CODE_TO_COLOR[stones[r, c]], r, c) for r in range(gsize) for c in range(gsize) if keep[r, c]]
            self.bulk_update(moves)

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

        def process_targets(self, targets):
            """SfNeural.process_targets (sf_neural.py:86-99) without the heat map: lopsided batches are dropped."""
            moves = self.predict_moves(targets)
            if not len(moves) or not self.get_color_ratio(moves) < 1:
                return
            if len(moves) == 1:
                try:
                    self.suggest(*moves.pop()[0:3], doprint=False)
                except Exception as de:  # DeletedError of whichever base is in use
                    if type(de).__name__ != "DeletedError":
                        raise
            elif len(moves):
                self.bulk_update([m[0:3] for m in moves])

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
