"""Host-side mirror of the part of CamKifu's Python interface that the stone-detection path touches.

The reference's plugin boundary is a Python class contract: finders subclass
`camkifu.stone.StonesFinder` (src/camkifu/stone/stonesfinder.py:18), get frames through `_doframe(frame)`, publish
results through `suggest` / `remove` / `bulk_update`, and read the goban through the controller of their `vmanager`.
When the reference package is importable, camkifu_b200.plugins derives from the real base class. This module is what
the plugins derive from when it is not (the GPU box has no /root/reference): same names, argument meaning, coordinate
conventions and error behaviour, written from the interface description — only what the hot path needs.

Coordinates: (r, c) numpy row / column in finders; controller calls take (x = c, y = r) (stonesfinder.py:305,349).
"""

import numpy as np

try:  # the author's Golib, when installed: use its very constants (finders compare colours with `is`)
    from golib.config.golib_conf import gsize, E, B, W
    from golib.model import Move
except ImportError:
    gsize, E, B, W = 19, 'E', 'B', 'W'

    class Move:
        """golib.model.Move('np', (color, r, c)): .color, .x (= column), .y (= row)."""

        def __init__(self, ctype, ctuple=None, string=None, number=-1):
            if ctype != 'np' or ctuple is None:
                raise NotImplementedError("only numpy-coordinate moves are mirrored")
            self.color, r, c = ctuple
            self.x, self.y = int(c), int(r)
            self.number = number

        def __repr__(self):
            return "%s[%d,%d]" % (self.color, self.y, self.x)

CODE_TO_COLOR = (E, B, W)          # uint8 codes of the C ABI -> the Golib colour constants
canonical_size = 20 * gsize        # cvconf.py:10


class DeletedError(ValueError):
    """camkifu.core.exceptions.DeletedError (exceptions.py:46-60): a move targets a location the user just deleted."""

    def __init__(self, locations, message=None):
        if message is None:
            message = "Location has been deleted by user, so it is locked until pixels change significantly."
        super().__init__(message)
        self.locations = locations
        self.message = message


def zone_rect(r: int, c: int, g: int = gsize):
    """StonesFinder.getrect(r, c, cursor=1.0) (stonesfinder.py:412-450): intersections sit at 10 + 20 k, a zone spans
    the 20 pixels around one, the last row / column stop one pixel short of the image edge."""
    S = 20 * g
    return 20 * r, 20 * c, (S - 1 if r == g - 1 else 20 * r + 20), (S - 1 if c == g - 1 else 20 * c + 20)


class PosGridMirror:
    """PosGrid (stonesfinder.py:938-1010): pixel position of every intersection of the canonical image, `mtx[i, j] =
    (x, y)` int16 — the grid of zone centres, (10 + 20 i, 10 + 20 j) for the 380-pixel, 19-line canonical frame."""

    def __init__(self, size, g=gsize):
        self.size = size
        self.mtx = np.zeros((g, g, 2), dtype=np.int16)
        self.adjust_vect = np.zeros(2, dtype=np.float32)
        self.adjust_contribs = 0
        first = size / g / 2
        last = size - first
        for i in range(g):
            for j in range(g):
                self.mtx[i, j, 0] = (first * (g - 1 - i) + last * i) / (g - 1)
                self.mtx[i, j, 1] = (first * (g - 1 - j) + last * j) / (g - 1)


class StonesFinderBase:
    """Mirror of camkifu.stone.StonesFinder + the bits of camkifu.core.video.VidProcessor finders rely on."""

    def __init__(self, vmanager, learn_bg=True):
        self.vmanager = vmanager
        self.total_f_processed = 0          # VidProcessor: incremented by the frame loop after each _doframe
        self.metadata = {}
        self.goban_img = None
        self.canonical_shape = (canonical_size, canonical_size)
        self._posgrid = PosGridMirror(canonical_size)
        self.mask_cache = None
        self.zone_area = None
        self.intersections = None
        if learn_bg:
            video = getattr(vmanager, "current_video", None)
            is_img = isinstance(video, str) and video.lower().endswith((".png", ".jpg", ".jpeg"))
            self.bg_init_frames = 0 if is_img else 50

    # ---- frame loop hooks
    def ready_to_read(self):
        bf = getattr(self.vmanager, "board_finder", None)
        return bf is not None and getattr(bf, "mtx", None) is not None

    def _doframe(self, frame):
        raise NotImplementedError("the B200 plugins provide _doframe (device warp)")

    def _find(self, goban_img):
        raise NotImplementedError("Abstract method meant to be extended")

    def _learn_bg(self):
        pass   # the B200 plugins keep the MOG2 background model on the device (plugins._DeviceFrames._learn_bg)

    def _learn(self):
        """Hook of the frame loop (stonesfinder.py:178). The reference's base class digests user corrections here
        (deletion watch, CorrectionWarning): GUI-driven, out of scope for the detection path; both
        B200 finders override it with a no-op exactly as SfClustering does (sf_clustering.py:180-181)."""

    def get_foreground(self):
        raise ValueError("This StonesFinder doesn't seem to be segmenting background. See self.__init__()")

    def _show(self, img, name=None, latency=True, thumbnail=True, loc=None, max_frequ=2):
        pass   # display is the GUI's business

    def display_message(self, message, name=None, force=False):
        pass

    def _window_name(self):
        return type(self).__name__

    # ---- geometry (constant tables)
    def getrect(self, r, c, cursor=1.0):
        if cursor != 1.0:
            raise NotImplementedError("only the default zone size is mirrored")
        return zone_rect(r, c)

    def getmask(self, depth=1):
        if self.mask_cache is None:
            S = canonical_size
            mask = np.zeros((S, S), np.uint8)
            for r in range(gsize):
                for c in range(gsize):
                    x0, y0, x1, y1 = zone_rect(r, c)
                    h, w = x1 - x0, y1 - y0
                    yy, xx = np.ogrid[-h / 2:h - h / 2, -w / 2:w - w / 2]
                    mask[x0:x1, y0:y1] = xx * xx + yy * yy <= min(h / 2, w / 2) ** 2
            self.zone_area = int(mask[0:20, 0:20].sum())
            self.mask_cache = mask.astype(np.float64)
        if depth > 1:
            return np.repeat(self.mask_cache[:, :, None], depth, axis=2)
        return self.mask_cache

    # ---- goban access and result submission
    def is_empty(self, r, c):
        return self.vmanager.controller.is_empty_blocking(c, r)

    def get_stones(self):
        return self.vmanager.controller.get_stones()

    def suggest(self, color, r, c, doprint=True):
        move = Move('np', ctuple=(color, r, c))
        if doprint:
            print(move)
        self.vmanager.controller.pipe("append", move)
        self.vmanager.controller.pipe("auto_save")

    def remove(self, r, c):
        assert not self.is_empty(r, c), "Can't remove stone from empty intersection."
        move = Move('np', ("", r, c))
        self.vmanager.controller.pipe("delete", move.x, move.y)

    def bulk_update(self, tuples):
        """E on an occupied point removes the stone; B / W on an empty point adds one; on a point of the other colour
        the old stone is removed first; unchanged points are skipped. One "bulk" + "auto_save" for the lot. (The
        reference's base class also refuses locations the user has just deleted in the GUI — DeletedError, raised by the
        reference's own StonesFinder when the plugins inherit from it; this stand-alone mirror has no GUI to learn from.)"""
        moves = []
        for color, r, c in tuples:
            occupied = not self.is_empty(r, c)
            if color is E or color == E:
                if occupied:
                    moves.append(Move('np', (E, r, c)))
                continue
            if color not in (B, W):
                continue
            if occupied:
                if self.vmanager.controller.locate(c, r).color == color:
                    continue
                moves.append(Move('np', (E, r, c)))
            moves.append(Move('np', (color, r, c)))
        if moves:
            self.vmanager.controller.pipe("bulk", moves)
            self.vmanager.controller.pipe("auto_save")
