"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference (ArnaudPel/CamKifu) from /root/reference.

Used by oracle/gen_golden.py (golden-vector generation in the authoring container) and by the `not gpu` tests that
check the drop-in boundary against the real `camkifu.core.VManager` reflection. /root/reference does not exist on the
GPU box, so nothing that runs there may call `load()`; `available()` says whether the tree is present.

The fakes below follow the patterns of the reference's own test doubles:
  - test/objects/controllerv_test.py:7-43  (ControllerVDev: inline pipe(), get_stones())
  - test/mains/benchmark.py:104-109        (DummyQueue: non-None imqueue => no cv2.imshow)
"""
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("CKB_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "camkifu", "stone", "stonesfinder.py"))


def load():
    """Put the shim + the reference on sys.path (idempotent) and return the `camkifu` package."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    for p in (os.path.join(REF_ROOT), os.path.join(REF_ROOT, "src"), _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)  # vmanager.py:315 has an invalid escape sequence
        import camkifu
        import camkifu.core
        import camkifu.stone
    return camkifu


class DummyQueue:
    def put(self, x):
        pass

    def put_nowait(self, x):
        pass


class FakeController:
    """Goban state + the controller calls the hot path makes (stonesfinder.py:250-349)."""

    def __init__(self, gsize=19, video="synthetic.avi"):
        self.gsize = gsize
        self.video = video
        self.bounds = (0, 1)
        self.stones = np.full((gsize, gsize), 'E', dtype=object)  # [r, c]
        self.piped = []

    def pipe(self, instruction, *args):
        self.piped.append((instruction, args))
        if instruction == "bulk":
            for mv in args[0]:
                self.stones[mv.y, mv.x] = mv.color
        elif instruction == "append":
            mv = args[0]
            self.stones[mv.y, mv.x] = mv.color
        elif instruction == "delete":
            x, y = args
            self.stones[y, x] = 'E'

    def is_empty_blocking(self, x, y):
        return self.stones[y, x] == 'E'

    def locate(self, x, y):
        from golib.model import Move
        col = self.stones[y, x]
        return None if col == 'E' else Move('np', (col, y, x))

    def get_stones(self):
        return self.stones.copy()


class FakeCapt:
    def get(self, prop):
        return 0


class FakeBoardFinder:
    def __init__(self, mtx=None):
        self.mtx = mtx


class FakeVManager:
    def __init__(self, mtx=None, gsize=19, video="synthetic.avi"):
        self.controller = FakeController(gsize, video)
        self.board_finder = FakeBoardFinder(mtx)
        self.imqueue = DummyQueue()
        self.current_video = video
        self.capt = FakeCapt()
        self.errors = []

    def error_raised(self, proc, exc):
        self.errors.append(exc)

    def confirm_stop(self, proc):
        pass
