"""Headless test doubles and a single-threaded frame loop for the plugins.

They follow the reference's own test doubles — test/objects/controllerv_test.py (ControllerVDev: inline `pipe`,
`get_stones`), test/objects/vmanager_test.py:47-67 (VManagerSeq: finders run sequentially on the calling thread),
test/mains/benchmark.py:104-109 (DummyQueue: a non-None image queue disables cv2.imshow) — so that a video can be
pushed through a finder without Tk, Golib's GUI or threads, here and on a GPU box where the reference is absent.
"""
import numpy as np

from .hostapi import E, gsize, Move


class DummyQueue:
    def put(self, item):
        pass

    put_nowait = put


class HeadlessController:
    """Holds the goban and executes the controller commands a StonesFinder emits (stonesfinder.py:250-349)."""

    def __init__(self, video="synthetic.avi"):
        self.video = video
        self.bounds = (0, 1)
        self.stones = np.full((gsize, gsize), E, dtype=object)   # [row, column]
        self.piped = []

    def pipe(self, instruction, *args):
        self.piped.append((instruction, args))
        if instruction == "bulk":
            for mv in args[0]:
                self.stones[mv.y, mv.x] = mv.color
        elif instruction == "append":
            self.stones[args[0].y, args[0].x] = args[0].color
        elif instruction == "delete":
            self.stones[args[1], args[0]] = E

    def is_empty_blocking(self, x, y):
        return self.stones[y, x] == E

    def locate(self, x, y):
        col = self.stones[y, x]
        return None if col == E else Move('np', (col, y, x))

    def get_stones(self):
        return self.stones.copy()

    def bulk_moves(self):
        """[(colour, row, column)] lists of every "bulk" command received, in order."""
        return [[(m.color, m.y, m.x) for m in a[0]] for ins, a in self.piped if ins == "bulk"]


class FixedBoardFinder:
    """The 'manual board finder' of the benchmark configs: a given frame -> canonical homography."""

    def __init__(self, mtx):
        self.mtx = mtx


class _Capt:
    def get(self, prop):
        return 0


class HeadlessVManager:
    def __init__(self, mtx=None, video="synthetic.avi"):
        self.controller = HeadlessController(video)
        self.board_finder = FixedBoardFinder(mtx)
        self.imqueue = DummyQueue()
        self.current_video = video
        self.capt = _Capt()
        self.full_speed = True
        self.errors = []

    def error_raised(self, processor, error):
        self.errors.append(error)

    def confirm_stop(self, processor):
        pass


def run_frames(finder, frames):
    """What VidProcessor.execute does per frame (core/video.py:98-107), on the calling thread."""
    for frame in frames:
        finder._doframe(frame.copy())       # frames handed to a finder are private copies (vmanager.py:584)
        finder.total_f_processed += 1
    return finder.vmanager.controller
