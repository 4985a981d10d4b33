#!/usr/bin/env python
"""Headless counterpart of the reference's test/mains/benchmark.py for the B200 plugins.

The reference runs every (reference.sgf, video) pair of a directory through VManager with the chosen board / stones
finders and prints `[name: match% in N s]` (test/mains/benchmark.py:68-101). Golib, Tk and recorded games are not available
here, so the "videos" are seeded synthetic game clips (camkifu_b200.synth.make_game_clip: a hand places stones on a fixed
board), the board finder is the manual one (a given homography), the finder runs frame by frame on the calling thread
exactly as VidProcessor.execute drives it (camkifu_b200.harness.run_frames), and the score is the share of
intersections of the final goban that equal the ground truth.

    python tools/detection_benchmark.py [--sf SfClusteringB200|SfNeuralB200] [--frames 90] [--height 1080 --width 1920]

SfNeuralB200 uses seeded random weights unless --weights points to a .npy blob (camkifu_b200.weights): the reference's
trained model does not ship, so its score only shows that the path runs.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from camkifu_b200 import plugins, synth, weights  # noqa: E402
from camkifu_b200.harness import HeadlessVManager, run_frames  # noqa: E402

CODE = {plugins.E: 0, plugins.B: 1, plugins.W: 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", default="SfClusteringB200", choices=["SfClusteringB200", "SfNeuralB200"])
    ap.add_argument("--frames", type=int, default=90)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--clips", type=int, default=3)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--bg-init-frames", type=int, default=None, dest="bg_init_frames")
    args = ap.parse_args()
    reports = []
    for k in range(args.clips):
        events = [(20 + 12 * j, 1 + j % 2, 3 + 4 * j, 5 + 3 * j) for j in range(4) if 20 + 12 * j + 8 < args.frames]
        frames, mtx, truth, _ = synth.make_game_clip(100 + k, args.frames, args.height, args.width, events=events)
        vm = HeadlessVManager(mtx, video="synthetic_%d.avi" % k)
        cls = getattr(plugins, args.sf)
        if args.sf == "SfNeuralB200":
            if args.weights and args.weights.endswith(".npz"):
                cls.cnn_params = np.load(args.weights)["params"]
            else:
                cls.cnn_params = np.load(args.weights) if args.weights else weights.glorot_params(seed=0)
        sf = cls(vm)
        if args.bg_init_frames is not None and hasattr(sf, "bg_init_frames"):
            sf.bg_init_frames = args.bg_init_frames
        if hasattr(sf, "set_rng_seed"):
            sf.set_rng_seed(k)
        t0 = time.time()
        run_frames(sf, frames[:2])          # the first frames create the CUDA context, pack the weights, ...
        t1 = time.time()
        ctl = run_frames(sf, frames[2:])
        dt, steady = time.time() - t0, time.time() - t1
        board = np.vectorize(CODE.get)(ctl.stones).astype(np.uint8)
        region = (slice(None), slice(6, 13)) if args.sf == "SfClusteringB200" else (slice(None), slice(None))
        match = float((board[region] == truth[-1][region]).mean())   # SfClustering._find looks at columns 6..12 only
        new = [(r, c) for (_, _, r, c) in events]
        found = sum(int(board[r, c] == truth[-1][r, c]) for r, c in new)
        reports.append("[synthetic_%d: %.1f%% in %.1f s; %.0f frames/s after start-up; %d of %d stones played during the "
                       "clip detected]" % (k, 100 * match, dt, (args.frames - 2) / steady, found, len(new)))
        print(reports[-1], flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
