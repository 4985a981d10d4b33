#!/usr/bin/env python
"""k-means kernel probe: per-phase clock64 breakdown of ckb_kmeans_cluster_u8 (library built with -DKC_TIMING) and the
per-kernel times of ckb_find_stones on 64 full-board 1080p-derived canonical images.

    python tools/km_probe.py build      # here (no GPU): tools/_probe/libckb_timing.so
    python tools/km_probe.py run        # on the GPU box
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROBE = os.path.join(ROOT, "tools", "_probe")
LIB = os.path.join(PROBE, "libckb_timing.so")

VARIANTS = [int(v) for v in os.environ.get("KC_EXP_VARIANTS", "0").split(",")]

if sys.argv[1:] == ["build"]:
    from camkifu_b200 import build
    os.makedirs(os.path.join(PROBE, "obj"), exist_ok=True)
    for v in VARIANTS:
        print(build.build(force=True, extra_flags=(["-DKC_TIMING"] if os.environ.get("KC_TIMING", "1") == "1" else []) + ["-DKC_EXP=%d" % v], out=LIB.replace(".so", "%d.so" % v),
                          bdir=os.path.join(PROBE, "obj")))
    sys.exit(0)

if sys.argv[1:] == ["run"]:     # the instrumented library prints from the kernel: run it in a child process
    for v in VARIANTS:
        print("---- KC_EXP =", v, flush=True)
        env = dict(os.environ, CAMKIFU_B200_LIB=LIB.replace(".so", "%d.so" % v))
        subprocess.run([sys.executable, __file__, "child", "1"], env=env, check=False)
    if os.environ.get("KC_TIMING", "1") == "1":
        subprocess.run([sys.executable, __file__, "child", "0"], check=False)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from camkifu_b200 import synth  # noqa: E402
from camkifu_b200.engine import StoneEngine, rng_seed, rng_states  # noqa: E402

timing = sys.argv[2] == "1"
eng = StoneEngine(19)
frames, mtx, truth, _ = synth.make_clip_parallel(1000, 64, 1080, 1920)
goban = eng.warp(torch.from_numpy(frames).cuda(), mtx)
states = rng_states(rng_seed(0), 0, 64)
n = 64
for rep in range(3):
    eng.profile_begin(256)
    res = eng.find_stones(goban[:n], states[:n])
    torch.cuda.synchronize()
    prof = eng.profile_end()
print("timing build" if timing else "product build", "n =", n, [(k, round(v, 4)) for k, v in prof])
print("accuracy vs truth", float((res["stones"].cpu().numpy() == truth[:n]).mean()))
