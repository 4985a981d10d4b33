#!/usr/bin/env python
"""A/B of the two warp kernels (direct gather vs TMA-staged source tiles, CKB_WARP_STAGED=1): parity against the oracle
and time per 64 x 1080p frames. Run each variant in its own process (the switch is read once).

    python tools/warp_stage_probe.py            # runs both children
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) == 1:
    for v in ("0", "1"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, CKB_WARP_STAGED=v), check=False)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from camkifu_b200 import synth  # noqa: E402
from camkifu_b200.engine import StoneEngine  # noqa: E402
from oracle import oracle as O  # noqa: E402

eng = StoneEngine(19)
rows = []
for (H, W) in ((1080, 1920), (480, 640), (2160, 3840)):
    n = 64 if H <= 1080 else 16
    frames, mtx, _, _ = synth.make_clip_parallel(7, n, H, W)
    d = torch.from_numpy(frames).cuda()
    d2 = torch.roll(d, 3, 0).contiguous()
    out = torch.empty((n, 380, 380, 3), dtype=torch.uint8, device="cuda")
    eng.warp(d, mtx, out=out)
    torch.cuda.synchronize()
    ok = all(np.array_equal(out[k].cpu().numpy(), O.c_warp(frames[k], mtx, 380)) for k in (0, n // 2, n - 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        eng.warp(d2, mtx, out=out)
    e0.record()
    for i in range(20):
        eng.warp(d if i & 1 else d2, mtx, out=out)
    e1.record()
    torch.cuda.synchronize()
    rows.append("%dx%d n=%d: %.4f ms per call, bit-exact %s" % (W, H, n, e0.elapsed_time(e1) / 20, ok))
# wild homographies (taps outside the image: the staged kernel must fall back per tile)
rng = np.random.default_rng(3)
frame = rng.integers(0, 256, (1, 480, 640, 3), dtype=np.uint8)
bad = 0
for M in synth.wild_homographies(rng, 12):
    g = eng.warp(torch.from_numpy(frame).cuda(), M)
    bad += int(not np.array_equal(g[0].cpu().numpy(), O.c_warp(frame[0], M, 380)))
print("CKB_WARP_STAGED=%s | %s | wild homographies differing: %d" % (os.environ.get("CKB_WARP_STAGED"), " | ".join(rows), bad))
