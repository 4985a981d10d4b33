#!/usr/bin/env python
"""MOG2 kernel: time per 64-frame launch for register allocations sized for 4 / 6 / 8 CTAs per SM (-DMOG2_MINB).
    python tools/mog2_probe.py build ; python tools/mog2_probe.py run (GPU box)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROBE = os.path.join(ROOT, "tools", "_probe")
VARIANTS = (4, 6, 8)
if sys.argv[1:] == ["build"]:
    from camkifu_b200 import build
    for v in VARIANTS:
        os.makedirs(os.path.join(PROBE, "obj_m%d" % v), exist_ok=True)
        print(build.build(force=True, extra_flags=["-DMOG2_MINB=%d" % v], out=os.path.join(PROBE, "libckb_mog2_%d.so" % v),
                          bdir=os.path.join(PROBE, "obj_m%d" % v)))
    sys.exit(0)
if sys.argv[1:] == ["run"]:
    for v in VARIANTS:
        subprocess.run([sys.executable, __file__, "child", str(v)],
                       env=dict(os.environ, CAMKIFU_B200_LIB=os.path.join(PROBE, "libckb_mog2_%d.so" % v)), check=False)
    sys.exit(0)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from camkifu_b200.engine import StoneEngine  # noqa: E402
eng = StoneEngine(19)
rng = np.random.default_rng(0)
gob = torch.from_numpy(rng.integers(0, 256, (64, 380, 380, 3), dtype=np.uint8)).cuda()
bg = eng.mog2_new_state()
fg = torch.empty((64, 380, 380), dtype=torch.uint8, device="cuda")
for i in range(3):
    eng.mog2_apply(gob, bg, 64 * i, 0.005, out=fg)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    eng.mog2_apply(gob, bg, 64 * (3 + i), 0.005, out=fg)
e1.record()
torch.cuda.synchronize()
print("MOG2_MINB=%s: %.4f ms per 64-frame launch, mask checksum %d" % (sys.argv[2], e0.elapsed_time(e1) / 20, int(fg.sum())))
