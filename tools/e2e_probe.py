"""Development probe: where does DetectPipeline.detect() spend its time? (H2D alone, compute alone, both)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from camkifu_b200 import synth, weights
from camkifu_b200.engine import StoneEngine
from camkifu_b200.pipeline import DetectPipeline, pinned_frames

H, W, N = 1080, 1920, 64
frames, mtx, _, _ = synth.make_clip(0, 4, H, W)
host = pinned_frames(N, H, W)
for i in range(N):
    host[i] = torch.from_numpy(frames[i % 4])
eng = StoneEngine(19)
eng.set_cnn_weights(weights.glorot_params(seed=0))
dev = torch.empty((N, H, W, 3), dtype=torch.uint8, device="cuda")
roi = eng.frame_roi(mtx, H, W)
print("roi", roi, "bytes/frame", (roi[1] - roi[0]) * (roi[3] - roi[2]) * 3)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

t = timeit(lambda: dev.copy_(host, non_blocking=True))
print("full contiguous H2D: %.2f ms  %.1f GB/s" % (t * 1e3, host.numel() / t / 1e9))
nb = [0]
def up():
    nb[0] = eng.upload_frames(host, dev, roi)
t = timeit(up)
print("ROI 2D H2D (per-frame cudaMemcpy2DAsync): %.2f ms  %.1f GB/s  -> %.0f frames/s" % (t * 1e3, nb[0] / t / 1e9, N / t))
# ROI packed on the host first? (rows contiguous) -- measure the DMA rate for one contiguous block of the same size
blk = torch.empty(nb[0], dtype=torch.uint8, pin_memory=True); dblk = torch.empty(nb[0], dtype=torch.uint8, device="cuda")
t = timeit(lambda: dblk.copy_(blk, non_blocking=True))
print("same bytes contiguous: %.2f ms  %.1f GB/s -> %.0f frames/s" % (t * 1e3, nb[0] / t / 1e9, N / t))
for sub in (8, 16, 32):
    pipe = DetectPipeline(H, W, 19, mode="neural", sub_batch=sub, engine=eng)
    t = timeit(lambda: pipe.detect(host, mtx))
    print("detect sub_batch=%d: %.2f ms -> %.0f frames/s" % (sub, t * 1e3, N / t))
g = eng.warp(dev[:N], mtx)
def comp():
    for f0 in range(0, N, 16):
        gg = eng.warp(dev[f0:f0 + 16], mtx)
        eng.cnn_forward(gg, want_softmax=False)
t = timeit(comp)
print("compute only (4 x 16 frames): %.2f ms -> %.0f frames/s" % (t * 1e3, N / t))
