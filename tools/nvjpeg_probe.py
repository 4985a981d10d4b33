#!/usr/bin/env python
"""Motion-JPEG ingest on the device (nvJPEG) against host decode (OpenCV / FFmpeg): pixel differences, board states,
frames/s through process_video.   python tools/nvjpeg_probe.py   (GPU box)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
from camkifu_b200 import synth, weights  # noqa: E402
from camkifu_b200.engine import StoneEngine  # noqa: E402
from camkifu_b200.pipeline import DetectPipeline  # noqa: E402
from camkifu_b200.video import MjpegAvi, process_video  # noqa: E402

H, W, n = 1080, 1920, 2048
frames, mtx, truth, _ = synth.make_clip_parallel(5, 64, H, W)
path = "/tmp/ckb_nvjpeg_probe.avi"
wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30, (W, H))
for i in range(n):
    wr.write(frames[i % 64])
wr.release()
eng = StoneEngine(19)
eng.set_cnn_weights(weights.glorot_params(seed=0))
print("backend:", eng.jpeg_backend(), "| file MB:", os.path.getsize(path) / 1e6)
avi = MjpegAvi(path)
assert len(avi) == n
d = torch.empty((8, H, W, 3), dtype=torch.uint8, device="cuda")
eng.jpeg_decode(avi.base_address, avi.offsets[:8], avi.sizes[:8], d)
torch.cuda.synchronize()
cap = cv2.VideoCapture(path)
diffs = []
for i in range(8):
    ok, f = cap.read()
    diffs.append(np.abs(d[i].cpu().numpy().astype(np.int16) - f.astype(np.int16)))
cap.release()
print("nvjpeg vs FFmpeg decode: max |diff| %d, mean %.3f, share of differing bytes %.3f; vs the source frame: nvjpeg mean %.3f, FFmpeg mean %.3f"
      % (max(x.max() for x in diffs), np.mean([x.mean() for x in diffs]), np.mean([(x > 0).mean() for x in diffs]),
         np.abs(d[0].cpu().numpy().astype(np.int16) - frames[0].astype(np.int16)).mean(),
         np.abs(f.astype(np.int16) - frames[7].astype(np.int16)).mean()))
pipe = DetectPipeline(H, W, 19, mode="both", sub_batch=16, engine=eng)
res = {}
for ingest, dec, vb in (("host", 8, 64), ("nvjpeg", 1, 256), ("nvjpeg", 2, 256), ("nvjpeg", 3, 256), ("nvjpeg", 4, 256),
                        ("nvjpeg", 6, 256), ("nvjpeg", 8, 256), ("nvjpeg", 8, 128)):
    for rep in range(2):        # the first pass creates the lane's decoder and the device ring; the second is the figure
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res[ingest] = process_video(path, mtx, mode="both", batch=vb, pipeline=pipe, decoders=dec, ingest=ingest)
        dt = time.perf_counter() - t0
    print("%-7s x%d batch %3d: %d frames in %.3f s = %.0f frames/s" % (ingest, dec, vb, res[ingest]["stones"].shape[0], dt, n / dt))
for k in ("km_stones", "stones", "keep"):
    a, b = res["host"][k], res["nvjpeg"][k]
    print("%-10s host vs nvjpeg ingest: %.5f of entries equal; vs ground truth: host %.5f nvjpeg %.5f"
          % (k, (a == b).mean(), (a == truth[np.arange(n) % 64]).mean(), (b == truth[np.arange(n) % 64]).mean()))
os.remove(path)
