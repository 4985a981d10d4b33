"""Offline batch stone detection: host frames in, board states out ("fast video file processing", README.md:35).

The reference handles one frame per `_doframe` call on a Python thread (core/video.py:88-120) and throttles file input
to 5 fps (vmanager.py:510-525). This module is the batch entry point that bypasses that cadence: frames of one video
segment (one board homography) are staged host -> device in sub-batches on a copy stream while the previous sub-batch
is warped and classified on the compute stream; only the part of each frame the warp can read is uploaded.

Everything that computes is a kernel of libcamkifu_b200.so; torch provides buffers, streams and events.
"""
import numpy as np
import torch

from .engine import StoneEngine, rng_seed, rng_advance


def pinned_frames(n: int, H: int, W: int) -> torch.Tensor:
    """Page-locked host buffer for n BGR frames (what the capture thread should decode into)."""
    return torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=True)


class DetectPipeline:
    """mode: "neural" (SfNeural.predict_all per frame), "clustering" (SfClustering.find_stones, full board, per frame)
    or "both". Results are numpy arrays in pinned host memory, valid until the next call."""

    def __init__(self, H: int, W: int, gsize: int = 19, mode: str = "neural", sub_batch: int = 16, device=None,
                 cnn_params=None, engine: StoneEngine = None):
        assert mode in ("neural", "clustering", "both")
        self.eng = engine or StoneEngine(gsize, device=device)
        self.H, self.W, self.mode, self.sub = H, W, mode, sub_batch
        dev = self.eng.device
        if mode != "clustering":
            if cnn_params is not None:
                self.eng.set_cnn_weights(cnn_params)
            if not self.eng.has_weights:
                raise ValueError("the neural mode needs CNN parameters (see camkifu_b200.weights)")
        g, S = gsize, 20 * gsize
        self.d_frames = [torch.empty((sub_batch, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        self.d_goban = torch.empty((sub_batch, S, S, 3), dtype=torch.uint8, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.comp_stream = torch.cuda.Stream(device=dev)
        self.ev_up = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self._cap = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _ensure_out(self, n):
        if n <= self._cap:
            return
        g = self.eng.gsize
        self.out = {}
        if self.mode != "clustering":
            self.out["stones"] = torch.empty((n, g, g), dtype=torch.uint8, pin_memory=True)
            self.out["keep"] = torch.empty((n, g, g), dtype=torch.uint8, pin_memory=True)
            self.out["conf"] = torch.empty((n, g, g), dtype=torch.float32, pin_memory=True)
        if self.mode != "neural":
            self.out["km_stones"] = torch.empty((n, g, g), dtype=torch.uint8, pin_memory=True)
            self.out["km_trusted"] = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
        self._cap = n

    def detect(self, frames: torch.Tensor, mtx, rng_state: int = None, crop: bool = True):
        """frames: HOST uint8 [n, H, W, 3] (torch tensor, ideally pinned; numpy arrays are wrapped), mtx: the 3x3
        frame -> canonical homography of the segment. rng_state: cv::RNG state before the first k-means call
        (clustering modes; defaults to cv2.setRNGSeed(0)'s). Returns {name: numpy array [n, ...]}."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        n = frames.shape[0]
        assert tuple(frames.shape[1:]) == (self.H, self.W, 3)
        self._ensure_out(n)
        eng = self.eng
        roi = eng.frame_roi(mtx, self.H, self.W) if crop else None
        st0 = rng_seed(0) if rng_state is None else rng_state
        self.h2d_bytes = self.d2h_bytes = 0
        cur = torch.cuda.current_stream(eng.device)
        self.copy_stream.wait_stream(cur)
        self.comp_stream.wait_stream(cur)
        for k, f0 in enumerate(range(0, n, self.sub)):
            m = min(self.sub, n - f0)
            b = k & 1
            with torch.cuda.stream(self.copy_stream):
                if k >= 2:
                    self.copy_stream.wait_event(self.ev_free[b])
                self.h2d_bytes += eng.upload_frames(frames[f0:f0 + m], self.d_frames[b], roi)
                self.ev_up[b].record(self.copy_stream)
            with torch.cuda.stream(self.comp_stream):
                self.comp_stream.wait_event(self.ev_up[b])
                goban = eng.warp(self.d_frames[b][:m], mtx, out=self.d_goban[:m])
                self.ev_free[b].record(self.comp_stream)
                if self.mode != "clustering":
                    r = eng.cnn_forward(goban, want_softmax=False)
                    for name in ("stones", "keep", "conf"):
                        self.out[name][f0:f0 + m].copy_(r[name], non_blocking=True)
                        self.d2h_bytes += r[name].numel() * r[name].element_size()
                if self.mode != "neural":
                    states = [rng_advance(st0, f0 + i) for i in range(m)]
                    r = eng.find_stones(goban, states)
                    self.out["km_stones"][f0:f0 + m].copy_(r["stones"], non_blocking=True)
                    self.out["km_trusted"][f0:f0 + m].copy_(r["trusted"], non_blocking=True)
                    self.d2h_bytes += r["stones"].numel() + r["trusted"].numel()
        cur.wait_stream(self.comp_stream)
        cur.wait_stream(self.copy_stream)
        cur.synchronize()
        return {k: v[:n].numpy() for k, v in self.out.items()}
