"""Offline batch stone detection: host frames in, board states out ("fast video file processing", README.md:35).

The reference handles one frame per `_doframe` call on a Python thread (core/video.py:88-120) and throttles file input
to 5 fps (vmanager.py:510-525). This module is the batch entry point that bypasses that cadence: frames of one video
segment (one board homography) are staged host -> device in sub-batches on a copy stream while the previous sub-batch
is warped and classified on the compute stream; only the part of each frame the warp can read is uploaded.

Everything that computes is a kernel of libcamkifu_b200.so; torch provides buffers, streams and events.
"""
import numpy as np
import torch

from .engine import StoneEngine, rng_seed, rng_states


def pinned_frames(n: int, H: int, W: int) -> torch.Tensor:
    """Page-locked host buffer for n BGR frames (what the capture thread should decode into). Without a CUDA device
    (host-side tests of the ingest code) the buffer is ordinary memory."""
    return torch.empty((n, H, W, 3), dtype=torch.uint8, pin_memory=torch.cuda.is_available())


class DetectPipeline:
    """mode: "neural" (SfNeural.predict_all per frame), "clustering" (SfClustering.find_stones, full board, per frame),
    "both", or "full" = everything a finder runs per frame (BASELINE.json configs[2]): warp, the MOG2 background model
    with the reference's learning-rate schedule and the per-zone foreground counts (stonesfinder.py:171-176,
    sf_neural.py:178-180), SfClustering's running average (sf_clustering.py:33-36), full-board find_stones and
    predict_all. The streaming state of "full" (background model, running average, frame count) lives in the pipeline
    and advances with every submitted frame. Results are numpy arrays in pinned host memory, valid until DEPTH more
    batches have been submitted. `h2d_bytes` / `d2h_bytes` count everything copied since construction."""

    def __init__(self, H: int, W: int, gsize: int = 19, mode: str = "neural", sub_batch: int = 16, device=None,
                 cnn_params=None, engine: StoneEngine = None):
        assert mode in ("neural", "clustering", "both", "full")
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
        self.side_stream = torch.cuda.Stream(device=dev)      # mode "full": the statistics branch next to the CNN branch
        self.ev_warped, self.ev_side = torch.cuda.Event(), torch.cuda.Event()
        self.ev_up = [torch.cuda.Event() for _ in range(2)]
        self.ev_free = [torch.cuda.Event() for _ in range(2)]
        self._k = 0                     # sub-batches enqueued so far (the device frame buffers are a ring of two)
        self._slots = []                # output slots (pinned host buffers + completion event), a ring of DEPTH
        self._ticket = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        if mode == "full":
            self.bg_state = self.eng.mog2_new_state()
            self.accu = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
            self.d_fg = torch.empty((sub_batch, S, S), dtype=torch.uint8, device=dev)
            self.frames_seen = 0

    DEPTH = 3   # batches that may be in flight (submit() without collect())

    def _slot(self, ticket, n):
        g = self.eng.gsize
        while len(self._slots) < self.DEPTH:
            self._slots.append({"cap": 0, "out": {}, "done": torch.cuda.Event(), "ticket": -1, "n": 0})
        sl = self._slots[ticket % self.DEPTH]
        if sl["ticket"] >= 0 and sl["ticket"] != ticket:
            sl["done"].synchronize()    # the slot's previous batch must have been produced before it is reused
        if n > sl["cap"]:
            out = {}
            if self.mode != "clustering":
                out["stones"] = torch.empty((n, g, g), dtype=torch.uint8, pin_memory=True)
                out["keep"] = torch.empty((n, g, g), dtype=torch.uint8, pin_memory=True)
                out["conf"] = torch.empty((n, g, g), dtype=torch.float32, pin_memory=True)
            if self.mode != "neural":
                out["km_stones"] = torch.empty((n, g, g), dtype=torch.uint8, pin_memory=True)
                out["km_trusted"] = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
                sl["h_states"] = torch.empty((n,), dtype=torch.int64, pin_memory=True)     # cv::RNG states, staged pinned
                sl["d_states"] = torch.empty((n,), dtype=torch.int64, device=self.eng.device)
            if self.mode == "full":
                out["fg_counts"] = torch.empty((n, g, g), dtype=torch.int32, pin_memory=True)
            sl["out"], sl["cap"] = out, n
        sl["ticket"], sl["n"] = ticket, n
        return sl

    def submit(self, frames: torch.Tensor, mtx, rng_state: int = None, crop: bool = True) -> int:
        """Enqueue one batch: frames HOST uint8 [n, H, W, 3] (torch tensor, ideally pinned; numpy arrays are wrapped),
        mtx the 3x3 frame -> canonical homography of the segment, rng_state the cv::RNG state before the first k-means
        call (clustering modes; defaults to cv2.setRNGSeed(0)'s). Returns a ticket for collect(). The call only
        enqueues copies and kernels: the upload of this batch overlaps the kernels of the previous one. `frames` must
        stay untouched until the batch is collected; at most DEPTH batches may be outstanding."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        on_device = frames.is_cuda      # frames already in HBM (e.g. decoded there): no upload, the warp reads them in place;
        n = frames.shape[0]             # whatever produced them must be ordered before comp_stream (enqueued on it, or waited for)
        assert tuple(frames.shape[1:]) == (self.H, self.W, 3)
        ticket = self._ticket
        self._ticket += 1
        sl = self._slot(ticket, n)
        out = sl["out"]
        eng = self.eng
        roi = eng.frame_roi(mtx, self.H, self.W) if (crop and not on_device) else None
        st0 = rng_seed(0) if rng_state is None else rng_state
        if self.mode != "neural":
            sl["h_states"][:n].copy_(torch.from_numpy(np.asarray(rng_states(st0, 0, n), dtype=np.uint64).astype(np.int64)))
        if ticket == 0 or self._idle:
            cur = torch.cuda.current_stream(eng.device)
            self.copy_stream.wait_stream(cur)
            self.comp_stream.wait_stream(cur)
            self._idle = False
        for f0 in range(0, n, self.sub):
            m = min(self.sub, n - f0)
            k = self._k
            self._k += 1
            b = k & 1
            if not on_device:
                with torch.cuda.stream(self.copy_stream):
                    if k >= 2:
                        self.copy_stream.wait_event(self.ev_free[b])
                    self.h2d_bytes += eng.upload_frames(frames[f0:f0 + m], self.d_frames[b], roi)
                    self.ev_up[b].record(self.copy_stream)
            with torch.cuda.stream(self.comp_stream):
                if on_device:
                    goban = eng.warp(frames[f0:f0 + m], mtx, out=self.d_goban[:m])
                else:
                    self.comp_stream.wait_event(self.ev_up[b])
                    goban = eng.warp(self.d_frames[b][:m], mtx, out=self.d_goban[:m])
                self.ev_free[b].record(self.comp_stream)
                # the statistics branch (background model, running average, k-means) and the CNN branch both only read
                # the warped images: in mode "full" the first runs on a side stream next to the second
                stats_stream = self.side_stream if self.mode == "full" else self.comp_stream
                if self.mode == "full":
                    self.ev_warped.record(self.comp_stream)
                    self.side_stream.wait_event(self.ev_warped)
                with torch.cuda.stream(stats_stream):
                    if self.mode == "full":
                        t0 = self.frames_seen     # stonesfinder.py:171-176: rate 0.01 while learning (bg_init_frames = 50), then 0.005
                        fg = eng.mog2_apply(goban, self.bg_state, t0, [0.01 if t0 + i < 50 else 0.005 for i in range(m)],
                                            out=self.d_fg[:m])
                        cnt = eng.zone_fg_counts(fg)
                        out["fg_counts"][f0:f0 + m].copy_(cnt, non_blocking=True)
                        self.d2h_bytes += cnt.numel() * 4
                        eng.accumulate(goban, self.accu, first=(t0 == 0))
                        self.frames_seen = t0 + m
                    if self.mode != "neural":
                        if f0 == 0:
                            sl["d_states"][:n].copy_(sl["h_states"][:n], non_blocking=True)
                        r = eng.find_stones(goban, sl["d_states"][f0:f0 + m])
                        out["km_stones"][f0:f0 + m].copy_(r["stones"], non_blocking=True)
                        out["km_trusted"][f0:f0 + m].copy_(r["trusted"], non_blocking=True)
                        self.d2h_bytes += r["stones"].numel() + r["trusted"].numel()
                    if self.mode == "full":
                        self.ev_side.record(self.side_stream)
                if self.mode != "clustering":
                    r = eng.cnn_forward(goban, want_softmax=False)
                    for name in ("stones", "keep", "conf"):
                        out[name][f0:f0 + m].copy_(r[name], non_blocking=True)
                        self.d2h_bytes += r[name].numel() * r[name].element_size()
                if self.mode == "full":
                    self.comp_stream.wait_event(self.ev_side)      # the next warp overwrites the images both branches read
        sl["done"].record(self.comp_stream)
        return ticket

    _idle = True

    def collect(self, ticket: int):
        """Wait for a submitted batch; returns {name: numpy array [n, ...]} (views of pinned host buffers, valid until
        DEPTH more batches have been submitted)."""
        sl = self._slots[ticket % self.DEPTH]
        if sl["ticket"] != ticket:
            raise ValueError("ticket %d is no longer available (at most %d batches may be outstanding)" % (ticket, self.DEPTH))
        sl["done"].synchronize()
        return {k: v[:sl["n"]].numpy() for k, v in sl["out"].items()}

    def detect(self, frames: torch.Tensor, mtx, rng_state: int = None, crop: bool = True):
        """Synchronous form: submit + collect of one batch (see submit)."""
        res = self.collect(self.submit(frames, mtx, rng_state, crop))
        cur = torch.cuda.current_stream(self.eng.device)
        cur.wait_stream(self.comp_stream)
        self._idle = True
        return res

    def detect_stream(self, batches, depth: int = 2):
        """Offline video: `batches` yields (frames, mtx) or (frames, mtx, rng_state); results are yielded in order while
        up to `depth` (<= DEPTH - 1) later batches are already uploading / computing."""
        depth = max(1, min(depth, self.DEPTH - 1))
        pending = []
        for item in batches:
            pending.append(self.submit(*item))
            if len(pending) > depth:
                yield self.collect(pending.pop(0))
        while pending:
            yield self.collect(pending.pop(0))
