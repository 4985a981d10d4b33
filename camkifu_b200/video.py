"""Offline video file processing: decode -> pinned host ring -> DetectPipeline.detect_stream -> per-frame board states.

The reference reads a file through `CaptureReader` (src/camkifu/core/vmanager.py:510-525,563-586): one `VideoCapture.read`
per processed frame on the finder's thread, throttled to `file_fps` = 5 frames per second of video by skipping frames
(cvconf.py:18-19). This module is the ingest side of the batch path that replaces that cadence ("fast video file
processing", README.md:35): every frame of [start, stop) is decoded by a background thread straight
into page-locked batch buffers while the previous batches upload and compute, and with several processes (one per GPU)
each rank takes a contiguous frame range (camkifu_b200.sharding) and the per-frame board states are gathered once at the
end. Decoding is OpenCV's FFmpeg reader on the host (this image has no NVDEC binding); at 1080p it, not the GPU, sets the
pace of a single process — which is why the shards matter.

Nothing here computes on the detection path: frames go to `DetectPipeline` untouched.
"""
import queue
import threading

import numpy as np
import torch

from . import sharding
from .pipeline import DetectPipeline, pinned_frames


def open_source(source):
    """(capture or array, n_frames, H, W) of a video file path or an indexable of BGR uint8 frames."""
    if isinstance(source, str):
        import cv2
        cap = cv2.VideoCapture(source)
        if not cap.isOpened():
            raise IOError("cannot open video " + source)
        return cap, int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), \
            int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    return source, len(source), int(source[0].shape[0]), int(source[0].shape[1])


class FrameSource:
    """Frames [start, stop) of a video as pinned batches, decoded by a daemon thread. `source` is a file path
    (cv2.VideoCapture) or any indexable of BGR uint8 frames (e.g. a numpy array [n, H, W, 3]). Iterating yields
    (pinned buffer [batch, H, W, 3], frames filled m, index of the first frame); the consumer hands every buffer back
    with `release()` once the batch has been consumed (its upload has completed), and the decoder blocks when all
    `depth` buffers are out."""

    def __init__(self, source, start: int = 0, stop: int = None, batch: int = 32, depth: int = 5):
        src, self.n_total, self.H, self.W = open_source(source)
        self._cap = src if isinstance(source, str) else None
        self._arr = None if isinstance(source, str) else src
        self.start = max(0, start)
        self.stop = self.n_total if stop is None else min(stop, self.n_total)
        self.batch = batch
        self._free = queue.Queue()
        for _ in range(depth):
            self._free.put(pinned_frames(batch, self.H, self.W))
        self._q = queue.Queue()
        self._err = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def __len__(self):
        return max(0, self.stop - self.start)

    def release(self, buf):
        self._free.put(buf)

    def _seek(self):
        import cv2
        cap = self._cap
        if self.start == 0:
            return
        cap.set(cv2.CAP_PROP_POS_FRAMES, self.start)
        if int(cap.get(cv2.CAP_PROP_POS_FRAMES)) != self.start:      # inexact seek (inter-frame codec): walk there
            cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
            for _ in range(self.start):
                if not cap.grab():
                    break

    def _run(self):
        try:
            if self._cap is not None:
                self._seek()
            pos = self.start
            while pos < self.stop:
                buf = self._free.get()
                m = min(self.batch, self.stop - pos)
                view = buf.numpy()
                for i in range(m):
                    if self._cap is not None:
                        ok, frame = self._cap.read()
                        if not ok:
                            raise IOError("decode failed at frame %d" % (pos + i))
                        view[i] = frame
                    else:
                        view[i] = self._arr[pos + i]
                self._q.put((buf, m, pos))
                pos += m
        except BaseException as e:   # surfaced on the consumer side
            self._err = e
        finally:
            self._q.put(None)

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is None:
                if self._err is not None:
                    raise self._err
                return
            yield item


def process_video(source, mtx, mode: str = "neural", gsize: int = 19, batch: int = 32, cnn_params=None, engine=None,
                  rng_state: int = None, rank: int = None, world: int = None, gather: bool = True, pipeline=None):
    """Board states of every frame of a video under one board homography `mtx` (a fixed camera: the reference's manual
    board finder). Returns {name: array [n_frames, ...]} — "stones"/"keep"/"conf" (neural), "km_stones"/"km_trusted"
    (clustering: full-board find_stones per frame, RNG state replayed per frame index so that the result does not depend
    on the sharding). With torch.distributed initialised (or rank/world given) each rank processes its frame range and,
    if `gather`, the states are all-gathered so that every rank returns the whole video. `pipeline`: an existing
    DetectPipeline to reuse (anything with its detect_stream / eng interface)."""
    import collections
    import torch.distributed as dist
    from .engine import rng_seed, rng_advance
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    _, n_total, H, W = open_source(source)
    start, stop = sharding.shard_range(n_total, rank, world)
    src = FrameSource(source, start, stop, batch=batch, depth=5)
    pipe = pipeline or DetectPipeline(H, W, gsize, mode=mode, sub_batch=min(16, batch), cnn_params=cnn_params, engine=engine)
    st0 = rng_seed(0) if rng_state is None else rng_state
    inflight = collections.deque()

    def batches():
        for buf, m, pos in src:
            inflight.append(buf)
            yield buf[:m], mtx, rng_advance(st0, pos)

    parts = {}
    for res in pipe.detect_stream(batches(), depth=2):
        for k, v in res.items():
            parts.setdefault(k, []).append(v.copy())
        src.release(inflight.popleft())             # results are ready, so this batch's uploads have completed
    names = {"neural": ("stones", "keep", "conf"), "clustering": ("km_stones", "km_trusted"),
             "both": ("stones", "keep", "conf", "km_stones", "km_trusted")}[mode]
    out = {}
    for k in names:
        if k in parts:
            out[k] = np.concatenate(parts[k])
        else:
            shape = (0,) if k == "km_trusted" else (0, gsize, gsize)
            out[k] = np.zeros(shape, np.float32 if k == "conf" else np.uint8)
    if gather and world > 1:
        dev = pipe.eng.device if dist.get_backend() == "nccl" else torch.device("cpu")   # gloo gathers on the host
        for k in names:
            out[k] = sharding.gather_board_states(torch.from_numpy(out[k]).to(dev), n_total).cpu().numpy()
    return out
