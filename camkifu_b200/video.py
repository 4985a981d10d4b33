"""Offline video file processing: decode -> pinned host ring -> DetectPipeline.detect_stream -> per-frame board states.

The reference reads a file through `CaptureReader` (src/camkifu/core/vmanager.py:510-525,563-586): one `VideoCapture.read`
per processed frame on the finder's thread, throttled to `file_fps` = 5 frames per second of video by skipping frames
(cvconf.py:18-19). This module is the ingest side of the batch path that replaces that cadence ("fast video file
processing", README.md:35): every frame of [start, stop) is decoded by background threads straight
into page-locked batch buffers (`VideoCapture.read(image=slot)`: no intermediate copy) while the previous batches upload
and compute; with several processes (one per GPU) each rank takes a contiguous frame range (camkifu_b200.sharding) and
the per-frame board states are gathered once at the end. Decoding is OpenCV's FFmpeg reader on the host (this image has no
NVDEC binding: see DESIGN.md for the probe); at 1080p it, not the GPU, sets the pace, which is why a rank may run several
decoder threads (`decoders`), each with its own capture over a contiguous part of the rank's range.

Frame counts. FFmpeg only estimates `CAP_PROP_FRAME_COUNT` for many containers, so a failed read is the end of the stream,
not an error: the decoder stops, the rank returns the frames it really got, and the ranks agree on the real counts before
the final gather (which therefore cannot dead-lock on a short file).

Nothing here computes on the detection path: frames go to `DetectPipeline` untouched.
"""
import queue
import threading

import numpy as np
import torch

from . import sharding
from .pipeline import DetectPipeline, pinned_frames


def open_source(source):
    """(capture or indexable, n_frames, H, W) of a video file path or an indexable of BGR uint8 frames. For a file
    n_frames is the container's estimate."""
    if isinstance(source, str):
        import cv2
        cap = cv2.VideoCapture(source)
        if not cap.isOpened():
            raise IOError("cannot open video " + source)
        return cap, int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), \
            int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    return source, len(source), int(source[0].shape[0]), int(source[0].shape[1])


def probe(source):
    """(n_frames, H, W) without keeping a capture open."""
    src, n, H, W = open_source(source)
    if isinstance(source, str):
        src.release()
    return n, H, W


class RingClip:
    """A long synthetic video held in host memory as a short ring: frame i is ring[i % len(ring)]. `ring` is a pinned
    torch tensor [k, H, W, 3]; FrameSource hands out views of it (no copy), so this source measures the path behind the
    decoder. Batches never straddle the end of the ring."""

    def __init__(self, ring: torch.Tensor, n_frames: int):
        assert ring.dim() == 4 and ring.dtype == torch.uint8
        self.ring, self.n = ring, int(n_frames)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.ring[i % self.ring.shape[0]].numpy()

    def view(self, pos: int, m: int):
        k = self.ring.shape[0]
        a = pos % k
        m = min(m, k - a)
        return self.ring[a:a + m], m


class MjpegAvi:
    """Index of a Motion-JPEG AVI: where each frame's JPEG sits in the (memory-mapped) file. RIFF walk: the video chunks
    ('..dc' / '..db') of every 'movi' list ('rec ' groups and OpenDML 'AVIX' segments included). Raises ValueError if the
    file is not an AVI whose first video chunk is a JPEG — other containers / codecs go through FrameSource."""

    def __init__(self, path: str):
        import struct
        self.path = path
        self.data = np.memmap(path, dtype=np.uint8, mode="r")
        buf = self.data
        size = buf.shape[0]
        if size < 12 or bytes(buf[0:4]) != b"RIFF" or bytes(buf[8:12]) != b"AVI ":
            raise ValueError(path + " is not an AVI file")
        offs, lens = [], []

        def walk(pos, end):
            while pos + 8 <= end:
                cc = bytes(buf[pos:pos + 4])
                ln = struct.unpack("<I", bytes(buf[pos + 4:pos + 8]))[0]
                body = pos + 8
                if cc in (b"RIFF", b"LIST"):
                    kind = bytes(buf[body:body + 4])
                    if cc == b"RIFF" or kind in (b"movi", b"rec "):
                        walk(body + 4, min(end, body + ln))
                elif cc[2:4] in (b"dc", b"db") and ln > 0:
                    offs.append(body)
                    lens.append(ln)
                pos = body + ln + (ln & 1)

        walk(0, size)
        if not offs or bytes(buf[offs[0]:offs[0] + 2]) != b"\xff\xd8":
            raise ValueError(path + " holds no Motion-JPEG video chunks")
        self.offsets = np.asarray(offs, dtype=np.int64)
        self.sizes = np.asarray(lens, dtype=np.int64)
        self.base_address = int(self.data.ctypes.data)
        self.H, self.W = self._frame_size(bytes(buf[offs[0]:offs[0] + min(lens[0], 65536)]))

    @staticmethod
    def _frame_size(jpeg: bytes):
        """(height, width) from the frame header (SOF0 / SOF1 / SOF2 marker) of a JPEG."""
        i = 2
        while i + 9 < len(jpeg):
            if jpeg[i] != 0xFF:
                i += 1
                continue
            m = jpeg[i + 1]
            if m in (0xC0, 0xC1, 0xC2):
                return (jpeg[i + 5] << 8) | jpeg[i + 6], (jpeg[i + 7] << 8) | jpeg[i + 8]
            if m == 0xFF or 0xD0 <= m <= 0xD9 or m == 0x01:
                i += 2 if m != 0xFF else 1
                continue
            i += 2 + ((jpeg[i + 2] << 8) | jpeg[i + 3])
        raise ValueError("no JPEG frame header found")

    def __len__(self):
        return len(self.offsets)

    def repeat(self, k: int):
        """The same frames k times over, as one long video (an index, no bytes are copied): benchmark input."""
        import copy
        out = copy.copy(self)
        out.offsets, out.sizes = np.tile(self.offsets, k), np.tile(self.sizes, k)
        return out


class FrameSource:
    """Frames [start, stop) of a video as pinned batches. `source` is a file path (cv2.VideoCapture), any indexable of
    BGR uint8 frames (e.g. a numpy array [n, H, W, 3]) or a RingClip. Iterating yields (pinned buffer [batch, H, W, 3],
    frames filled m, index of the first frame); the consumer hands every buffer back with `release()` once the batch has
    been consumed (its upload has completed), and a decoder blocks when all its `depth` buffers are out.

    decoders = 1: batches come in frame order. decoders = k > 1: k threads, each with its own capture over a contiguous
    part of the range; batches are yielded as they complete, so their order interleaves the parts (every batch carries
    its first frame index) — for consumers whose result does not depend on the order of the frames."""

    def __init__(self, source, start: int = 0, stop: int = None, batch: int = 32, depth: int = 5, decoders: int = 1):
        self.source = source
        self.n_total, self.H, self.W = probe(source)
        self.start = max(0, start)
        self.stop = self.n_total if stop is None else min(stop, self.n_total)
        self.batch = batch
        self.frames_read = 0                 # frames really delivered (a file may be shorter than its header says)
        self._ring = source if isinstance(source, RingClip) else None
        self._is_file = isinstance(source, str)
        self._q = queue.Queue()
        self._err = None
        self._threads = []
        self._free = {}
        n = max(0, self.stop - self.start)
        k = max(1, min(decoders, (n + batch - 1) // batch)) if (self._is_file and n) else 1
        per = ((n + k - 1) // k + batch - 1) // batch * batch if n else 0      # whole batches per decoder
        self._pending = k
        self._lock = threading.Lock()
        for j in range(k):
            a = self.start + j * per
            b = min(self.stop, a + per)
            fq = queue.Queue()
            if self._ring is None:
                for _ in range(depth):
                    buf = pinned_frames(batch, self.H, self.W)
                    self._free[buf.data_ptr()] = fq
                    fq.put(buf)
            t = threading.Thread(target=self._run, args=(a, b, fq), daemon=True)
            self._threads.append(t)
            t.start()

    def __len__(self):
        return max(0, self.stop - self.start)

    def release(self, buf):
        fq = self._free.get(buf.data_ptr())
        if fq is not None:                   # views of a RingClip are not recycled
            fq.put(buf)

    @staticmethod
    def _seek(cap, start):
        import cv2
        if start == 0:
            return
        cap.set(cv2.CAP_PROP_POS_FRAMES, start)
        if int(cap.get(cv2.CAP_PROP_POS_FRAMES)) != start:      # inexact seek (inter-frame codec): walk there
            cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
            for _ in range(start):
                if not cap.grab():
                    break

    def _run(self, a, b, fq):
        cap = None
        try:
            if self._is_file:
                import cv2
                cap = cv2.VideoCapture(self.source)
                if not cap.isOpened():
                    raise IOError("cannot open video " + self.source)
                self._seek(cap, a)
            pos = a
            while pos < b:
                m = min(self.batch, b - pos)
                if self._ring is not None:
                    buf, m = self._ring.view(pos, m)
                    self._q.put((buf, m, pos))
                    pos += m
                    continue
                buf = fq.get()
                view = buf.numpy()
                got = 0
                for i in range(m):
                    if cap is not None:
                        dst = view[i]
                        ok, frame = cap.read(dst)              # decode straight into the pinned slot
                        if not ok:
                            break                              # end of the stream (the header's count was an estimate)
                        if frame.ctypes.data != dst.ctypes.data:
                            dst[...] = frame                   # OpenCV allocated its own image: copy it in
                    else:
                        try:
                            view[i] = self.source[pos + i]
                        except IndexError:                     # shorter than len() claimed: same rule as for files
                            break
                    got += 1
                if got:
                    self._q.put((buf, got, pos))
                else:
                    fq.put(buf)
                pos += got
                if got < m:
                    break
        except BaseException as e:   # surfaced on the consumer side
            self._err = e
        finally:
            if cap is not None:
                cap.release()
            with self._lock:
                self._pending -= 1
                last = self._pending == 0
            if last:
                self._q.put(None)

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is None:
                if self._err is not None:
                    raise self._err
                return
            self.frames_read += item[1]
            yield item


def process_video(source, mtx, mode: str = "neural", gsize: int = 19, batch: int = 32, cnn_params=None, engine=None,
                  rng_state: int = None, rank: int = None, world: int = None, gather: bool = True, pipeline=None,
                  decoders: int = 1, depth: int = 3, stats: dict = None, ingest: str = "host"):
    """Board states of every frame of a video under one board homography `mtx` (a fixed camera: the reference's manual
    board finder). Returns {name: array [n_frames, ...]} — "stones"/"keep"/"conf" (neural), "km_stones"/"km_trusted"
    (clustering: full-board find_stones per frame, RNG state replayed per frame index so that the result does not depend
    on the sharding or on the order in which batches are decoded). With torch.distributed initialised (or rank/world
    given) each rank processes its frame range and, if `gather` and a process group exists, the states are all-gathered
    so that every rank returns the whole video; a file shorter than its header claims simply yields fewer frames.
    `pipeline`: an existing DetectPipeline to reuse (anything with its detect_stream / eng interface). `decoders`: decoder
    threads of this rank, each with `depth` pinned batch buffers (page-locking memory is slow: keep batch x depth x
    decoders modest, e.g. 16 x 3 x 8 frames of 1080p = 2.4 GB) (file sources; the streaming mode "full" needs frame order and uses one). `stats`, if given,
    receives {"frames": frames this rank processed, "range": (start, stop)}.
    ingest = "nvjpeg" (Motion-JPEG AVI files only, stateless modes): the compressed frames cross PCIe and are decoded on
    the device by nvJPEG straight into the buffer the warp reads (csrc/jpeg_ingest.cu) — no host decode, a tenth of the
    upload; the decoded pixels can differ from FFmpeg's by a level or two, hence not the default. `decoders` is then the
    number of nvJPEG lanes (threads + streams) decoding batches side by side; use batches of 256 frames (nvJPEG decodes
    batches of less than ~100 images partly on the host) — the ring holds decoders + 3 such batches on the device."""
    import collections
    import torch.distributed as dist
    from .engine import rng_seed, rng_advance
    have_group = dist.is_available() and dist.is_initialized()
    if rank is None or world is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if have_group else (0, 1)
    avi = None
    if ingest == "nvjpeg":
        if not isinstance(source, (str, MjpegAvi)) or mode == "full":
            raise ValueError('ingest="nvjpeg" needs a Motion-JPEG AVI file (path or MjpegAvi) and a stateless mode')
        avi = source if isinstance(source, MjpegAvi) else MjpegAvi(source)
        n_total, H, W = len(avi), avi.H, avi.W      # the index is exact (the header's frame count is an estimate)
    else:
        n_total, H, W = probe(source)
    if have_group and world > 1:       # every rank must shard the same count: rank 0's view of the file wins
        obj = [n_total]
        dist.broadcast_object_list(obj, src=0)
        n_total = int(obj[0])
    start, stop = sharding.shard_range(n_total, rank, world)
    pipe = pipeline or DetectPipeline(H, W, gsize, mode=mode, sub_batch=min(16, batch), cnn_params=cnn_params, engine=engine)
    st0 = rng_seed(0) if rng_state is None else rng_state
    inflight = collections.deque()
    if avi is None:
        src = FrameSource(source, start, stop, batch=batch, depth=depth, decoders=1 if mode == "full" else decoders)

        def batches():
            for buf, m, pos in src:
                inflight.append((buf, m, pos))
                yield buf[:m], mtx, rng_advance(st0, pos)
    else:
        # `decoders` nvJPEG lanes, each a host thread with its own CUDA stream: a batch keeps the GPU busy for a fraction of
        # its decode time only, so the lanes overlap (csrc/jpeg_ingest.cu). Batch b goes to lane b % lanes and into ring
        # slot b % slots once batch b - slots has been consumed; the pipeline takes the batches in order.
        import threading
        eng = pipe.eng
        lanes = max(1, min(int(decoders), eng.JPEG_LANES))
        positions = list(range(start, stop, batch))
        slots = lanes + pipe.DEPTH + 1
        ring = [torch.empty((batch, H, W, 3), dtype=torch.uint8, device=eng.device) for _ in range(min(slots, max(1, len(positions))))]
        slots = len(ring)
        cond = threading.Condition()
        ready, state = {}, {"consumed": 0, "error": None}
        eng.jpeg_backend()                               # creates lane 0 / loads the library before the threads start

        def lane_main(k):
            try:
                torch.cuda.set_device(eng.device)
                stream = torch.cuda.Stream(device=eng.device)
                for bi in range(k, len(positions), lanes):
                    with cond:
                        cond.wait_for(lambda: state["consumed"] > bi - slots or state["error"] is not None)
                        if state["error"] is not None:
                            return
                    pos = positions[bi]
                    m = min(batch, stop - pos)
                    with torch.cuda.stream(stream):
                        eng.jpeg_decode(avi.base_address, avi.offsets[pos:pos + m], avi.sizes[pos:pos + m], ring[bi % slots],
                                        cpu_threads=2, lane=k)
                        ev = torch.cuda.Event()
                        ev.record(stream)
                    with cond:
                        ready[bi] = ev
                        cond.notify_all()
            except BaseException as e:                   # noqa: BLE001 - handed to the consumer
                with cond:
                    state["error"] = e
                    cond.notify_all()

        threads = [threading.Thread(target=lane_main, args=(k,), daemon=True) for k in range(lanes)]
        for t in threads:
            t.start()

        class _Ring:
            @staticmethod
            def release(buf):                            # the batch's results are on the host: its slot may be decoded into
                with cond:
                    state["consumed"] += 1
                    cond.notify_all()
        src = _Ring

        def batches():
            try:
                for bi, pos in enumerate(positions):
                    with cond:
                        cond.wait_for(lambda: bi in ready or state["error"] is not None)
                        if state["error"] is not None:
                            raise state["error"]
                        ev = ready.pop(bi)
                    pipe.comp_stream.wait_event(ev)      # the pipeline's kernels run after the decode of this batch
                    m = min(batch, stop - pos)
                    inflight.append((None, m, pos))
                    yield ring[bi % slots][:m], mtx, rng_advance(st0, pos)
            finally:                                     # abandoned half way (an error downstream): let the lanes go
                with cond:
                    if state["error"] is None:
                        state["error"] = RuntimeError("ingest stopped")
                    cond.notify_all()

    names = {"neural": ("stones", "keep", "conf"), "clustering": ("km_stones", "km_trusted"),
             "both": ("stones", "keep", "conf", "km_stones", "km_trusted"),
             "full": ("stones", "keep", "conf", "km_stones", "km_trusted", "fg_counts")}[mode]
    shapes = {"km_trusted": (), "fg_counts": (gsize, gsize)}
    dtypes = {"conf": np.float32, "fg_counts": np.int32}
    out = {k: np.zeros((max(0, stop - start),) + shapes.get(k, (gsize, gsize)), dtypes.get(k, np.uint8)) for k in names}
    done = 0
    for res in pipe.detect_stream(batches(), depth=2):
        buf, m, pos = inflight.popleft()
        for k in names:
            out[k][pos - start:pos - start + m] = res[k]
        done = max(done, pos - start + m)
        src.release(buf)                            # results are ready, so this batch's uploads have completed
    for k in names:                                 # a short file: keep what was really read
        out[k] = out[k][:done]
    if stats is not None:
        stats.update(frames=done, range=(start, stop))
    if gather and world > 1 and have_group:
        dev = pipe.eng.device if dist.get_backend() == "nccl" else torch.device("cpu")   # gloo gathers on the host
        for k in names:
            out[k] = sharding.gather_ragged(torch.from_numpy(out[k]).to(dev)).cpu().numpy()
    return out
