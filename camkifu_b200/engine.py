"""StoneEngine — host-side driver of the CUDA stone-detection path.

Thin layer over the C ABI (include/camkifu_b200.h): PyTorch provides device buffers and the CUDA stream, every
computation is a hand-written sm_100a kernel inside libcamkifu_b200.so. One engine = one `ckb_ctx` = one finder
instance; it is safe to use from the finder's own thread (ctypes releases the GIL during calls).

Colour codes: 0 = E, 1 = B, 2 = W (the reference's Golib constants 'E', 'B', 'W').
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CkbError

DRAWS_PER_KMEANS = 39  # CKB_RNG_DRAWS_PER_KMEANS


def rng_seed(seed: int) -> int:
    """cv::RNG state installed by cv2.setRNGSeed(seed)."""
    return int(_lib.lib().ckb_rng_seed(seed & 0xffffffff))


# cv::RNG is a multiply-with-carry generator: with t = carry * 2^32 + x (the 64-bit state itself), one draw is
# t' = A * x + carry = A * t mod (A * 2^32 - 1). n draws are therefore one modular power, whatever n.
_MWC_A = 4164903690
_MWC_M = _MWC_A * (1 << 32) - 1


def rng_advance(state: int, n_kmeans_calls: int) -> int:
    """State after `n_kmeans_calls` cv2.kmeans(K=3, attempts=3, KMEANS_PP_CENTERS) calls (39 draws each). O(log n):
    equal to ckb_rng_advance(state, 39 n), which steps draw by draw (tests/test_abi.py checks the two against each other)."""
    if n_kmeans_calls <= 0:
        return int(state)
    t = (int(state) * pow(_MWC_A, DRAWS_PER_KMEANS * int(n_kmeans_calls), _MWC_M)) % _MWC_M
    return t if t != 0 else _MWC_M


def rng_states(state: int, first_call: int, n: int):
    """States before k-means calls first_call, first_call + 1, ... (n of them) of a stream that starts at `state`."""
    step = pow(_MWC_A, DRAWS_PER_KMEANS, _MWC_M)
    t = rng_advance(state, first_call)
    out = []
    for _ in range(n):
        out.append(t)
        t = (t * step) % _MWC_M or _MWC_M
    return out


class StoneEngine:
    def __init__(self, gsize: int = 19, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("camkifu_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        if self.device.index is None:       # torch.device("cuda"): the context must bind to the device tensors live on
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.gsize = gsize
        self.S = 20 * gsize
        h = C.c_void_p()
        rc = self.L.ckb_create(C.byref(h), self.device.index, gsize)
        self._h = h
        if rc != 0:
            msg = self.L.ckb_last_error(h).decode() if h else "ckb_create failed"
            if h:
                self.L.ckb_destroy(h)
            self._h = None
            raise CkbError(rc, msg)
        self._work = None
        self.has_weights = False

    # ------------------------------------------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.L.ckb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CkbError(rc, self.L.ckb_last_error(self._h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, nbytes: int, key: str = "cnn") -> torch.Tensor:
        """Scratch buffers, one per kernel family: the k-means and the CNN branches of a frame batch may run on two
        streams at once (DetectPipeline, bench.py), so they must not share one."""
        if self._work is None:
            self._work = {}
        w = self._work.get(key)
        if w is None or w.numel() < nbytes:
            w = self._work[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return w

    @property
    def launches(self) -> int:
        return int(self.L.ckb_launch_count(self._h))

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    # ------------------------------------------------------------------------------------------------------------ K1
    def warp(self, frames: torch.Tensor, mtx, out: torch.Tensor = None) -> torch.Tensor:
        """frames: uint8 [n, H, W, 3] (or [H, W, 3]) on the device; mtx: 3x3 or [n, 3, 3] float64 frame->canonical
        homography (BoardFinder.mtx). Returns uint8 [n, S, S, 3] — cv2.warpPerspective(frame, mtx, (S, S)) per frame."""
        if frames.dim() == 3:
            frames = frames.unsqueeze(0)
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.shape[-1] == 3
        assert frames.stride(3) == 1 and frames.stride(2) == 3, "frames must be interleaved BGR"
        n, H, W, _ = frames.shape
        m = np.ascontiguousarray(np.asarray(mtx, dtype=np.float64)).reshape(-1, 9)
        if out is None:
            out = torch.empty((n, self.S, self.S, 3), dtype=torch.uint8, device=self.device)
        self._check(self.L.ckb_warp(self._h, self._ptr(frames), n, H, W, frames.stride(1), frames.stride(0) if n > 1
                                    else frames.stride(1) * H, m.ctypes.data_as(C.c_void_p), m.shape[0], self._ptr(out),
                                    self._stream()))
        return out

    def accumulate(self, goban: torch.Tensor, accu: torch.Tensor, first: bool, alpha: float = 0.2,
                   snap_every: int = 0, snap_phase: int = 0):
        """Running average of sf_clustering.py:33-36 applied to the n images in order. Returns the snapshots tensor
        [k, S, S, 3] float32 for frames i with (i + snap_phase) % snap_every == 0 (None when snap_every == 0)."""
        if goban.dim() == 3:
            goban = goban.unsqueeze(0)
        n = goban.shape[0]
        assert goban.is_contiguous() and accu.is_contiguous() and accu.dtype == torch.float32
        snaps = None
        if snap_every > 0:
            k = sum(1 for i in range(n) if (i + snap_phase) % snap_every == 0)
            snaps = torch.empty((k, self.S, self.S, 3), dtype=torch.float32, device=self.device)
        self._check(self.L.ckb_accumulate(self._h, self._ptr(goban), n, self._ptr(accu), alpha, int(first),
                                          self._ptr(snaps), max(snap_every, 1), snap_phase, self._stream()))
        return snaps

    # ---------------------------------------------------------------------------------------------- background model
    def mog2_new_state(self) -> torch.Tensor:
        """Device state of a fresh cv2.createBackgroundSubtractorMOG2(detectShadows=False) (stonesfinder.py:113-115)."""
        st = torch.empty(int(self.L.ckb_mog2_state_bytes(self._h)), dtype=torch.uint8, device=self.device)
        self._check(self.L.ckb_mog2_reset(self._h, self._ptr(st), self._stream()))
        return st

    def mog2_apply(self, goban: torch.Tensor, state: torch.Tensor, frames_before: int, learning_rates, out=None):
        """bg_model.apply(goban_img, learningRate=lr) for n canonical images in order (stonesfinder.py:171-176).
        frames_before: frames this model has already seen. Returns the foreground masks uint8 [n, S, S] (0 / 255)."""
        if goban.dim() == 3:
            goban = goban.unsqueeze(0)
        n = goban.shape[0]
        assert goban.is_contiguous() and goban.dtype == torch.uint8 and goban.shape[1:] == (self.S, self.S, 3)
        lr = np.ascontiguousarray(np.broadcast_to(np.asarray(learning_rates, dtype=np.float64), (n,)))
        if out is None:
            out = torch.empty((n, self.S, self.S), dtype=torch.uint8, device=self.device)
        self._check(self.L.ckb_mog2_apply(self._h, self._ptr(goban), n, self._ptr(state), int(frames_before),
                                          lr.ctypes.data_as(C.c_void_p), self._ptr(out), self._stream()))
        return out

    def zone_fg_counts(self, fgmask: torch.Tensor) -> torch.Tensor:
        """Foreground pixels per zone rectangle, int32 [n, g, g] (= np.sum(fg[a0:a1, b0:b1]) / 255, sf_neural.py:178-180)."""
        if fgmask.dim() == 2:
            fgmask = fgmask.unsqueeze(0)
        n = fgmask.shape[0]
        assert fgmask.is_contiguous() and fgmask.dtype == torch.uint8 and fgmask.shape[1:] == (self.S, self.S)
        out = torch.empty((n, self.gsize, self.gsize), dtype=torch.int32, device=self.device)
        self._check(self.L.ckb_zone_fg_counts(self._h, self._ptr(fgmask), n, self._ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------------------------------------------- K3 + K2
    def find_stones(self, imgs: torch.Tensor, rng_states, rs=0, re=None, cs=0, ce=None, want=("stones", "trusted")):
        """SfClustering.find_stones on n canonical images (uint8 or float32 [n, S, S, 3]).
        rng_states: n cv::RNG states (ints) — see rng_seed / rng_advance. Returns a dict of device tensors."""
        g = self.gsize
        re = g if re is None else re
        ce = g if ce is None else ce
        if imgs.dim() == 3:
            imgs = imgs.unsqueeze(0)
        assert imgs.is_contiguous() and imgs.shape[1:] == (self.S, self.S, 3)
        assert imgs.dtype in (torch.uint8, torch.float32)
        n = imgs.shape[0]
        if isinstance(rng_states, torch.Tensor):
            # states already on the device (int64 bit patterns): nothing is copied. A list goes through a pageable
            # host -> device copy, which makes the driver wait for the stream first; callers that must not stall
            # (DetectPipeline) stage their states in pinned memory and pass the device tensor.
            st = rng_states
            assert st.is_cuda and st.dtype == torch.int64 and st.is_contiguous()
        else:
            st = torch.as_tensor(np.asarray(rng_states, dtype=np.uint64).astype(np.int64), device=self.device)
        assert st.numel() == n
        dev = self.device
        out = {"stones": torch.empty((n, g, g), dtype=torch.uint8, device=dev),
               "trusted": torch.empty((n,), dtype=torch.uint8, device=dev)}
        if "ratios" in want:
            out["ratios"] = torch.empty((n, g, g, 3), dtype=torch.uint8, device=dev)
        if "centers" in want:
            out["centers"] = torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        if "compactness" in want:
            out["compactness"] = torch.empty((n,), dtype=torch.float64, device=dev)
        if "labels" in want:
            x0, y0 = 20 * rs, 20 * cs
            x1 = self.S - 1 if re == g else 20 * re
            y1 = self.S - 1 if ce == g else 20 * ce
            out["labels"] = torch.empty((n, x1 - x0, y1 - y0), dtype=torch.int32, device=dev)
        wb = self.L.ckb_find_stones_workspace(self._h, n)
        work = self._workspace(wb, "kmeans")
        self._check(self.L.ckb_find_stones(self._h, self._ptr(imgs), int(imgs.dtype == torch.float32), n, rs, re, cs,
                                           ce, self._ptr(st), self._ptr(work), work.numel(), self._ptr(out["stones"]),
                                           self._ptr(out["trusted"]), self._ptr(out.get("ratios")),
                                           self._ptr(out.get("centers")), self._ptr(out.get("compactness")),
                                           self._ptr(out.get("labels")), self._stream()))
        return out

    def find_stones_regions(self, imgs: torch.Tensor, regions, rng_states, want=("stones", "trusted")):
        """SfClustering.find_stones(img, rs, re, cs, ce) for every region of `regions` ([(rs, re, cs, ce)], at most 16) on
        each of the n uint8 canonical images, in one set of launches (SfMeta's 3 x 3 Regions, sf_meta.py:245-262).
        rng_states: [n][n_regions] cv::RNG states (ints) or a CUDA int64 tensor of that shape. Returns device tensors
        indexed [image, region]: stones [n, R, g, g] (E outside each region), trusted [n, R], ..."""
        g = self.gsize
        if imgs.dim() == 3:
            imgs = imgs.unsqueeze(0)
        assert imgs.is_contiguous() and imgs.dtype == torch.uint8 and imgs.shape[1:] == (self.S, self.S, 3)
        n, R = imgs.shape[0], len(regions)
        reg = np.ascontiguousarray(np.asarray(regions, dtype=np.int32).reshape(R, 4))
        if isinstance(rng_states, torch.Tensor):
            st = rng_states
            assert st.is_cuda and st.dtype == torch.int64 and st.is_contiguous()
        else:
            st = torch.as_tensor(np.asarray(rng_states, dtype=np.uint64).astype(np.int64), device=self.device)
        assert st.numel() == n * R
        dev = self.device
        out = {"stones": torch.empty((n, R, g, g), dtype=torch.uint8, device=dev),
               "trusted": torch.empty((n, R), dtype=torch.uint8, device=dev)}
        if "ratios" in want:
            out["ratios"] = torch.empty((n, R, g, g, 3), dtype=torch.uint8, device=dev)
        if "centers" in want:
            out["centers"] = torch.empty((n, R, 3, 3), dtype=torch.float32, device=dev)
        if "compactness" in want:
            out["compactness"] = torch.empty((n, R), dtype=torch.float64, device=dev)
        wb = self.L.ckb_find_stones_regions_workspace(self._h, n, R)
        work = self._workspace(wb, "kmeans")
        self._check(self.L.ckb_find_stones_regions(self._h, self._ptr(imgs), n, R, reg.ctypes.data_as(C.c_void_p),
                                                   self._ptr(st), self._ptr(work), work.numel(), self._ptr(out["stones"]),
                                                   self._ptr(out["trusted"]), self._ptr(out.get("ratios")),
                                                   self._ptr(out.get("centers")), self._ptr(out.get("compactness")),
                                                   self._stream()))
        return out

    # ------------------------------------------------------------------------------------ SfMeta / SfContours statistics
    def zone_means(self, imgs: torch.Tensor, masks: torch.Tensor, rs=0, re=None, cs=0, ce=None) -> torch.Tensor:
        """The zone table of SfContours.find_stones (sf_contours.py:87-102,113-126): int16 [n, re-rs, ce-cs, 4] = visible
        flag + mean B, G, R. masks: uint8 [n, S, S], non-zero under the filled convex hulls (built on the host)."""
        g = self.gsize
        re = g if re is None else re
        ce = g if ce is None else ce
        if imgs.dim() == 3:
            imgs, masks = imgs.unsqueeze(0), masks.unsqueeze(0)
        n = imgs.shape[0]
        assert imgs.is_contiguous() and imgs.dtype == torch.uint8 and imgs.shape[1:] == (self.S, self.S, 3)
        assert masks.is_contiguous() and masks.dtype == torch.uint8 and masks.shape == (n, self.S, self.S)
        out = torch.empty((n, re - rs, ce - cs, 4), dtype=torch.int16, device=self.device)
        self._check(self.L.ckb_zone_means(self._h, self._ptr(imgs), self._ptr(masks), n, rs, re, cs, ce, self._ptr(out),
                                          self._stream()))
        return out

    def history_vote(self, history: torch.Tensor, is_empty: torch.Tensor) -> torch.Tensor:
        """Region.commit's vote (sf_meta.py:305-340). history: uint8 [..., histo] colour codes, is_empty: uint8 [...]
        (both on the device). Returns uint8 [...]: 0 = nothing to submit, 1 = B, 2 = W."""
        assert history.is_cuda and history.dtype == torch.uint8 and history.is_contiguous()
        assert is_empty.is_cuda and is_empty.dtype == torch.uint8 and is_empty.is_contiguous()
        assert tuple(history.shape[:-1]) == tuple(is_empty.shape)
        out = torch.empty_like(is_empty)
        self._check(self.L.ckb_history_vote(self._h, self._ptr(history), self._ptr(is_empty), is_empty.numel(),
                                            history.shape[-1], self._ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------------------------------------------------ K4
    def set_cnn_weights(self, params: np.ndarray):
        p = np.ascontiguousarray(params, dtype=np.float32).ravel()
        self._check(self.L.ckb_set_cnn_weights(self._h, p.ctypes.data_as(C.c_void_p), p.size))
        self.has_weights = True

    def cnn_forward(self, goban: torch.Tensor, want_softmax: bool = True, simt: bool = False):
        """NNCache.predict_all_stones + the 0.6 confidence rule on n canonical images (uint8 [n, 380, 380, 3]).
        Returns dict(stones [n,19,19] u8, conf [n,19,19] f32, keep [n,19,19] u8, softmax [n,100,81] f32)."""
        if goban.dim() == 3:
            goban = goban.unsqueeze(0)
        assert goban.is_contiguous() and goban.dtype == torch.uint8 and goban.shape[1:] == (380, 380, 3)
        n = goban.shape[0]
        dev = self.device
        out = {"stones": torch.empty((n, 19, 19), dtype=torch.uint8, device=dev),
               "conf": torch.empty((n, 19, 19), dtype=torch.float32, device=dev),
               "keep": torch.empty((n, 19, 19), dtype=torch.uint8, device=dev)}
        if want_softmax:
            out["softmax"] = torch.empty((n, 100, 81), dtype=torch.float32, device=dev)
        wb = (self.L.ckb_cnn_workspace_simt if simt else self.L.ckb_cnn_workspace)(self._h, n)
        work = self._workspace(wb, "cnn")
        fn = self.L.ckb_cnn_forward_simt if simt else self.L.ckb_cnn_forward
        self._check(fn(self._h, self._ptr(goban), n, self._ptr(work), work.numel(), self._ptr(out.get("softmax")),
                       self._ptr(out["stones"]), self._ptr(out["conf"]), self._ptr(out["keep"]), self._stream()))
        return out

    def cnn_set_debug(self, on: bool):
        """Test aid: keep conv1's activations in the workspace during the following cnn_forward calls."""
        self._check(self.L.ckb_cnn_set_debug(self._h, int(bool(on))))

    def cnn_debug_activation(self, n: int, layer: int) -> torch.Tensor:
        """Test aid: dense float32 copy of an intermediate activation of the last cnn_forward (n <= 64 frames)."""
        shape = {1: (36, 36, 32), 2: (16, 16, 32), 3: (14, 14, 90), 4: (6, 6, 90), 5: (160,)}[layer]
        out = torch.zeros((n * 100,) + shape, dtype=torch.float32, device=self.device)
        self._check(self.L.ckb_cnn_debug_activation(self._h, self._ptr(self._work["cnn"]), n, layer, self._ptr(out),
                                                    self._stream()))
        return out

    # --------------------------------------------------------------------------------- Motion-JPEG ingest (optional)
    def jpeg_backend(self) -> str:
        return self.L.ckb_jpeg_backend(self._h).decode()

    JPEG_LANES = 8

    def jpeg_decode(self, base_address: int, offsets, sizes, out: torch.Tensor, cpu_threads: int = 4, lane: int = 0) -> torch.Tensor:
        """Decode len(offsets) JPEG frames that sit in HOST memory at base_address + offsets[i] (sizes[i] bytes each: e.g. a
        memory-mapped Motion-JPEG file) into `out`, uint8 [>= n, H, W, 3] BGR on the device, on the current stream
        (nvJPEG; see csrc/jpeg_ingest.cu). `lane` < JPEG_LANES picks one of the engine's independent decoders: calls on
        different lanes may run at the same time from different threads (on different streams). Returns out[:n]."""
        n = len(offsets)
        assert out.is_cuda and out.dtype == torch.uint8 and out.dim() == 4 and out.shape[3] == 3 and out.shape[0] >= n
        assert out.stride(3) == 1 and out.stride(2) == 3
        ptrs = (C.c_void_p * n)(*[base_address + int(o) for o in offsets])
        lens = (C.c_size_t * n)(*[int(v) for v in sizes])
        self._check(self.L.ckb_jpeg_decode(self._h, ptrs, lens, n, out.shape[1], out.shape[2], self._ptr(out), out.stride(1),
                                           out.stride(0), cpu_threads, lane, self._stream()))
        return out[:n]

    # ------------------------------------------------------------------------------------------------ host staging
    def frame_roi(self, mtx, H: int, W: int):
        """(y0, y1, x0, x1): the part of an H x W frame the warp can read under homography mtx."""
        m = np.ascontiguousarray(np.asarray(mtx, dtype=np.float64)).reshape(9)
        roi = np.zeros(4, np.int32)
        rc = self.L.ckb_frame_roi(m.ctypes.data_as(C.c_void_p), H, W, self.S, roi.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise CkbError(rc, "ckb_frame_roi: bad argument")
        return tuple(int(v) for v in roi)

    def upload_frames(self, host: torch.Tensor, dev: torch.Tensor, roi=None):
        """Asynchronous H2D copy (current stream) of `roi` of each HOST frame [n, H, W, 3] into dev [>= n, H, W, 3]."""
        assert not host.is_cuda and dev.is_cuda and host.dtype == torch.uint8 and dev.dtype == torch.uint8
        assert host.dim() == 4 and host.stride(3) == 1 and host.stride(2) == 3 and dev.stride(2) == 3
        n, H, W, _ = host.shape
        assert dev.shape[0] >= n and tuple(dev.shape[1:]) == (H, W, 3)
        r = None
        if roi is not None:
            r = np.asarray(roi, dtype=np.int32)
        self._check(self.L.ckb_upload_frames(self._h, C.c_void_p(host.data_ptr()), n, H, W, host.stride(1),
                                             host.stride(0), r.ctypes.data_as(C.c_void_p) if r is not None else None,
                                             self._ptr(dev), dev.stride(1), dev.stride(0), self._stream()))
        bytes_per_frame = (H * W * 3) if roi is None else (roi[1] - roi[0]) * (roi[3] - roi[2]) * 3
        return n * bytes_per_frame

    # ------------------------------------------------------------------------------------------- per-kernel timing
    def profile_begin(self, capacity: int = 8192):
        self._check(self.L.ckb_profile_begin(self._h, capacity))
        self._prof_cap = capacity

    def profile_end(self):
        """[(kernel name, ms)] for every launch since profile_begin, in issue order."""
        cap = getattr(self, "_prof_cap", 8192)
        names = C.create_string_buffer(cap * 32)
        ms = np.zeros(cap, np.float32)
        n = C.c_int(0)
        self._check(self.L.ckb_profile_end(self._h, cap, names, ms.ctypes.data_as(C.c_void_p), C.byref(n)))
        raw = names.raw
        return [(raw[i * 32:(i + 1) * 32].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n.value)]
