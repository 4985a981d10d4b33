"""Host side of the offline video path (camkifu_b200/video.py): decode thread, pinned ring, frame ranges."""
import os

import numpy as np
import pytest

from camkifu_b200 import synth


def write_lossless(path, frames):
    import cv2
    h, w = frames.shape[1:3]
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30, (w, h))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()


@pytest.fixture(scope="module")
def clip(tmp_path_factory):
    frames, mtx, truth, _ = synth.make_clip(21, 23, 120, 160)
    path = str(tmp_path_factory.mktemp("video") / "clip.avi")
    write_lossless(path, frames)
    return path, frames, mtx


@pytest.mark.parametrize("rng", [(0, None), (6, 19), (9, 9), (20, 400)])
def test_frame_source_reads_ranges_bit_exact(clip, rng):
    from camkifu_b200.video import FrameSource
    path, frames, _ = clip
    start, stop = rng
    src = FrameSource(path, start, stop, batch=4, depth=3)
    assert (src.n_total, src.H, src.W) == (23, 120, 160)
    got, pos = [], start
    for buf, m, first in src:
        assert first == pos and 1 <= m <= 4
        got.append(buf[:m].numpy().copy())
        pos += m
        src.release(buf)                     # without this the decoder would stall after `depth` batches
    want = frames[start:min(stop if stop is not None else 23, 23)]
    assert len(src) == len(want)
    if len(want):
        assert np.array_equal(np.concatenate(got), want)
    else:
        assert got == []


def test_frame_source_array_and_errors(clip):
    from camkifu_b200.video import FrameSource, open_source
    _, frames, _ = clip
    src = FrameSource(frames, 3, 11, batch=5, depth=2)
    out = []
    for buf, m, first in src:
        out.append(buf[:m].numpy().copy())
        src.release(buf)
    assert np.array_equal(np.concatenate(out), frames[3:11])
    with pytest.raises(IOError):
        open_source(os.path.join(os.path.dirname(__file__), "no_such_video.avi"))


def test_frame_source_several_decoders_cover_the_range(clip):
    """decoders = 3: three captures over contiguous parts of the range; batches arrive interleaved, each tagged with its
    first frame index, every frame exactly once and bit-exact."""
    from camkifu_b200.video import FrameSource
    path, frames, _ = clip
    src = FrameSource(path, 2, 23, batch=4, depth=2, decoders=3)
    seen = np.zeros(23, bool)
    for buf, m, first in src:
        assert np.array_equal(buf[:m].numpy(), frames[first:first + m])
        assert not seen[first:first + m].any()
        seen[first:first + m] = True
        src.release(buf)
    assert seen[2:].all() and not seen[:2].any() and src.frames_read == 21


def test_ring_clip_is_zero_copy():
    import torch
    from camkifu_b200.video import FrameSource, RingClip
    ring = torch.arange(5 * 4 * 6 * 3, dtype=torch.uint8).reshape(5, 4, 6, 3)
    src = FrameSource(RingClip(ring, 23), 3, 23, batch=4)
    pos = 3
    for buf, m, first in src:
        assert first == pos and buf.data_ptr() == ring[first % 5].data_ptr()      # a view of the ring, not a copy
        assert torch.equal(buf[:m], ring[first % 5:first % 5 + m])
        pos += m
        src.release(buf)
    assert pos == 23


def test_mjpeg_avi_index(tmp_path):
    """MjpegAvi (the index the device-side JPEG ingest reads): one entry per frame in file order, each a complete JPEG
    (SOI ... EOI) of the size the container declares; repeat() tiles the index; other containers are refused."""
    import cv2
    from camkifu_b200.video import MjpegAvi
    n, H, W = 7, 48, 64
    rng = np.random.default_rng(0)
    path = str(tmp_path / "m.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30, (W, H))
    ramp = np.add.outer(np.arange(H), np.arange(W))[:, :, None] % 64
    frames = [(ramp + rng.integers(0, 192, 3)[None, None, :]).astype(np.uint8) for _ in range(n)]   # smooth, distinct
    for f in frames:
        wr.write(f)
    wr.release()
    avi = MjpegAvi(path)
    assert (len(avi), avi.H, avi.W) == (n, H, W)
    assert np.all(np.diff(avi.offsets) > 0) and np.all(avi.sizes > 0)
    raw = np.fromfile(path, dtype=np.uint8)
    for i in range(n):
        blob = raw[avi.offsets[i]: avi.offsets[i] + avi.sizes[i]]
        assert blob[0] == 0xFF and blob[1] == 0xD8
        img = cv2.imdecode(blob, cv2.IMREAD_COLOR)
        assert img.shape == (H, W, 3)
    cap = cv2.VideoCapture(path)
    ok, first = cap.read()
    cap.release()
    # the indexed chunk is the frame the container's decoder returns (two JPEG decoders: chroma upsampling differs)
    mine = cv2.imdecode(raw[avi.offsets[0]: avi.offsets[0] + avi.sizes[0]], cv2.IMREAD_COLOR)
    assert ok and np.abs(first.astype(np.int16) - mine.astype(np.int16)).mean() < 3
    assert np.abs(first.astype(np.int16) - frames[1].astype(np.int16)).mean() > 10
    rep = avi.repeat(3)
    assert len(rep) == 3 * n and np.array_equal(rep.offsets[n:2 * n], avi.offsets) and (rep.H, rep.W) == (H, W)
    other = tmp_path / "not_avi.bin"
    other.write_bytes(b"\x00" * 64)
    with pytest.raises(ValueError):
        MjpegAvi(str(other))


def test_nvjpeg_ingest_argument_checks(tmp_path):
    """ingest="nvjpeg" is for Motion-JPEG AVI files and stateless modes: anything else is refused before any device work."""
    from camkifu_b200.video import process_video
    frames = np.zeros((4, 24, 32, 3), np.uint8)
    with pytest.raises(ValueError):
        process_video(frames, np.eye(3), mode="neural", ingest="nvjpeg")
    other = tmp_path / "clip.bin"
    other.write_bytes(b"RIFF" + b"\x00" * 60)
    with pytest.raises(ValueError):
        process_video(str(other), np.eye(3), mode="neural", ingest="nvjpeg")
    with pytest.raises(ValueError):
        process_video(str(other), np.eye(3), mode="full", ingest="nvjpeg")
