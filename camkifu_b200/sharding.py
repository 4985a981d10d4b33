"""Frame-range sharding of an offline video across the GPUs of one box (one process per GPU).

The path has no exchange step while it computes: frames are independent units for the stateless calls (warp,
find_stones on a given image, predict_all), so each rank takes a contiguous frame range and the only collective is the
final gather of per-frame board states (uint8 [n, 361] — 36 MB per 100 k frames). Streaming state is handled per shard:
  * the every-third-frame cadence of SfClustering._find (sf_clustering.py:37): shard starts are multiples of 3;
  * cv2's process-global RNG carried across k-means calls: `rng_state_at` replays the draw count (39 per call);
  * the running average (alpha = 0.2): `halo` frames before the shard start warm it up (0.8^80 < 2^-24).
"""
import numpy as np

from .engine import DRAWS_PER_KMEANS

ACCU_HALO = 80


def shard_range(n_frames: int, rank: int, world: int, align: int = 3):
    """[start, stop) of `rank`: near-equal contiguous ranges whose starts are multiples of `align`."""
    assert 0 <= rank < world and n_frames >= 0 and align >= 1
    units = (n_frames + align - 1) // align
    lo = (units * rank) // world
    hi = (units * (rank + 1)) // world
    return min(lo * align, n_frames), min(hi * align, n_frames)


def halo_start(start: int, halo: int = ACCU_HALO, align: int = 3) -> int:
    """First frame a shard must feed its running average to reproduce the streaming state at `start`."""
    return max(0, (start - halo) // align * align)


def kmeans_calls_before(frame: int, every: int = 3) -> int:
    """Number of k-means calls the streaming finder has made before processing `frame` (it detects on frames 0, 3, …)."""
    return (frame + every - 1) // every


def rng_state_at(state0: int, frame: int, every: int = 3) -> int:
    """cv::RNG state at the k-means call of `frame`, given the state before frame 0."""
    from .engine import rng_advance
    return rng_advance(state0, kmeans_calls_before(frame, every))


def gather_board_states(local, n_frames: int, align: int = 3, group=None):
    """All ranks -> the full [n_frames, ...] tensor of board states, in frame order, on every rank.
    `local` is this rank's torch tensor [stop - start, ...] (CPU with gloo, CUDA with NCCL)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = [shard_range(n_frames, r, world, align) for r in range(world)]
    longest = max(b - a for a, b in ranges)
    assert local.shape[0] == ranges[rank][1] - ranges[rank][0]
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:b - a] for r, (a, b) in enumerate(ranges)], dim=0)
