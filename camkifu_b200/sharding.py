"""Frame-range sharding of an offline video across the GPUs of one box (one process per GPU).

The path has no exchange step while it computes: frames are independent units for the stateless calls (warp,
find_stones on a given image, predict_all), so each rank takes a contiguous frame range and the only collective is the
final gather of per-frame board states (uint8 [n, 361] — 36 MB per 100 k frames). Streaming state is handled per shard:
  * the every-third-frame cadence of SfClustering._find (sf_clustering.py:37): shard starts are multiples of 3;
  * cv2's process-global RNG carried across k-means calls: `rng_state_at` replays the draw count (39 per call);
  * the running average (alpha = 0.2): `halo` frames before the shard start warm it up. The restarted float32 recurrence
    converges to the streaming one geometrically (255 x 0.8^120 = 6e-10, below half an ulp of every accumulator value
    above 0.01), after which the two are equal bit for bit in practice (tests/test_sharding.py compares them); it is a
    convergence argument, not an identity: callers that need a guarantee hand the accumulator itself across the seam.
"""
import numpy as np

from .engine import DRAWS_PER_KMEANS

ACCU_HALO = 120


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


def gather_ragged(local, group=None):
    """All ranks -> the concatenation, in rank order, of every rank's [k_r, ...] tensor (k_r may differ and be zero), on
    every rank: the ranks first agree on the counts, so a rank that got fewer frames than planned (a video file shorter
    than its header says) cannot stall the others."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c) for c in counts]
    longest = max(max(counts), 1)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:counts[r]] for r in range(world)], dim=0)
