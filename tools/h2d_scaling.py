"""Aggregate pinned-host -> device copy bandwidth with one process per GPU (torchrun): the ceiling of the end-to-end path."""
import os
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
d = torch.empty_like(h, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([20 * h.numel() / dt / 1e9], device="cuda")
if world > 1:
    dist.all_reduce(gbs)
    dist.destroy_process_group()
if rank == 0:
    print("ranks %d: aggregate H2D %.1f GB/s (%.1f per GPU)" % (world, float(gbs), float(gbs) / world))
