"""Host-side placement for one-process-per-GPU runs: bind the process to the CPUs (and so, by first touch, the pinned
staging buffers to the memory) of the NUMA node its GPU hangs off. With 8 ranks streaming frames over PCIe at once, host
memory that sits on the other socket costs a large part of the aggregate host-to-device bandwidth. Plumbing only."""
import os


def bind_to_gpu(index: int) -> bool:
    """Pin the calling process to the CPU set NVML reports as local to GPU `index` (nvmlDeviceSetCpuAffinity).
    Returns False (and changes nothing) when NVML or the affinity call is unavailable, e.g. inside a restricted cgroup."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = os.sched_getaffinity(0)
        if not after:                      # never leave the process without CPUs
            os.sched_setaffinity(0, before)
            return False
        return True
    except Exception:
        return False
