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


def pcie_levels(n_gpus: int = None):
    """Pairwise PCIe distance of the visible GPUs: NVML's common-ancestor level (10 = one bridge, 20 = several bridges
    of one switch, 30 = host bridge, 40 = same NUMA node, 50 = across sockets), NVLink ignored. None without NVML."""
    try:
        import pynvml
        pynvml.nvmlInit()
        n = pynvml.nvmlDeviceGetCount() if n_gpus is None else n_gpus
        hs = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(n)]
        return [[0 if i == j else int(pynvml.nvmlDeviceGetTopologyCommonAncestor(hs[i], hs[j])) for j in range(n)]
                for i in range(n)]
    except Exception:
        return None


def pick_gpus(world: int, levels=None):
    """GPU index for each local rank of a `world`-process run on one box: when fewer processes than GPUs run, spread them
    over distinct PCIe uplinks (GPUs behind one switch share its host link, and the end-to-end path is bound by host ->
    device copies). Greedy farthest-first on the NVML common-ancestor level; the identity map when NVML gives nothing or
    every GPU is in use."""
    levels = pcie_levels() if levels is None else levels
    if not levels or world >= len(levels):
        return list(range(world))
    n = len(levels)
    chosen = [0]
    while len(chosen) < world:
        best, best_key = None, None
        for g in range(n):
            if g in chosen:
                continue
            d = [levels[g][c] for c in chosen]
            key = (min(d), sum(d), -g)          # farthest from its nearest chosen neighbour, then overall, then low index
            if best_key is None or key > best_key:
                best, best_key = g, key
        chosen.append(best)
    return chosen
