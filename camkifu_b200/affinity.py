"""Host-side placement for one-process-per-GPU runs: bind the process to the CPUs (and so, by first touch, the pinned
staging buffers to the memory) of the NUMA node its GPU hangs off. With 8 ranks streaming frames over PCIe at once, host
memory that sits on the other socket costs a large part of the aggregate host-to-device bandwidth. Plumbing only."""
import os


def bind_to_gpu(index: int, min_cpus: int = 4) -> bool:
    """Pin the calling process to the CPU set NVML reports as local to GPU `index` (nvmlDeviceSetCpuAffinity), when that
    helps: only if the set NVML would install is a proper subset of the CPUs the process may use now (i.e. the box really
    has several NUMA nodes) and still holds at least `min_cpus` CPUs — the decoder threads of the offline-video path
    live in this process. Returns False (and changes nothing) otherwise, or when NVML / the affinity call is
    unavailable, e.g. inside a restricted cgroup."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(before) // 64) + 1)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        target = ideal & before
        if len(target) < min_cpus or target == before:
            return False
        os.sched_setaffinity(0, target)
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


def measure_gpu_choice(world: int, n_gpus: int = None, mbytes: int = 64, reps: int = 4):
    """Which `world` GPUs of the box to use, MEASURED. NVML reports every GPU pair of these boxes at the same PCIe level and
    any TWO GPUs copy from the host at twice the rate of one (110 GB/s), yet four ranks on GPUs 0-3 together get 115 GB/s
    while GPUs 0, 1, 4, 5 get 220: the GPUs hang off host bridges with a shared ceiling that only shows with three or
    more streams. So the choice is greedy on the measured aggregate: start with GPU 0; add, one at a time, the GPU with
    which all chosen GPUs together copy pinned host buffers fastest (ties within 5 % go to the lowest index). Returns
    (gpus, log) with log = [(candidate set, GB/s)]. Creates a CUDA context on every GPU: call from ONE process."""
    import time
    import torch
    n = torch.cuda.device_count() if n_gpus is None else n_gpus
    nbytes = mbytes << 20
    host = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(world)]
    dev = [torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % g) for g in range(n)]
    streams = [torch.cuda.Stream(device="cuda:%d" % g) for g in range(n)]

    def run(gpus):
        dt = 1.0
        for _ in range(2):                                   # warm-up, then timed
            for g in gpus:
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for _ in range(reps):
                for k, g in enumerate(gpus):
                    with torch.cuda.stream(streams[g]):
                        dev[g].copy_(host[k], non_blocking=True)
            for g in gpus:
                streams[g].synchronize()
            dt = time.perf_counter() - t0
        return len(gpus) * reps * nbytes / dt / 1e9

    chosen, log = [0], []
    while len(chosen) < world:
        best, best_bw = None, 0.0
        for g in range(n):
            if g in chosen:
                continue
            bw = run(chosen + [g])
            log.append((chosen + [g], round(bw, 1)))
            if bw > 1.05 * best_bw:
                best, best_bw = g, bw
        chosen.append(best)
    return chosen, log


def gpu_map_for_run(world: int, local_rank: int, token: str, timeout_s: float = 90.0):
    """The LOCAL_RANK -> GPU map of a `world`-process run on one box, agreed through a small file: local rank 0 measures
    which GPUs to use (only when fewer processes than GPUs run — otherwise the identity is the only map) and publishes the
    map; the other ranks wait for it. Falls back to the identity map on any failure."""
    import json
    import time
    import torch
    ident = list(range(world))
    n = torch.cuda.device_count()
    if world <= 1 or world >= n:
        return ident, None
    path = os.path.join("/tmp", "ckb_gpu_map_%s_w%d.json" % (token, world))
    t_start = time.time()          # a file left by an earlier run with the same token is older than this run: ignored
    if local_rank == 0:
        info = {"map": ident}
        try:
            gpus, log = measure_gpu_choice(world, n)
            info = {"map": gpus, "measured_h2d_gbs": [["+".join(str(g) for g in c), bw] for c, bw in log]}
        except Exception as e:                                  # never let the probe take the run down
            info["error"] = repr(e)
        with open(path + ".tmp", "w") as f:
            json.dump(info, f)
        os.replace(path + ".tmp", path)
        return info["map"], info
    t0 = time.time()
    while time.time() - t0 < timeout_s:
        if os.path.exists(path) and os.path.getmtime(path) > t_start - 3.0:
            try:
                with open(path) as f:
                    info = json.load(f)
                return info["map"], info
            except Exception:
                pass
        time.sleep(0.05)
    return ident, None
