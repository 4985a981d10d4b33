#!/usr/bin/env python
"""bench.py — stone-detect frames/s @1080p 19x19 (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] — SfNeural CNN stone classification on synthetic 1080p 19x19 frames,
64 frames per step on each GPU: ckb_warp (cv2.warpPerspective, stonesfinder.py:140) -> ckb_cnn_forward
(NNCache.predict_all_stones + the 0.6 confidence rule, nn_cache.py:25-52, sf_neural.py:57-70). Random-init (Glorot) weights
of the reference architecture: the trained weights do not ship with the reference.

`value`   frames/s with the frames resident in HBM, CUDA events on the launching stream, max over ranks.
`e2e`     the same metric through camkifu_b200.pipeline.DetectPipeline.detect_stream() with HOST (pinned) frames: H2D of
          the frames and D2H of the board states inside the timed region (batch k+1 uploads while batch k computes).
`roofline` dominant kernel = cnn_tc_front (patch gather + conv1 + conv2 + pool on tcgen05); achieved = algorithmic FLOP / mean launch time measured with
          CUDA events in the timed region (ckb_profile_begin/end); peak from MEASURED_PEAKS.json.
`cpu_baseline` the reference's CPU path (cv2 warp + fp32 CNN, oracle/) on a bounded sample, timed on this host.
`--impl reference` times that CPU path alone with all host threads and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "stone-detect frames/s @1080p 19x19"
UNIT = "frames/s"
H, W, GSIZE, BATCH = 1080, 1920, 19, 64
WORKLOAD = "SfNeural CNN stone classification, synthetic 1080p 19x19 frames, 64 frames per step per GPU (warp + CNN + decode)"
CNN_MAC_PER_PATCH = {"conv1": 36 * 36 * 75 * 32, "conv2": 32 * 32 * 800 * 32, "conv3": 14 * 14 * 288 * 90,
                     "conv4": 12 * 12 * 810 * 90, "fc1": 3240 * 160, "fc2": 160 * 81}
assert sum(CNN_MAC_PER_PATCH.values()) == 45434080
# dram__bytes_read.sum + dram__bytes_write.sum of one cnn_tc_front launch (64 frames), from the ncu --set full capture
# committed under profiles/ (None until a capture exists for the current kernel)
FRONT_DRAM_TRAFFIC_BYTES = 28010400 + 151286000   # profiles/r1q_kernels_ncu_full_selected.csv (dram__bytes_read.sum + dram__bytes_write.sum)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "MEASURED_PEAKS.json (sustained bf16)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)            # first queries are slow: pay for
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)             # them outside the timed region
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.004)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_fps(frames, mtx, params, threads: int, budget_s: float = 15.0, max_frames: int = 64):
    """The reference's per-frame CPU path on this host: cv2.warpPerspective (stonesfinder.py:140) + the SfNeural net on
    the 100 patches + decode (nn_cache.py:25-52). Keras/Theano are not installed anywhere here, so `net.predict` is the
    oracle's fp32 torch-CPU stand-in, fed one 100-patch batch per frame (kinder than the reference's 100 batch-1 calls).
    Returns (frames/s, frames timed)."""
    import cv2
    import torch
    from oracle import oracle as O
    cv2.setNumThreads(threads)
    torch.set_num_threads(threads)
    predict = O.torch_cnn(params)
    done, t0 = 0, time.perf_counter()
    # one untimed frame (thread pools, allocator)
    g = cv2.warpPerspective(frames[0], mtx, (380, 380))
    predict(O.c_nn_gather(g))
    t0 = time.perf_counter()
    while done < max_frames:
        f = frames[done % len(frames)]
        g = cv2.warpPerspective(f, mtx, (380, 380))
        y = predict(O.c_nn_gather(g))
        O.c_nn_decode(y)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done


def run_reference(args, rank, world):
    if rank != 0:
        return
    from camkifu_b200 import synth, weights
    threads = os.cpu_count() or 1
    frames, mtx, truth, _ = synth.make_clip_parallel(1000, 8, H, W)
    params = weights.glorot_params(seed=0)
    per_step = max(4, min(16, BATCH))
    times = []
    for s in range(args.warmup + args.steps):
        fps, n = cpu_reference_fps(frames, mtx, params, threads, budget_s=8.0, max_frames=per_step)
        if s >= args.warmup:
            times.append(n / fps)
    ms = 1e3 * sum(times) / len(times)
    value = per_step / (ms / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "frame": [H, W], "gsize": GSIZE, "frames_per_step": per_step,
                       "note": "bounded sample of the workload: %d frames per step on the host CPU" % per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d frames/step x %d steps: cv2.warpPerspective + fp32 CNN (torch-CPU stand-in "
                                       "for Keras predict, one 100-patch batch per frame) + decode" % (per_step, args.steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------- B200 path
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from camkifu_b200 import synth, weights
    from camkifu_b200.engine import StoneEngine
    from camkifu_b200.pipeline import DetectPipeline, pinned_frames

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from camkifu_b200.affinity import bind_to_gpu
    numa_bound = bind_to_gpu(local_rank) if world > 1 else False   # pinned staging buffers on the GPU's own NUMA node
    eng = StoneEngine(GSIZE, device=dev)
    params = weights.glorot_params(seed=0)
    eng.set_cnn_weights(params)

    # synthetic 1080p clip: 64 distinct frames (one homography), pinned on the host and resident in HBM
    frames_np, mtx, truth, _ = synth.make_clip_parallel(1000 + rank, BATCH, H, W)
    host = pinned_frames(BATCH, H, W)
    host.copy_(torch.from_numpy(frames_np))
    n_rot = 2   # two resident batches (796 MB > the 126 MB L2): consecutive steps never read the same frames
    resident = [host.to(dev, non_blocking=True)]
    resident.append(torch.roll(resident[0], shifts=7, dims=0).contiguous())
    goban = torch.empty((BATCH, 380, 380, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step(i):
        eng.warp(resident[i % n_rot], mtx, out=goban)
        return eng.cnn_forward(goban, want_softmax=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # parity spot check outside the timed region (rank 0): warp bit-exact, board state of frame 0 against the oracle
    check = None
    if rank == 0:
        from oracle import oracle as O
        out = step(0)
        torch.cuda.synchronize()
        g0 = goban[0].cpu().numpy()
        warp_ok = bool(np.array_equal(g0, O.c_warp(frames_np[0], mtx, 380)))
        y = O.c_cnn_forward(O.c_nn_gather(g0), params)
        s_ref, c_ref, k_ref = O.c_nn_decode(y)
        check = {"warp_bit_exact": warp_ok, "stones_equal": bool(np.array_equal(out["stones"][0].cpu().numpy(), s_ref)),
                 "conf_max_abs_err": float(np.abs(out["conf"][0].cpu().numpy() - c_ref).max())}

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launches
    eng.profile_begin(capacity=max(64, 16 * args.steps + 16))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(args.warmup + i)
    stones_all = None
    if world > 1:   # the one collective of the path: final gather of the per-frame board states
        mine = out["stones"].reshape(BATCH, 361)
        stones_all = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(stones_all, mine)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = eng.profile_end()
    launches = eng.launches - launches0
    clocks = sampler.finish()

    # ---- end to end: host frames through the public batch API
    pipe = DetectPipeline(H, W, GSIZE, mode="neural", sub_batch=16, engine=eng)
    for _ in range(max(1, min(args.warmup, 3))):
        pipe.detect(host, mtx)
    barrier()
    e2e_steps = args.steps
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    # offline-video form of the API: batch k+1 uploads while batch k computes; every result is read back on the host
    for res in pipe.detect_stream(((host, mtx) for _ in range(e2e_steps)), depth=2):
        pass
    f1.record()
    barrier()
    e2e_s = f0.elapsed_time(f1) / 1e3
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_ok = bool(np.array_equal(res["stones"], out["stones"].cpu().numpy())) if (args.steps + args.warmup - 1) % n_rot == 0 else None

    # ---- max over ranks
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        return
    ms_step = ms_total / args.steps
    value = world * BATCH * args.steps / (ms_total / 1e3)
    e2e_value = world * BATCH * e2e_steps / e2e_s

    # ---- per-kernel shares and the roofline of the dominant kernel
    agg = {}
    for name, ms in prof:
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
    kern = sorted(((n, v[0] / v[1], v[1]) for n, v in agg.items()), key=lambda x: -x[1] * x[2])
    ksum = sum(v[0] for v in agg.values())
    peaks = measured_peaks()
    front = "cnn_tc_front"          # gather + conv1 + conv2 + pool: the dominant kernel
    front_ms = agg[front][0] / agg[front][1]
    front_flop = 2.0 * (CNN_MAC_PER_PATCH["conv1"] + CNN_MAC_PER_PATCH["conv2"]) * 100 * BATCH
    achieved = front_flop / (front_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": front, "achieved": achieved,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                "traffic": FRONT_DRAM_TRAFFIC_BYTES, "peak_source": peaks["source"],
                "algorithmic_flop_per_launch": front_flop, "ms_per_launch": front_ms,
                "share_of_step": agg[front][0] / ksum,
                "note": "algorithmic FLOP = 2 x (3 110 400 conv1 + 26 214 400 conv2) MAC x 6400 patches; the kernel issues 3 "
                        "bf16 products per conv2 MAC and 2 per conv1 MAC (hi/lo operand split for the 1e-3 softmax bar), "
                        "so ~1/3 is its ceiling in these units"}
    cnn_flop = 2.0 * sum(CNN_MAC_PER_PATCH.values()) * 100 * BATCH
    cnn_ms = sum(v[0] for n, v in agg.items() if n.startswith("cnn_")) / args.steps

    # ---- CPU baseline on this host (bounded sample; rank 0 of the N = 1 run only)
    threads = os.cpu_count() or 1
    if world == 1:
        cpu_fps, cpu_n = cpu_reference_fps(frames_np, mtx, params, threads, budget_s=15.0, max_frames=64)
        cpu_line = {"value": cpu_fps, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "%d frames of the same clip: cv2.warpPerspective + fp32 CNN (torch-CPU stand-in "
                              "for Keras predict, one 100-patch batch per frame) + decode" % cpu_n}
    else:
        cpu_line = {"value": None, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "not timed at N > 1: see the N = 1 line and --impl reference"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [H, W], "gsize": GSIZE, "frames_per_step_per_gpu": BATCH,
                       "weights": "glorot_uniform seed 0 (reference architecture, nn_manager.py:277-298)",
                       "l2": "inputs larger than L2: two resident 398 MB batches used alternately",
                       "parallelism": "frames sharded across %d GPU(s), final all_gather of board states" % world,
                       "numa_bound": numa_bound},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "camkifu_b200.pipeline.DetectPipeline.detect_stream (pinned host frames, ROI upload, 16-frame "
                           "sub-batches double buffered, results of every batch read back to the host)", "matches_resident_path": e2e_ok},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_line,
            "kernels": [{"name": n, "ms": round(ms, 4), "launches_per_step": c / args.steps,
                         "share": round(ms * c / ksum, 4)} for n, ms, c in kern],
            "cnn": {"tflops_algorithmic": cnn_flop / (cnn_ms * 1e-3) / 1e12, "ms_per_step": cnn_ms},
            "parity_check": check}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
