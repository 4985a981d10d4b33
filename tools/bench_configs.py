#!/usr/bin/env python
"""Measurements of the BASELINE.json configs that bench.py does not time (bench.py = configs[1], the headline line).

    python tools/bench_configs.py [--frames N] > gpurun_out/configs.jsonl

One JSON line per config, inputs resident in HBM, CUDA events on the launching stream, per-kernel times from the library's
own event log (ckb_profile_begin/end), HBM roofline fractions for the byte-bound kernels against MEASURED_PEAKS.json:

  config 1  SfClustering on 640x480 frames: (i) the reference's streaming semantics (running average every frame, k-means
            of columns 6..12 every 3rd frame, sf_clustering.py:23-46), (ii) full-board find_stones on every frame
  config 3  the whole pipeline on 1080p frames: warp + MOG2 + running average + full-board k-means + CNN predict_all
  config 4  3840x2160 frames, boards of 9 / 13 / 19 lines: warp + k-means (bit-exactness of this config against the
            oracle is a test: tests/test_gpu_scale.py::test_config4_4k_mixed_board_sizes)
Algorithmic bytes per frame are SURVEY.md section 8(d)'s figures (stated in DESIGN.md section 4).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from camkifu_b200 import synth, weights  # noqa: E402
from camkifu_b200.engine import StoneEngine, rng_seed, rng_advance  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_table(eng, fn):
    eng.profile_begin(capacity=8192)
    fn()
    torch.cuda.synchronize()
    agg = {}
    for name, ms in eng.profile_end():
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
    return agg


def quad_area(mtx, S):
    inv = np.linalg.inv(mtx)
    pts = []
    for x, y in ((0, 0), (S, 0), (S, S), (0, S)):
        v = inv @ np.array([x, y, 1.0])
        pts.append((v[0] / v[2], v[1] / v[2]))
    a = 0.0
    for i in range(4):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % 4]
        a += x0 * y1 - x1 * y0
    return abs(a) / 2


def warp_bytes(mtx, S):
    return 3 * S * S + 3 * min(4 * S * S, quad_area(mtx, S))


def clip(seed, n, H, W, gsize=19, distinct=16):
    frames, mtx, truth, _ = synth.make_clip_parallel(seed, distinct, H, W, gsize=gsize)
    reps = (n + distinct - 1) // distinct
    return np.concatenate([frames] * reps)[:n], mtx, np.concatenate([truth] * reps)[:n]


def config1(n):
    H, W, S = 480, 640, 380
    eng = StoneEngine(19)
    frames, mtx, truth = clip(1, n, H, W)
    d = torch.from_numpy(frames).cuda()
    goban = torch.empty((n, S, S, 3), dtype=torch.uint8, device="cuda")
    accu = torch.empty((S, S, 3), dtype=torch.float32, device="cuda")
    st0 = rng_seed(0)
    k = (n + 2) // 3
    states_stream = [rng_advance(st0, i) for i in range(k)]
    states_full = [rng_advance(st0, i) for i in range(n)]
    res = {}

    def stream():
        eng.warp(d, mtx, out=goban)
        snaps = eng.accumulate(goban, accu, first=True, snap_every=3, snap_phase=0)
        res["s"] = eng.find_stones(snaps, states_stream, 0, 19, 6, 13)

    def full():
        eng.warp(d, mtx, out=goban)
        res["f"] = eng.find_stones(goban, states_full)

    ms_s, ms_f = timed(stream, 3), timed(full, 3)
    agg = kernel_table(eng, full)
    hbm, src = peaks()
    wb = warp_bytes(mtx, S)
    kb = 3 * 143641 + 4 * 143641 + 1083
    warp_ms = agg["ckb_warp_kernel"][0]
    km_ms = sum(v[0] for name, v in agg.items() if name != "ckb_warp_kernel")
    acc = float((res["f"]["stones"].cpu().numpy() == truth).mean())
    return {"config": "1: SfClustering k-means stone detection, synthetic 640x480 19x19, homography given", "frames": n,
            "stream_semantics": {"frames_per_s": n / (ms_s / 1e3), "note": "warp + running average every frame, k-means of "
                                 "columns 6..12 on every 3rd frame (sf_clustering.py:23-46)"},
            "full_board_every_frame": {"frames_per_s": n / (ms_f / 1e3), "label_accuracy_vs_truth": acc},
            "roofline": {"warp": {"bound": "hbm", "bytes_per_frame": wb, "achieved_gbs": wb * n / (warp_ms / 1e3) / 1e9,
                                  "frac": wb * n / (warp_ms / 1e3) / 1e9 / hbm},
                         "kmeans+zones": {"bound": "hbm (compulsory bytes; the stage is latency/sync bound)", "bytes_per_frame": kb,
                                          "achieved_gbs": kb * n / (km_ms / 1e3) / 1e9, "frac": kb * n / (km_ms / 1e3) / 1e9 / hbm},
                         "peak_gbs": hbm, "peak_source": src},
            "kernels_ms": {k_: round(v[0], 4) for k_, v in agg.items()}}


def config3(n):
    H, W, S, B = 1080, 1920, 380, 64
    eng = StoneEngine(19)
    eng.set_cnn_weights(weights.glorot_params(seed=0))
    frames, mtx, truth = clip(3, B, H, W)
    d = torch.from_numpy(frames).cuda()
    goban = torch.empty((B, S, S, 3), dtype=torch.uint8, device="cuda")
    accu = torch.empty((S, S, 3), dtype=torch.float32, device="cuda")
    fg = torch.empty((B, S, S), dtype=torch.uint8, device="cuda")
    bg = eng.mog2_new_state()
    st0 = rng_seed(0)
    state = {"frames": 0}
    out = {}

    def batch():
        f0 = state["frames"]
        eng.warp(d, mtx, out=goban)
        eng.mog2_apply(goban, bg, f0, [0.01 if f0 + i < 50 else 0.005 for i in range(B)], out=fg)
        eng.zone_fg_counts(fg)
        eng.accumulate(goban, accu, first=(f0 == 0))
        out["km"] = eng.find_stones(goban, [rng_advance(st0, f0 + i) for i in range(B)])
        out["nn"] = eng.cnn_forward(goban, want_softmax=False)
        state["frames"] = f0 + B

    steps = max(1, n // B)
    ms = timed(batch, steps)
    agg = kernel_table(eng, batch)
    hbm, src = peaks()
    mog_ms = agg["ckb_mog2_kernel"][0]
    mog_bytes = 2 * (25 * 4 + 1) * S * S + B * (3 + 1) * S * S
    fps = B / (ms / 1e3)
    return {"config": "3: full warp + background + k-means + CNN pipeline, synthetic 1080p 30 fps video, 1 B200",
            "frames": steps * B, "frames_per_s": fps, "real_time_factor_at_30fps": fps / 30.0,
            "kmeans_label_accuracy_vs_truth": float((out["km"]["stones"].cpu().numpy() == truth).mean()),
            "roofline": {"mog2": {"bound": "hbm", "bytes_per_launch": mog_bytes, "achieved_gbs": mog_bytes / (mog_ms / 1e3) / 1e9,
                                  "frac": mog_bytes / (mog_ms / 1e3) / 1e9 / hbm},
                         "peak_gbs": hbm, "peak_source": src},
            "kernels_ms_per_64_frames": {k_: round(v[0], 4) for k_, v in agg.items()}}


def config4(n):
    H, W = 2160, 3840
    rows = []
    for gsize in (9, 13, 19):
        S = 20 * gsize
        eng = StoneEngine(gsize)
        frames, mtx, truth = clip(40 + gsize, n, H, W, gsize=gsize, distinct=4)
        d = torch.from_numpy(frames).cuda()
        goban = torch.empty((n, S, S, 3), dtype=torch.uint8, device="cuda")
        st0 = rng_seed(0)
        states = [rng_advance(st0, i) for i in range(n)]
        res = {}

        def run():
            eng.warp(d, mtx, out=goban)
            res["r"] = eng.find_stones(goban, states, want=("stones", "trusted", "labels"))

        ms = timed(run, 3)
        rows.append({"gsize": gsize, "frames_per_s": n / (ms / 1e3),
                     "label_accuracy_vs_truth": float((res["r"]["stones"].cpu().numpy() == truth).mean())})
        del eng
    return {"config": "4: 4K synthetic frames, mixed 9x9 / 13x13 / 19x19 boards, perspective jitter", "frames_per_size": n,
            "sizes": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--only", type=int, default=0)
    args = ap.parse_args()
    t0 = time.time()
    for k, fn, n in ((1, config1, args.frames), (3, config3, args.frames), (4, config4, 16)):
        if args.only and args.only != k:
            continue
        line = fn(n)
        line["gpu"] = torch.cuda.get_device_name(0)
        line["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
