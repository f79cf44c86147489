#!/usr/bin/env python3
"""Sweep in the reference benchmark tool's CSV schema (tools/film_grain_bench.py:53-67 of
joseph-wardle/film_grain), with device = "b200", so tools/highsample_analysis.py of the reference can
read our numbers next to its own rows.

Differences from the reference tool, on purpose: it times whole process launches (PNG decode/encode
and start-up included); here `runtime_seconds` is the in-process time of `render_with_input_image`
(8-bit image in, 8-bit image out; load / lambda / store fused on the device, or the reference's
host data flow of lib.rs:134-173 with --host-color) on a decoded image, median of `--repeats` calls after one warm-up.  Inputs are the tool's synthetic
intensity fields (constant / step / ramp / natural, :147-177).

usage (on a GPU box):  python tools/bench_sweep_b200.py --out gpurun_out/b200_benchmark_results.csv
"""
import argparse
import csv
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CSV_HEADER = ("algorithm", "device", "thread_mode", "m", "n", "N", "mu_r", "sigma_r_ratio", "s",
              "intensity_pattern", "alpha", "delta", "runtime_seconds")


def bilinear_resize(field, size):
    src_h, src_w = field.shape
    xs = np.linspace(0, src_w - 1, num=size)
    inter = np.stack([np.interp(xs, np.arange(src_w, dtype=np.float64), field[r]) for r in range(src_h)])
    ys = np.linspace(0, src_h - 1, num=size)
    return np.stack([np.interp(ys, np.arange(src_h, dtype=np.float64), inter[:, c]) for c in range(size)], axis=1)


def intensity_field(pattern: str, size: int) -> np.ndarray:
    if pattern == "constant":
        f = np.full((size, size), 0.5)
    elif pattern == "step":
        f = np.zeros((size, size))
        f[size // 2:, :] = 1.0
    elif pattern == "ramp":
        f = np.tile(np.linspace(0.0, 1.0, num=size), (size, 1))
    elif pattern == "natural":
        base = np.clip(np.random.default_rng(20240611).normal(0.5, 0.2, size=(128, 128)), 0.0, 1.0)
        f = base if size == 128 else bilinear_resize(base, size)
    else:
        raise ValueError(pattern)
    g = np.rint(np.clip(f, 0.0, 1.0) * 255.0).astype(np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2)


def fmt(v):
    return f"{v:.6e}" if (abs(v) >= 1e6 or 0 < abs(v) < 1e-3) else f"{v:.10g}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="b200_benchmark_results.csv")
    ap.add_argument("--resolutions", type=int, nargs="+", default=[256, 512, 1024])
    ap.add_argument("--N", type=int, nargs="+", default=[16, 64, 256, 1024, 4096])
    ap.add_argument("--mu", type=float, nargs="+", default=[0.1, 0.5])
    ap.add_argument("--sigma-ratio", type=float, nargs="+", default=[0.0, 0.5])
    ap.add_argument("--zoom", type=int, nargs="+", default=[1, 4])
    ap.add_argument("--patterns", nargs="+", default=["constant", "ramp", "natural"])
    ap.add_argument("--algos", nargs="+", default=["pixel", "grain"])
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--host-color", action="store_true", help="host load/lambda/store (the reference data flow) instead of the fused device path")
    ap.add_argument("--budget-evals", type=float, default=6e9, help="skip configs above this many sample evaluations")
    args = ap.parse_args()

    from film_grain_b200 import host as H

    with open(args.out, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=CSV_HEADER)
        w.writeheader()
        for m in args.resolutions:
            for pattern in args.patterns:
                img = intensity_field(pattern, m)
                for algo in args.algos:
                    for N in args.N:
                        for mu in args.mu:
                            for ratio in args.sigma_ratio:
                                for s in args.zoom:
                                    if (m * s) ** 2 * N > args.budget_evals:
                                        continue
                                    pb = H.ParamsBuilder(radius_mean=mu, n_samples=N, zoom=float(s), sigma_px=0.8,
                                                         algo=H.Algo.Pixel if algo == "pixel" else H.Algo.Grain,
                                                         radius_dist=H.RadiusDist.Lognorm if ratio > 0 else H.RadiusDist.Const,
                                                         radius_stddev=mu * ratio, color_mode=H.ColorMode.Luma)
                                    p = pb.build()
                                    H.render_with_input_image(img, p, fused=not args.host_color)  # warm-up
                                    ts = []
                                    for rep in range(args.repeats):
                                        p = H.ParamsBuilder(**{**pb.__dict__, "seed": 5489 + rep}).build()
                                        t0 = time.perf_counter()
                                        H.render_with_input_image(img, p, fused=not args.host_color)
                                        ts.append(time.perf_counter() - t0)
                                    w.writerow({"algorithm": algo, "device": "b200", "thread_mode": "gpu", "m": m, "n": m, "N": N,
                                                "mu_r": fmt(mu), "sigma_r_ratio": fmt(ratio), "s": s, "intensity_pattern": pattern,
                                                "alpha": "", "delta": "", "runtime_seconds": fmt(statistics.median(ts))})
                                    fh.flush()
    print("wrote", args.out)


if __name__ == "__main__":
    main()
