#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz.

The reference (Rust) cannot be built or imported in this image and ships no fixtures of its own
(SURVEY.md F2, F3), so these vectors are produced by the CPU oracle -- whose random arithmetic is
pinned by the crates' published known-answer vectors (tests/test_oracle_kat.py).  They freeze the
oracle's behaviour (any later change to it shows up as a diff here) and give the GPU tests a
committed target.  Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.helpers import gradient_u8, lambda_from_u8, noise_u8  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # 1. make_offsets (src/rng.rs:9-24)
    np.savez_compressed(os.path.join(HERE, "offsets.npz"),
                        seed5489_n64_s08=O.make_offsets(5489, 64, 0.8),
                        seed12345678901234_n7_s05=O.make_offsets(12345678901234, 7, 0.5),
                        seed0_n3_s2=O.make_offsets(0, 3, 2.0))
    # 2. grain realisations per cell (both streams; const and lognormal radii; both Poisson branches)
    rng = np.random.default_rng(2024)
    out = {}
    for tag, kw in (("const", dict(radius=0.1, n_samples=1, seed=5489)),
                    ("lognorm", dict(radius=0.1, radius_dist=O.DIST_LOGNORM, radius_stddev=0.05, n_samples=1, seed=77))):
        p = O.make_params(**kw)
        d, _, _ = O.derive_common(p, 16, 16)
        for sname, stream, scale in (("cell", O.STREAM_CELL, 1.0 / (d.delta * d.delta)), ("pixel", O.STREAM_PIXEL, 1.0)):
            n = 1500
            ij = rng.integers(-5000, 5000, (n, 2)).astype(np.int32)
            lam = (np.concatenate([rng.uniform(0, 5, n - 300), rng.uniform(11, 40, 300)]) * scale).astype(np.float32)
            q, g = O.gen_cells(p, d, stream, ij, lam, 6)
            out[f"{tag}_{sname}_ij"] = ij
            out[f"{tag}_{sname}_lam"] = lam
            out[f"{tag}_{sname}_q"] = q
            out[f"{tag}_{sname}_g"] = g
    np.savez_compressed(os.path.join(HERE, "cells.npz"), **out)
    # 3. pixel-wise plane (src/pixelwise.rs) -- a small cousin of BASELINE configs[0]
    w, h = 64, 48
    p = O.make_params(radius=0.1, n_samples=16, algo=O.ALGO_PIXEL, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(gradient_u8(w, h)[:, :, 0], d.inv_e_pi_r2)
    np.savez_compressed(os.path.join(HERE, "pixelwise_64x48_r0.1_N16.npz"), lam=lam, offsets_input=off_in,
                        out=O.render_pixelwise(lam, p, d, off_in))
    p = O.make_params(radius=0.05, n_samples=8, zoom=2.5, algo=O.ALGO_PIXEL, seed=99)
    d, off, off_in = O.derive_common(p, 40, 28)
    lam = lambda_from_u8(noise_u8(40, 28, seed=4)[:, :, 0], d.inv_e_pi_r2)
    np.savez_compressed(os.path.join(HERE, "pixelwise_40x28_zoom2.5_r0.05_N8.npz"), lam=lam, offsets_input=off_in,
                        out=O.render_pixelwise(lam, p, d, off_in))
    # 4. grain-wise plane (src/grainwise.rs)
    p = O.make_params(radius=0.5, n_samples=16, algo=O.ALGO_GRAIN, seed=5489)
    d, off, off_in = O.derive_common(p, 40, 40)
    lam = lambda_from_u8(noise_u8(40, 40, seed=6)[:, :, 0], d.inv_e_pi_r2)
    np.savez_compressed(os.path.join(HERE, "grainwise_40x40_r0.5_N16.npz"), lam=lam, offsets=off,
                        out=O.render_grainwise(lam, p, d, off))
    # 5. whole pipeline on 8-bit images (src/lib.rs:134-173), both colour modes
    img = noise_u8(24, 20, seed=8)
    p = O.make_params(radius=0.1, n_samples=16, zoom=1.5, algo=O.ALGO_PIXEL, seed=5489)
    rgb, _ = O.render_rgb8(img, p, 1)
    luma, _ = O.render_rgb8(img, p, 0)
    np.savez_compressed(os.path.join(HERE, "rgb8_24x20_zoom1.5.npz"), img=img, out_rgb=rgb, out_luma=luma)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
