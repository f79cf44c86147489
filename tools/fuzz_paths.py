#!/usr/bin/env python3
"""Randomised parity check on a GPU box: for random parameter sets (radius, distribution, zoom, N, cell size,
seed, image content) every kernel family must equal the CPU ORACLE bit for bit -- pixel-wise direct / tiled /
staged / auto, grain-wise global mask / tile / auto -- on the full render and on a random row band.  Prints one
line per mismatch and a summary; exit code 1 on any mismatch or CUDA error.
usage: python tools/fuzz_paths.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import film_grain_b200 as fg
    from oracle import oracle as O
    from tests.helpers import fg_params_from, lambda_from_u8

    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    ctx = fg.Context(0)
    t0, n_cases, bad = time.time(), 0, 0
    while time.time() - t0 < budget:
        w, h = int(rng.integers(40, 420)), int(rng.integers(30, 300))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        kind = rng.integers(0, 4)
        if kind == 1:
            img = np.tile(np.rint(np.linspace(0, 255, w)).astype(np.uint8), (h, 1))
        elif kind == 2:
            img[h // 3:, w // 4:] = 255
        elif kind == 3:
            img[:, :] = int(rng.integers(0, 256))
        radius = float(rng.choice([0.05, 0.07, 0.1, 0.12, 0.15, 0.2, 0.3, 0.5, 0.8, 1.3]))
        kw = dict(radius=radius, n_samples=int(rng.choice([1, 2, 5, 16, 33, 64, 100, 130, 200, 257])),
                  zoom=float(rng.choice([0.6, 1.0, 1.0, 1.0, 1.5, 2.0, 3.0])), seed=int(rng.integers(0, 2**32)))
        if rng.random() < 0.3:
            kw.update(radius_dist=O.DIST_LOGNORM, radius_stddev=radius * float(rng.choice([0.2, 0.5])))
        algo = "pixel" if rng.random() < 0.65 else "grain"
        if algo == "pixel" and rng.random() < 0.2:
            kw["cell_delta"] = float(rng.choice([0.05, 0.33, 0.7]))
        if algo == "grain" and radius < 0.1:
            kw["n_samples"] = min(kw["n_samples"], 33)
        try:
            p = O.make_params(algo=O.ALGO_PIXEL if algo == "pixel" else O.ALGO_GRAIN, **kw)
            d, off, off_in = O.derive_common(p, w, h)
            if d.output_width * d.output_height * p.n_samples > 8e6:  # keeps the oracle (the reference of every case) well under a second
                continue
            lam = lambda_from_u8(img, d.inv_e_pi_r2)
            oh = d.output_height
            a = int(rng.integers(0, oh))
            b = int(rng.integers(a + 1, oh + 1))
            if algo == "pixel":
                ref = O.render_pixelwise(lam, p, d, off_in)  # the oracle is the reference, not another GPU path
                for path in (1, 2, 3, 0):
                    got = ctx.render_pixelwise(fg_params_from(p, d, path=path), lam, off_in)
                    band = np.array(ref)
                    ctx.render_pixelwise(fg_params_from(p, d, path=path, rows=(a, b)), lam, off_in, out=band)
                    if not (np.array_equal(got, ref) and np.array_equal(band, ref)):
                        bad += 1
                        print("MISMATCH pixel path", path, w, h, kw, (a, b), flush=True)
                if rng.random() < 0.5:  # three planes at once: the joint table generation (k_gen_rows<., 3>), each plane against the oracle
                    img2 = rng.integers(0, 256, (h, w), dtype=np.uint8)
                    if rng.random() < 0.3:
                        img2[: h // 2] = 255
                    lam2 = lambda_from_u8(img2, d.inv_e_pi_r2)
                    ref2 = O.render_pixelwise(lam2, p, d, off_in)
                    outs = ctx.render_planes(fg_params_from(p, d, path=3), O.ALGO_PIXEL, [lam, lam2, lam], off_in)
                    if not (np.array_equal(outs[0], ref) and np.array_equal(outs[1], ref2) and np.array_equal(outs[2], ref)):
                        bad += 1
                        print("MISMATCH pixel 3-plane", w, h, kw, flush=True)
            else:
                ref = O.render_grainwise(lam, p, d, off)
                for path in (1, 3, 0):
                    got = ctx.render_grainwise(fg_params_from(p, d, path=path), lam, off)
                    band = np.array(ref)
                    ctx.render_grainwise(fg_params_from(p, d, path=path, rows=(a, b)), lam, off, out=band)
                    if not (np.array_equal(got, ref) and np.array_equal(band, ref)):
                        bad += 1
                        print("MISMATCH grain path", path, w, h, kw, (a, b), flush=True)
            n_cases += 1
        except Exception as e:  # a CUDA fault poisons the context: report and stop
            print("ERROR", algo, w, h, kw, repr(e)[:300], flush=True)
            bad += 1
            break
    print(f"fuzz: {n_cases} cases (each against the CPU oracle), {bad} failures, {time.time() - t0:.1f} s", flush=True)
    if not bad:
        ctx.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
