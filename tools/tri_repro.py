#!/usr/bin/env python3
"""Reproduce one k_pixelwise_tri case (debug aid; run under compute-sanitizer)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("FG_B200_TRI_MIN_HEADROOM", "0.05")
from oracle import oracle as O
from tests.helpers import fg_params_from, lambda_from_u8, noise_u8
import film_grain_b200 as fg
w, h = 96, 80
p = O.make_params(algo=O.ALGO_PIXEL, radius=0.1, n_samples=160)
d, off, off_in = O.derive_common(p, w, h)
img = noise_u8(w, h, seed=23)
img[20:60, 30:, :] = 255
lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
ctx = fg.Context(0)
got = ctx.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
print(ctx.eval_kernel_name(), float(got.mean()), "lam max", float(lam.max()), float(lam.max()) * d.delta * d.delta)
ref = O.render_pixelwise(lam, p, d, off_in)
print("equal", np.array_equal(ref, got))
