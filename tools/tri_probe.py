#!/usr/bin/env python3
"""Render one plane of BASELINE configs[1] through the default path and print the engine statistics (debug aid for
k_pixelwise_tri: build with FG_NVCC_EXTRA=-DFG_TRI_DEBUG to see why segments go to the fallback list)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O
from tests.helpers import fg_params_from, lambda_from_u8, noise_u8
import film_grain_b200 as fg

w, h, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160, 256)
p = O.make_params(radius=0.1, n_samples=n, algo=O.ALGO_PIXEL, seed=5489)
d, off, off_in = O.derive_common(p, w, h)
img = noise_u8(w, h)
lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
ctx = fg.Context(0)
full = ctx.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
st = ctx.stats()
print("tiles", st.tiles_total, "fallback", st.tiles_fallback, "kernel", ctx.last_eval_kernel() if hasattr(ctx, "last_eval_kernel") else "?", "mean", float(full.mean()))
ctx.close()
