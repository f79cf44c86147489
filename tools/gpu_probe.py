"""Quick on-GPU probe: issue-rate peaks, and timing of one pixel-wise render (used during bring-up)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import film_grain_b200 as fg
from oracle import oracle as O
from tests.helpers import fg_params_from, gradient_u8, noise_u8, lambda_from_u8

ctx = fg.Context(0)
print("issue peak (Glane-op/s):", ctx.measure_issue_peak())
for (w, h, n, path) in [(512, 512, 64, 1), (512, 512, 64, 2), (1024, 1024, 64, 2)]:
    p = O.make_params(radius=0.1, n_samples=n, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h)[:, :, 0], d.inv_e_pi_r2)
    q = fg_params_from(p, d, path=path)
    for it in range(2):
        t = time.time(); out = ctx.render_pixelwise(q, lam, off_in); dt = time.time() - t
    s = ctx.stats()
    print(f"{w}x{h} N={n} path={path}: kernel {s.kernel_ms:.2f} ms  e2e {dt*1e3:.2f} ms  "
          f"{w*h*n/s.kernel_ms/1e3:.1f} Mpx-smp/s  launches {s.launches} tiles {s.tiles_total} fb {s.tiles_fallback} mean {out.mean():.4f}")
