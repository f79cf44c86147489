import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import film_grain_b200 as fg
from film_grain_b200 import host as H
from tools.bench_sweep_b200 import intensity_field
ctx = fg.Context(0)
for (m, N, mu, ratio, s, path) in [(512,4096,0.1,0.5,1,2),(512,4096,0.1,0.5,1,1),(512,256,0.1,0.5,4,2),(1024,256,0.1,0.25,1,2),(1024,256,0.1,1.0,1,2),(1024,256,0.1,0.0,1,2)]:
    img = intensity_field('natural', m)
    p = H.ParamsBuilder(radius_mean=mu, n_samples=N, zoom=float(s), algo=H.Algo.Pixel, radius_dist=H.RadiusDist.Lognorm if ratio>0 else H.RadiusDist.Const, radius_stddev=mu*ratio).build()
    d = H.derive_common(p, (m, m))
    lam = H.lambda_plane((img[:,:,0].astype(np.float32)/np.float32(255)).astype(np.float32), d.inv_e_pi_r2)
    blk = H._band(d.block, None); blk.path = path
    for _ in range(2): out = ctx.render_pixelwise(blk, lam, d.offsets_input)
    st = ctx.stats()
    print(f"m={m} N={N} ratio={ratio} zoom={s} path={path}: kernel {st.kernel_ms:.2f} ms rm={d.rm:.3f} tiles {st.tiles_total} fb {st.tiles_fallback} -> {m*s*m*s*N/st.kernel_ms/1e3:.0f} Mpx-smp/s")
