#!/usr/bin/env python3
"""Time one 1/N row band of the C2 workload on one GPU (what a rank of an N-GPU run executes):
   python tools/band_probe.py [N]   -- prints kernel/strip/table ms; run under ncu for a launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import film_grain_b200 as fg
from film_grain_b200 import host as H
from tests.helpers import noise_u8

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3  # which of the n bands
w, h, N = 3840, 2160, 256
img = noise_u8(w, h)
p = H.ParamsBuilder(radius_mean=0.1, n_samples=N, algo=H.Algo.Pixel, seed=5489, color_mode=H.ColorMode.Rgb).build()
d = H.derive_common(p, (w, h))
lam = np.stack([H.lambda_plane((img[:, :, c].astype(np.float32) / np.float32(255.0)).astype(np.float32), d.inv_e_pi_r2) for c in range(3)])
dev = torch.device("cuda:0")
d_lam = torch.from_numpy(lam).to(dev)
d_off = torch.from_numpy(np.ascontiguousarray(d.offsets_input, np.float32)).to(dev)
d_out = torch.zeros((3, h, w), dtype=torch.float32, device=dev)
rows = h // n
with fg.Context(0) as ctx:
    blk = d.block
    blk.row_begin, blk.row_end = which * rows, (which + 1) * rows
    if os.environ.get("ROWS"):  # explicit band: ROWS=r0:r1
        blk.row_begin, blk.row_end = [int(v) for v in os.environ["ROWS"].split(":")]
    for i in range(6):
        ctx.render_planes_device(blk, fg.FG_ALGO_PIXEL, 3, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=True)
        st = ctx.stats()
        print(f"rows {blk.row_begin}:{blk.row_end} kernel {st.kernel_ms:.3f} ms  strip {st.strip_ms:.3f}  table {st.table_ms:.3f}  launches {st.launches}")
