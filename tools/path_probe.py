#!/usr/bin/env python3
"""Small-N probe (run on a GPU box): device-resident pixel-wise render of a 4K luma noise plane at N = 1..64
through each kernel family (FG_PATH_DIRECT / TILED / STAGED) and through FG_PATH_AUTO; prints ms per render.  Used to place the
crossover below which the cell table (whose cost does not depend on N) stops paying."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import film_grain_b200 as fg
    from film_grain_b200 import host as H

    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
    img = np.random.default_rng(20240611).integers(0, 256, (h, w), dtype=np.uint8)
    ctx = fg.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for radius in (0.1, 0.05):
        for n in (1, 2, 4, 8, 16, 32, 64):
            params = H.ParamsBuilder(radius_mean=radius, n_samples=n, algo=H.Algo.Pixel, color_mode=H.ColorMode.Luma).build()
            d = H.derive_common(params, (w, h))
            lam = H.lambda_plane((img.astype(np.float32) / np.float32(255.0)).astype(np.float32), d.inv_e_pi_r2)
            d_lam = torch.from_numpy(lam).to(dev)
            d_off = torch.from_numpy(np.ascontiguousarray(d.offsets_input)).to(dev)
            d_out = torch.zeros((d.output_height, d.output_width), dtype=torch.float32, device=dev)
            row = []
            for path in (fg.FG_PATH_DIRECT, fg.FG_PATH_TILED, fg.FG_PATH_STAGED, fg.FG_PATH_AUTO):
                blk = H._band(d.block, None)
                blk.path = path
                with torch.cuda.stream(stream):
                    for _ in range(2):
                        ctx.render_planes_device(blk, fg.FG_ALGO_PIXEL, 1, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    for _ in range(3):
                        ctx.render_planes_device(blk, fg.FG_ALGO_PIXEL, 1, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
                    b.record(stream)
                    stream.synchronize()
                row.append(a.elapsed_time(b) / 3)
            print(f"r={radius} N={n:3d}  direct {row[0]:8.2f} ms   tiled {row[1]:8.2f} ms   staged {row[2]:8.2f} ms   auto {row[3]:8.2f} ms", flush=True)
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
