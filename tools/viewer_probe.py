#!/usr/bin/env python3
"""Latencies of the viewer path on a 4K RGB frame (BASELINE config 2 geometry), through the host ABI:
cold render, re-renders from the cached cell table (N / sigma / zoom change), progressive refinement 16 -> 64 -> 256, and
the time from raising the cancel flag to the call's return.  Also: the cell-table pass with the three planes generated
jointly (k_gen_rows<., 3>) against per plane, on independent-noise planes and on a grey image (R = G = B).

usage: python tools/viewer_probe.py > gpurun_out/viewer_probe.json"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import film_grain_b200 as fg  # noqa: E402
from film_grain_b200 import host as H  # noqa: E402


def params(n, sigma=0.8, zoom=1.0, seed=5489):
    return H.ParamsBuilder(radius_mean=0.1, n_samples=n, sigma_px=sigma, zoom=zoom, algo=H.Algo.Pixel, seed=seed).build()


def timed(f, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = f()
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    return best, r


def main():
    w, h = 3840, 2160
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    out = {"frame": f"{w}x{h} RGB noise r=0.1"}

    def setup(p, image):
        d = H.derive_common(p, (w, h))
        lams = [H.lambda_plane(image[:, :, k].astype(np.float32) / np.float32(255.0), d.inv_e_pi_r2) for k in range(3)]
        return d, lams

    with fg.Context(0) as c:
        p = params(256)
        d, lams = setup(p, img)
        q = d.block
        offs = d.offsets_input
        outs = [np.zeros((q.out_h, q.out_w), np.float32) for _ in range(3)]
        c.render_planes(q, 2, lams, offs, outs)
        out["cold_N256_ms"], _ = timed(lambda: c.render_planes(q, 2, lams, offs, outs))
        st = c.stats()
        out["cold_N256_table_ms"] = st.table_ms
        c.set_table_cache(True)
        c.render_planes(q, 2, lams, offs, outs)  # builds and keeps the table
        out["cached_N256_ms"], _ = timed(lambda: c.render_planes(q, 2, lams, offs, outs))
        out["cached_N256_reused"] = int(c.stats().table_reused)
        for n in (16, 64):
            pn = params(n)
            dn = H.derive_common(pn, (w, h))
            out[f"cached_N{n}_ms"], _ = timed(lambda: c.render_planes(dn.block, 2, lams, dn.offsets_input, outs))
            out[f"cached_N{n}_reused"] = int(c.stats().table_reused)
        ps = params(256, sigma=0.6)
        ds = H.derive_common(ps, (w, h))
        out["cached_sigma0.6_N256_ms"], _ = timed(lambda: c.render_planes(ds.block, 2, lams, ds.offsets_input, outs))
        out["cached_sigma0.6_reused"] = int(c.stats().table_reused)
        # progressive refinement
        steps = []
        k0 = 0
        for k1 in (16, 64, 256):
            t0 = time.perf_counter()
            c.refine_planes(q, 2, lams, offs, k0, k1, outs)
            steps.append({"samples": [k0, k1], "ms": (time.perf_counter() - t0) * 1e3, "table_reused": int(c.stats().table_reused)})
            k0 = k1
        out["refine"] = steps
        c.set_table_cache(False)
        # cancel latency: flag raised 15 ms into a cold render
        flag = C.c_int(0)
        c.set_cancel_flag(flag)
        lat = []
        for delay in (0.005, 0.015, 0.030):
            flag.value = 0
            t_raise = [0.0]

            def raise_flag():
                t_raise[0] = time.perf_counter()
                flag.value = 1
            th = threading.Timer(delay, raise_flag)
            th.start()
            try:
                c.render_planes(q, 2, lams, offs, outs)
                lat.append({"raised_after_ms": delay * 1e3, "cancelled": False})
            except fg.Cancelled:
                lat.append({"raised_after_ms": delay * 1e3, "cancelled": True, "return_after_raise_ms": (time.perf_counter() - t_raise[0]) * 1e3})
            th.join()
        flag.value = 0
        c.set_cancel_flag(None)
        out["cancel"] = lat
    # joint vs per-plane table generation
    grey = np.repeat(img[:, :, :1], 3, axis=2)
    gen = {}
    for name, image in (("independent_noise", img), ("grey_R=G=B", grey)):
        for joint in ("1", "0"):
            os.environ["FG_B200_GEN_JOINT"] = joint
            with fg.Context(0) as c:
                p = params(256)
                d, lams = setup(p, image)
                outs = [np.zeros((d.block.out_h, d.block.out_w), np.float32) for _ in range(3)]
                c.render_planes(d.block, 2, lams, d.offsets_input, outs)
                best = 1e9
                for _ in range(3):
                    c.render_planes(d.block, 2, lams, d.offsets_input, outs)
                    best = min(best, c.stats().table_ms)
                gen[f"{name}_joint{joint}_table_ms"] = best
    os.environ.pop("FG_B200_GEN_JOINT", None)
    out["table_pass"] = gen
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
