"""Shared helpers for the parity tests: synthetic inputs and the glue that turns
the oracle's Params/Derived into the C-ABI parameter block (what the Rust side does in
build_uniforms, src/wgpu/mod.rs:661-692)."""
import ctypes as C

import numpy as np

from oracle import oracle as O


def gradient_u8(w: int, h: int) -> np.ndarray:
    """tools/film_grain_bench.py 'ramp' (:153-162): horizontal 0..1 ramp, 8-bit, 3 channels."""
    row = np.rint(np.tile(np.linspace(0, 1, w), (h, 1)) * 255).astype(np.uint8)
    return np.repeat(row[:, :, None], 3, axis=2)


def noise_u8(w: int, h: int, seed: int = 20240611) -> np.ndarray:
    """iid uniform 8-bit RGB noise (tool seed, tools/film_grain_bench.py:172)."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def lambda_from_u8(chan_u8: np.ndarray, inv_e_pi_r2: float) -> np.ndarray:
    """load (c/255, clamp) -> normalize_plane -> lambda_plane, all by the oracle."""
    plane = (chan_u8.astype(np.float32) / np.float32(255.0)).astype(np.float32)
    return O.lambda_plane(O.normalize_plane(plane), inv_e_pi_r2)


def fg_params_from(p: O.Params, d: O.Derived, path: int = 0, rows=None, seeding: int = 0):
    from film_grain_b200 import FgParams
    q = FgParams()
    q.struct_size = C.sizeof(FgParams)
    q.in_w, q.in_h = d.input_width, d.input_height
    q.out_w, q.out_h = d.output_width, d.output_height
    q.n_samples = p.n_samples
    q.dist_kind = p.radius_dist
    q.seeding = seeding
    q.seed = p.seed
    q.zoom = p.zoom
    q.delta = d.delta
    q.rm = d.rm
    q.inv_e_pi_r2 = d.inv_e_pi_r2
    q.radius_mean = p.radius_mean
    q.has_log = p.has_log
    q.radius_log_mu = float(np.float32(p.radius_log_mu))
    q.radius_log_sigma = float(np.float32(p.radius_log_sigma))
    if rows is not None:
        q.row_begin, q.row_end = rows
    q.path = path
    return q
