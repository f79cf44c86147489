"""ctypes binding of the CPU oracle (oracle/libfg_oracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (film_grain_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfg_oracle.so")

DIST_CONST, DIST_LOGNORM = 0, 1
ALGO_AUTO, ALGO_GRAIN, ALGO_PIXEL = 0, 1, 2
MAXR_ABSOLUTE, MAXR_QUANTILE = 0, 1
RNG_XOSHIRO, RNG_CHACHA12, RNG_PCG32 = 0, 1, 2
STREAM_OFFSET, STREAM_CELL, STREAM_PIXEL = 0, 1, 2


class Rng(C.Structure):
    _fields_ = [("kind", C.c_int), ("s", C.c_uint64 * 4), ("pcg_state", C.c_uint64),
                ("pcg_inc", C.c_uint64), ("cc_key", C.c_uint32 * 8), ("cc_counter", C.c_uint64),
                ("cc_buf", C.c_uint32 * 64), ("cc_index", C.c_int)]


class Params(C.Structure):
    _fields_ = [("radius_dist", C.c_int), ("radius_mean", C.c_float), ("radius_stddev", C.c_float),
                ("has_log", C.c_int), ("radius_log_mu", C.c_float), ("radius_log_sigma", C.c_float),
                ("zoom", C.c_float), ("sigma_px", C.c_float), ("n_samples", C.c_uint32),
                ("algo", C.c_int), ("max_radius_kind", C.c_int), ("max_radius_value", C.c_float),
                ("has_cell_delta", C.c_int), ("cell_delta", C.c_float), ("has_size", C.c_int),
                ("size_w", C.c_uint32), ("has_size_h", C.c_int), ("size_h", C.c_uint32),
                ("seed", C.c_uint64)]


class Derived(C.Structure):
    _fields_ = [("input_width", C.c_int64), ("input_height", C.c_int64), ("output_width", C.c_int64),
                ("output_height", C.c_int64), ("inv_e_pi_r2", C.c_float), ("rm", C.c_float),
                ("delta", C.c_float), ("radius_dist", C.c_int), ("mean_linear", C.c_float),
                ("has_log", C.c_int), ("log_mu", C.c_double), ("log_sigma", C.c_double)]


class Counters(C.Structure):
    _fields_ = [("sample_evals", C.c_uint64), ("cell_visits", C.c_uint64),
                ("grain_tests", C.c_uint64), ("grains_drawn", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile if the .so is missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in ("fg_oracle.c", "fg_oracle.h", "zig_tables.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u64, u32, i32, i64, f32, f64 = C.c_uint64, C.c_uint32, C.c_int32, C.c_int64, C.c_float, C.c_double
    P = C.POINTER
    sig = {
        "fgo_splitmix64": (u64, [u64]), "fgo_mix": (u64, [u64, u64]), "fgo_mix3": (u64, [u64, u64, i64, i64]),
        "fgo_stream_const": (u64, [C.c_int]),
        "fgo_seed_bytes_from_u64": (None, [u64, P(C.c_uint8)]),
        "fgo_xoshiro_from_seed": (None, [P(Rng), P(C.c_uint8)]),
        "fgo_xoshiro_from_state": (None, [P(Rng), P(u64)]),
        "fgo_small_rng_seed_from_u64": (None, [P(Rng), u64]),
        "fgo_small_rng_seed_from_u64_variant_b": (None, [P(Rng), u64]),
        "fgo_chacha12_from_seed": (None, [P(Rng), P(C.c_uint8)]),
        "fgo_std_rng_seed_from_u64": (None, [P(Rng), u64]),
        "fgo_pcg32_new": (None, [P(Rng), u64, u64]),
        "fgo_next_u64": (u64, [P(Rng)]), "fgo_next_u32": (u32, [P(Rng)]),
        "fgo_cell_rng": (None, [P(Rng), u64, i32, i32]), "fgo_pixel_rng": (None, [P(Rng), u64, i32, i32]),
        "fgo_set_seeding_variant": (None, [C.c_int]),
        "fgo_standard_f64": (f64, [P(Rng)]), "fgo_standard_f32": (f32, [P(Rng)]),
        "fgo_open01_f64": (f64, [P(Rng)]),
        "fgo_uniform_f32_scale": (f32, [f32, f32]), "fgo_uniform_f32_sample": (f32, [P(Rng), f32, f32]),
        "fgo_log_gamma_f64": (f64, [f64]),
        "fgo_poisson_f64_sample": (f64, [P(Rng), f64]), "fgo_poisson_f32_sample": (f32, [P(Rng), f32]),
        "fgo_standard_normal_f64": (f64, [P(Rng)]), "fgo_normal_f64_sample": (f64, [P(Rng), f64, f64]),
        "fgo_lognormal_f64_sample": (f64, [P(Rng), f64, f64]), "fgo_norm_inv_cdf": (f64, [f64]),
        "fgo_default_cell_delta": (f32, [f32]),
        "fgo_params_build": (C.c_int, [P(Params), C.c_char_p, C.c_size_t]),
        "fgo_derive_common": (C.c_int, [P(Params), i64, i64, P(Derived), P(f32), P(f32), C.c_char_p, C.c_size_t]),
        "fgo_make_offsets": (None, [u64, C.c_size_t, f32, P(f32)]),
        "fgo_choose_algorithm": (C.c_int, [P(Params), P(Derived)]),
        "fgo_normalize_plane": (f32, [P(f32), C.c_size_t, P(f32)]),
        "fgo_lambda_plane": (None, [P(f32), C.c_size_t, f32, P(f32)]),
        "fgo_resize_nearest": (None, [P(f32), i64, i64, i64, i64, P(f32)]),
        "fgo_render_pixelwise": (C.c_int, [P(f32), P(Params), P(Derived), P(f32), P(f32), i64, i64, C.c_int, P(Counters)]),
        "fgo_render_grainwise": (C.c_int, [P(f32), P(Params), P(Derived), P(f32), P(f32), C.c_int, P(Counters)]),
        "fgo_gen_cell": (u32, [P(Params), P(Derived), C.c_int, i32, i32, f32, P(f32), P(f32), P(f32), u32]),
        "fgo_gen_cells": (None, [P(Params), P(Derived), C.c_int, P(i32), P(f32), C.c_size_t, u32, P(u32), P(f32)]),
        "fgo_load_rgb_u8": (None, [P(C.c_uint8), C.c_size_t, P(f32), P(f32), P(f32)]),
        "fgo_load_luma_u8": (None, [P(C.c_uint8), C.c_size_t, P(f32), P(f32), P(f32)]),
        "fgo_store_rgb_u8": (None, [P(f32), P(f32), P(f32), C.c_size_t, P(C.c_uint8)]),
        "fgo_store_luma_u8": (None, [P(f32), P(f32), P(f32), C.c_size_t, P(C.c_uint8)]),
        "fgo_to_u8": (C.c_uint8, [f32]),
        "fgo_render_rgb8": (C.c_int, [P(C.c_uint8), i64, i64, P(Params), C.c_int, P(C.c_uint8), C.c_int,
                                      P(C.c_int), P(Counters), C.c_char_p, C.c_size_t]),
        "fgo_max_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class OracleError(RuntimeError):
    pass


def make_params(radius=0.1, radius_dist=DIST_CONST, radius_stddev=0.0, zoom=1.0, sigma_px=0.8,
                n_samples=32, algo=ALGO_AUTO, max_radius=("quantile", 0.999), cell_delta=None,
                size=None, seed=5489) -> Params:
    """ParamsBuilder{..}.build() (params.rs:141-180) with the CLI defaults (main.rs:94-246)."""
    p = Params()
    p.radius_dist = radius_dist
    p.radius_mean = radius
    p.radius_stddev = radius_stddev
    p.zoom = zoom
    p.sigma_px = sigma_px
    p.n_samples = n_samples
    p.algo = algo
    p.max_radius_kind = MAXR_ABSOLUTE if max_radius[0] == "absolute" else MAXR_QUANTILE
    p.max_radius_value = max_radius[1]
    if cell_delta is not None:
        p.has_cell_delta, p.cell_delta = 1, cell_delta
    if size is not None:
        p.has_size, p.size_w = 1, size[0]
        if size[1] is not None:
            p.has_size_h, p.size_h = 1, size[1]
    p.seed = seed
    msg = C.create_string_buffer(256)
    if lib().fgo_params_build(C.byref(p), msg, 256) != 0:
        raise OracleError(msg.value.decode())
    return p


def derive_common(p: Params, in_w: int, in_h: int):
    """model.rs:181-226 -> (Derived, offsets[N,2], offsets_input[N,2])."""
    d = Derived()
    n = int(p.n_samples)
    off = np.zeros((n, 2), np.float32)
    off_in = np.zeros((n, 2), np.float32)
    msg = C.create_string_buffer(256)
    if lib().fgo_derive_common(C.byref(p), in_w, in_h, C.byref(d), _fp(off), _fp(off_in), msg, 256) != 0:
        raise OracleError(msg.value.decode())
    return d, off, off_in


def choose_algorithm(p: Params, d: Derived) -> int:
    return lib().fgo_choose_algorithm(C.byref(p), C.byref(d))


def make_offsets(seed: int, n: int, sigma: float) -> np.ndarray:
    out = np.zeros((n, 2), np.float32)
    lib().fgo_make_offsets(seed, n, sigma, _fp(out))
    return out


def normalize_plane(plane: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(plane, np.float32)
    dst = np.empty_like(src)
    lib().fgo_normalize_plane(_fp(src), src.size, _fp(dst))
    return dst


def lambda_plane(norm: np.ndarray, inv_e_pi_r2: float) -> np.ndarray:
    src = np.ascontiguousarray(norm, np.float32)
    dst = np.empty_like(src)
    lib().fgo_lambda_plane(_fp(src), src.size, inv_e_pi_r2, _fp(dst))
    return dst


def render_pixelwise(lam: np.ndarray, p: Params, d: Derived, offsets_input: np.ndarray,
                     y0: int = 0, y1: int | None = None, nthreads: int = 0, counters: Counters | None = None):
    """pixelwise.rs:11-45; rows outside [y0,y1) are left 0."""
    lam = np.ascontiguousarray(lam, np.float32)
    assert lam.shape == (d.input_height, d.input_width)
    out = np.zeros((d.output_height, d.output_width), np.float32)
    oi = np.ascontiguousarray(offsets_input, np.float32)
    if y1 is None:
        y1 = d.output_height
    rc = lib().fgo_render_pixelwise(_fp(lam), C.byref(p), C.byref(d), _fp(oi), _fp(out), y0, y1, nthreads,
                                    C.byref(counters) if counters is not None else None)
    if rc:
        raise OracleError(f"render_pixelwise rc={rc}")
    return out


def render_grainwise(lam: np.ndarray, p: Params, d: Derived, offsets: np.ndarray, nthreads: int = 0,
                     counters: Counters | None = None):
    """grainwise.rs:12-124."""
    lam = np.ascontiguousarray(lam, np.float32)
    assert lam.shape == (d.input_height, d.input_width)
    out = np.zeros((d.output_height, d.output_width), np.float32)
    of = np.ascontiguousarray(offsets, np.float32)
    rc = lib().fgo_render_grainwise(_fp(lam), C.byref(p), C.byref(d), _fp(of), _fp(out), nthreads,
                                    C.byref(counters) if counters is not None else None)
    if rc:
        raise OracleError(f"render_grainwise rc={rc}")
    return out


def gen_cell(p: Params, d: Derived, stream: int, i: int, j: int, lam: float, cap: int = 64):
    cx = np.zeros(cap, np.float32)
    cy = np.zeros(cap, np.float32)
    r = np.zeros(cap, np.float32)
    q = lib().fgo_gen_cell(C.byref(p), C.byref(d), stream, i, j, lam, _fp(cx), _fp(cy), _fp(r), cap)
    n = min(q, cap)
    return q, cx[:n], cy[:n], r[:n]


def gen_cells(p: Params, d: Derived, stream: int, ij: np.ndarray, lam: np.ndarray, cap: int = 8):
    """batched gen_cell -> (q[n] u32, grains[n,cap,3] f32)."""
    ij = np.ascontiguousarray(ij, np.int32)
    lam = np.ascontiguousarray(lam, np.float32)
    n = ij.shape[0]
    q = np.zeros(n, np.uint32)
    g = np.zeros((n, cap, 3), np.float32)
    lib().fgo_gen_cells(C.byref(p), C.byref(d), stream, ij.ctypes.data_as(C.POINTER(C.c_int32)), _fp(lam), n, cap,
                        q.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(g))
    return q, g


def render_rgb8(rgb: np.ndarray, p: Params, color_mode: int, nthreads: int = 0, counters: Counters | None = None):
    """lib.rs:134-173 on a decoded 8-bit RGB image -> (out u8 [H,W,3], algo used)."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    d, _, _ = derive_common(p, w, h)
    out = np.zeros((d.output_height, d.output_width, 3), np.uint8)
    algo = C.c_int(0)
    msg = C.create_string_buffer(256)
    rc = lib().fgo_render_rgb8(_u8p(rgb), w, h, C.byref(p), color_mode, _u8p(out), nthreads, C.byref(algo),
                               C.byref(counters) if counters is not None else None, msg, 256)
    if rc:
        raise OracleError(msg.value.decode() or f"rc={rc}")
    return out, algo.value
