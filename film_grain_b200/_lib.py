"""ctypes loader for libfg_b200.so (the C-ABI engine declared in include/fg.h).

The product path has no CPU fallback: if the shared library is missing or does not export
the ABI, importing a render entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("FG_B200_LIB") or os.path.join(HERE, "libfg_b200.so")  # override: kernel experiments

FG_OK, FG_ERR_INVALID, FG_ERR_OOM, FG_ERR_CUDA_STICKY = 0, -1, -2, -3
FG_ERR_NO_DEVICE, FG_ERR_CANCELLED, FG_ERR_CUDA = -4, -5, -6
FG_DIST_CONST, FG_DIST_LOGNORM = 0, 1
FG_STREAM_CELL, FG_STREAM_PIXEL = 1, 2
FG_COLOR_LUMA, FG_COLOR_RGB = 0, 1
FG_ALGO_GRAIN, FG_ALGO_PIXEL = 1, 2
FG_PATH_AUTO, FG_PATH_DIRECT, FG_PATH_TILED, FG_PATH_STAGED = 0, 1, 2, 3


class FgParams(C.Structure):
    """struct fg_params (include/fg.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("in_w", C.c_uint32), ("in_h", C.c_uint32),
                ("out_w", C.c_uint32), ("out_h", C.c_uint32), ("n_samples", C.c_uint32),
                ("dist_kind", C.c_uint32), ("seeding", C.c_uint32), ("seed", C.c_uint64),
                ("zoom", C.c_float), ("delta", C.c_float), ("rm", C.c_float), ("inv_e_pi_r2", C.c_float),
                ("radius_mean", C.c_float), ("has_log", C.c_uint32), ("radius_log_mu", C.c_double),
                ("radius_log_sigma", C.c_double), ("row_begin", C.c_uint32), ("row_end", C.c_uint32),
                ("path", C.c_uint32), ("reserved", C.c_uint32)]


class FgStats(C.Structure):
    """struct fg_stats (include/fg.h)."""
    _fields_ = [("kernel_ms", C.c_float), ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("launches", C.c_uint32), ("tiles_total", C.c_uint32), ("tiles_fallback", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("strip_ms", C.c_float), ("strip_launches", C.c_uint32),
                ("table_ms", C.c_float), ("table_reused", C.c_uint32)]


class FghParams(C.Structure):
    """struct fgh_params (include/fg_host.h) = ParamsBuilder (src/params.rs:70-91)."""
    _fields_ = [("radius_dist", C.c_int32), ("radius_mean", C.c_float), ("radius_stddev", C.c_float),
                ("zoom", C.c_float), ("sigma_px", C.c_float), ("n_samples", C.c_uint32), ("algo", C.c_int32),
                ("max_radius_kind", C.c_int32), ("max_radius_value", C.c_float), ("has_cell_delta", C.c_int32),
                ("cell_delta", C.c_float), ("color_mode", C.c_int32), ("has_size", C.c_int32),
                ("size_w", C.c_uint32), ("has_size_h", C.c_int32), ("size_h", C.c_uint32), ("seed", C.c_uint64)]


class FghDerived(C.Structure):
    """struct fgh_derived (include/fg_host.h) = Derived (src/model.rs:167-179) + resolved algorithm."""
    _fields_ = [("input_width", C.c_uint64), ("input_height", C.c_uint64), ("output_width", C.c_uint64),
                ("output_height", C.c_uint64), ("inv_e_pi_r2", C.c_float), ("rm", C.c_float), ("delta", C.c_float),
                ("radius_stddev", C.c_float), ("has_log", C.c_int32), ("algorithm", C.c_int32),
                ("log_mu", C.c_double), ("log_sigma", C.c_double), ("block", FgParams)]


# every symbol include/fg.h declares: name -> (restype, argtypes)
_P = C.POINTER
_VP = C.c_void_p
ABI = {
    "fg_abi_version": (C.c_int, []),
    "fg_device_count": (C.c_int, []),
    "fg_error_string": (C.c_char_p, [C.c_int]),
    "fg_context_create": (C.c_int, [_P(_VP), C.c_int]),
    "fg_context_create_multi": (C.c_int, [_P(_VP), _P(C.c_int), C.c_int]),
    "fg_context_device_count": (C.c_int, [_VP]),
    "fg_context_destroy": (None, [_VP]),
    "fg_last_error": (C.c_char_p, [_VP]),
    "fg_last_eval_kernel": (C.c_char_p, [_VP]),
    "fg_set_cancel_flag": (None, [_VP, _P(C.c_int)]),
    "fg_get_stats": (None, [_VP, _P(FgStats)]),
    "fg_render_pixelwise": (C.c_int, [_VP, _P(FgParams), _VP, _VP, _VP]),
    "fg_render_grainwise": (C.c_int, [_VP, _P(FgParams), _VP, _VP, _VP]),
    "fg_render_planes": (C.c_int, [_VP, _P(FgParams), C.c_int, C.c_int, _P(_VP), _VP, _P(_VP)]),
    "fg_render_planes_cancelable": (C.c_int, [_VP, _P(FgParams), C.c_int, C.c_int, _P(_VP), _VP, _P(_VP), _P(C.c_int)]),
    "fg_set_table_cache": (None, [_VP, C.c_int]),
    "fg_refine_planes": (C.c_int, [_VP, _P(FgParams), C.c_int, C.c_int, _P(_VP), _VP, C.c_uint32, C.c_uint32, _P(_VP), _P(C.c_int)]),
    "fg_render_planes_device": (C.c_int, [_VP, _P(FgParams), C.c_int, C.c_int, _VP, _VP, _VP, C.c_int]),
    "fg_context_stream": (C.c_uint64, [_VP]),
    "fg_context_synchronize": (C.c_int, [_VP]),
    "fg_render_rgb8": (C.c_int, [_VP, _P(FgParams), C.c_int, C.c_int, _VP, _VP, _VP]),
    "fg_render_rgb8_device": (C.c_int, [_VP, _P(FgParams), C.c_int, C.c_int, _VP, _VP, _VP, C.c_int]),
    "fg_dump_cells": (C.c_int, [_VP, _P(FgParams), C.c_int, _VP, _VP, C.c_size_t, C.c_uint32, _VP, _VP]),
    "fg_measure_issue_peak": (C.c_int, [_VP, _P(C.c_double)]),
}
# include/fg_host.h (host-side mirror of the reference's library API)
HOST_ABI = {
    "fgh_last_error": (C.c_char_p, []),
    "fgh_derive": (C.c_int, [_P(FghParams), C.c_uint64, C.c_uint64, _P(FghDerived), _VP, _VP]),
    "fgh_lambda_from_plane": (C.c_int, [_VP, C.c_uint64, C.c_uint64, C.c_float, _VP]),
    "fgh_render_with_input_image": (C.c_int, [_P(FghParams), _VP, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                              _P(C.c_int), _VP, C.c_uint64, _P(FghDerived)]),
    "fgh_context": (_VP, [C.c_int]),
    "fgh_invalidate_context": (None, []),
    "fgh_logf_restated": (None, [_VP, C.c_uint64, _VP]),
    "fgh_render_file": (C.c_int, [_P(FghParams), C.c_char_p, C.c_char_p, C.c_char_p, _P(C.c_uint32), C.c_int, C.c_int,
                                  _P(C.c_int), _P(FghDerived)]),
    "fgh_load_image": (C.c_int, [C.c_char_p, _P(_VP), _P(C.c_uint64), _P(C.c_uint64)]),
    "fgh_free": (None, [_VP]),
    "fgh_save_image": (C.c_int, [C.c_char_p, _VP, C.c_uint64, C.c_uint64, C.c_char_p]),
}

_lib = None


class EngineMissing(ImportError):
    pass


def load() -> C.CDLL:
    """dlopen the engine; raises EngineMissing (never falls back to a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise EngineMissing(f"{SO_PATH} not built: run `python film_grain_b200/build.py` "
                            "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in list(ABI.items()) + list(HOST_ABI.items()):
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise EngineMissing(f"{SO_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.fg_abi_version() != 1:
        raise EngineMissing("libfg_b200.so ABI version mismatch")
    _lib = lib
    return lib
