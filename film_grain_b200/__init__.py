"""film_grain_b200 -- B200-native engine for the Monte-Carlo film-grain hot path of
joseph-wardle/film_grain (pixel-wise and grain-wise Boolean-model integrators).

Layout: csrc/ (CUDA kernels + the C ABI of include/fg.h), _lib.py (ctypes loader),
engine.py (context handle), host.py (host-side mirror of the reference's library API).
"""
from ._lib import (FG_ALGO_GRAIN, FG_ALGO_PIXEL, FG_COLOR_LUMA, FG_COLOR_RGB, FG_DIST_CONST, FG_DIST_LOGNORM,
                   FG_PATH_AUTO, FG_PATH_DIRECT, FG_PATH_TILED, FG_PATH_STAGED, FG_STREAM_CELL, FG_STREAM_PIXEL, EngineMissing,
                   FgParams, FgStats)
from .engine import Cancelled, Context, GpuError, device_count

__all__ = ["Context", "GpuError", "Cancelled", "FgParams", "FgStats", "EngineMissing", "device_count"]
