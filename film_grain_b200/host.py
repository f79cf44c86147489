"""Python face of the host-side mirror (film_grain_b200/host/film_grain.{hpp,cpp}, exported through
include/fg_host.h): the reference's library API for the hot path with the same names and meaning.

    ParamsBuilder(...).build() -> Params          src/params.rs:70-180
    derive_common(params, (w, h)) -> Derived      src/model.rs:181-226   (offsets via make_offsets, src/rng.rs:9-24)
    choose_algorithm(params, derived) -> Algo     src/choose.rs:4-26
    lambda_plane(normalize_plane(plane))          src/model.rs:228-265
    context() / render_pixelwise_gpu / render_grainwise_gpu   src/wgpu/mod.rs:84-86, 336-345, 473-482
    render_with_input_image(image_u8, params)     src/lib.rs:78-84, 134-173 (Device::Gpu)

All arithmetic happens in the C++ mirror / the CUDA engine; this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum
import os

import numpy as np

from . import _lib
from ._lib import FghDerived, FghParams, FgParams
from .engine import Cancelled, Context, GpuError


class RadiusDist(enum.IntEnum):
    Const = 0
    Lognorm = 1


class Algo(enum.IntEnum):
    Auto = 0
    Grain = 1
    Pixel = 2


class ColorMode(enum.IntEnum):
    Luma = 0
    Rgb = 1


class ParamsError(ValueError):
    """ParamsError (src/params.rs:116-139)."""


class RenderError(RuntimeError):
    """RenderError::Message (src/lib.rs:42-43)."""


@dataclasses.dataclass
class ParamsBuilder:
    """ParamsBuilder (src/params.rs:70-91) with the CLI defaults (src/main.rs:94-246)."""
    radius_dist: RadiusDist = RadiusDist.Const
    radius_mean: float = 0.10
    radius_stddev: float = 0.0
    zoom: float = 1.0
    sigma_px: float = 0.8
    n_samples: int = 32
    algo: Algo = Algo.Auto
    max_radius: tuple = ("quantile", 0.999)
    cell_delta: float | None = None
    color_mode: ColorMode = ColorMode.Luma
    size: tuple | None = None
    seed: int = 5489

    def _c(self) -> FghParams:
        p = FghParams()
        p.radius_dist = int(self.radius_dist)
        p.radius_mean = self.radius_mean
        p.radius_stddev = self.radius_stddev
        p.zoom = self.zoom
        p.sigma_px = self.sigma_px
        p.n_samples = self.n_samples
        p.algo = int(self.algo)
        p.max_radius_kind = 0 if self.max_radius[0] == "absolute" else 1
        p.max_radius_value = self.max_radius[1]
        if self.cell_delta is not None:
            p.has_cell_delta, p.cell_delta = 1, self.cell_delta
        p.color_mode = int(self.color_mode)
        if self.size is not None:
            p.has_size, p.size_w = 1, self.size[0]
            if self.size[1] is not None:
                p.has_size_h, p.size_h = 1, self.size[1]
        p.seed = self.seed
        return p

    def build(self) -> "Params":
        """ParamsBuilder::build (src/params.rs:141-180): validates; raises ParamsError."""
        d = FghDerived()
        rc = _lib.load().fgh_derive(C.byref(self._c()), 1, 1, C.byref(d), None, None)
        if rc == -101:
            raise ParamsError(_lib.load().fgh_last_error().decode())
        if rc != 0:
            raise RenderError(_lib.load().fgh_last_error().decode())
        return Params(dataclasses.replace(self, n_samples=max(1, self.n_samples)))


@dataclasses.dataclass
class Params:
    """Validated Params (src/params.rs:45-68)."""
    builder: ParamsBuilder

    def __getattr__(self, name):
        return getattr(self.builder, name)


@dataclasses.dataclass
class Derived:
    """Derived (src/model.rs:167-179)."""
    input_width: int
    input_height: int
    output_width: int
    output_height: int
    inv_e_pi_r2: float
    rm: float
    delta: float
    offsets: np.ndarray        # [N,2] f32, output pixels
    offsets_input: np.ndarray  # [N,2] f32, = offsets / zoom
    algorithm: Algo
    block: FgParams            # the engine parameter block (build_uniforms, src/wgpu/mod.rs:661-692)


def derive_common(params: Params, input_size) -> Derived:
    w, h = input_size
    n = max(1, params.n_samples)
    off = np.zeros((n, 2), np.float32)
    off_in = np.zeros((n, 2), np.float32)
    d = FghDerived()
    lib = _lib.load()
    rc = lib.fgh_derive(C.byref(params.builder._c()), w, h, C.byref(d), C.c_void_p(off.ctypes.data), C.c_void_p(off_in.ctypes.data))
    if rc == -101:
        raise ParamsError(lib.fgh_last_error().decode())
    if rc != 0:
        raise RenderError(lib.fgh_last_error().decode())
    blk = FgParams()
    C.memmove(C.byref(blk), C.byref(d.block), C.sizeof(FgParams))
    return Derived(d.input_width, d.input_height, d.output_width, d.output_height, d.inv_e_pi_r2, d.rm, d.delta,
                   off, off_in, Algo(d.algorithm), blk)


def choose_algorithm(params: Params, derived: Derived) -> Algo:
    return derived.algorithm


def lambda_plane(plane: np.ndarray, inv_e_pi_r2: float) -> np.ndarray:
    """lambda_plane(normalize_plane(plane).0, inv_e_pi_r2) (src/lib.rs:148-149)."""
    src = np.ascontiguousarray(plane, np.float32)
    out = np.empty_like(src)
    lib = _lib.load()
    if lib.fgh_lambda_from_plane(C.c_void_p(src.ctypes.data), src.shape[1], src.shape[0], inv_e_pi_r2, C.c_void_p(out.ctypes.data)):
        raise RenderError(lib.fgh_last_error().decode())
    return out


def context(device: int = 0) -> Context:
    """wgpu::context() (src/wgpu/mod.rs:84-86): raises GpuError when no device is usable."""
    return Context(device)


def render_pixelwise_gpu(ctx: Context, lam: np.ndarray, params: Params, derived: Derived, rows=None) -> np.ndarray:
    blk = _band(derived.block, rows)
    return ctx.render_pixelwise(blk, lam, derived.offsets_input)


def render_grainwise_gpu(ctx: Context, lam: np.ndarray, params: Params, derived: Derived, rows=None) -> np.ndarray:
    blk = _band(derived.block, rows)
    return ctx.render_grainwise(blk, lam, derived.offsets)


def _band(block: FgParams, rows) -> FgParams:
    blk = FgParams()
    C.memmove(C.byref(blk), C.byref(block), C.sizeof(FgParams))
    if rows is not None:
        blk.row_begin, blk.row_end = rows
    return blk


def render_with_input_image(image_u8: np.ndarray, params: Params, fused: bool = False, device: int = 0, cancel=None):
    """render_with_input_image (src/lib.rs:78-84) on a decoded 8-bit RGB image [H,W,3] -> (u8 [H',W',3], Derived-lite)."""
    img = np.ascontiguousarray(image_u8, np.uint8)
    h, w, _ = img.shape
    d = derive_common(params, (w, h))
    out = np.zeros((d.output_height, d.output_width, 3), np.uint8)
    info = FghDerived()
    lib = _lib.load()
    rc = lib.fgh_render_with_input_image(C.byref(params.builder._c()), C.c_void_p(img.ctypes.data), w, h, 1 if fused else 0,
                                         device, C.byref(cancel) if cancel is not None else None,
                                         C.c_void_p(out.ctypes.data), out.size, C.byref(info))
    if rc == -103:
        raise Cancelled()
    if rc == -102:
        raise GpuError(-6, lib.fgh_last_error().decode())
    if rc == -101:
        raise ParamsError(lib.fgh_last_error().decode())
    if rc != 0:
        raise RenderError(lib.fgh_last_error().decode())
    return out, d


def _raise(rc: int):
    lib = _lib.load()
    if rc == -103:
        raise Cancelled()
    if rc == -102:
        raise GpuError(-6, lib.fgh_last_error().decode())
    if rc == -101:
        raise ParamsError(lib.fgh_last_error().decode())
    if rc != 0:
        raise RenderError(lib.fgh_last_error().decode())


def load_image(path: str) -> np.ndarray:
    """image::open + to_rgb (src/color.rs:26-29): PNG / binary PNM -> u8 [H,W,3]."""
    lib = _lib.load()
    buf, w, h = C.c_void_p(), C.c_uint64(), C.c_uint64()
    _raise(lib.fgh_load_image(os.fsencode(path), C.byref(buf), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), (h.value, w.value, 3)).copy()
    finally:
        lib.fgh_free(buf)


def save_image(path: str, image_u8: np.ndarray, fmt: str | None = None):
    """save_with_format (src/lib.rs:65-68); the format comes from `fmt`, else the extension, else PNG."""
    img = np.ascontiguousarray(image_u8, np.uint8)
    h, w, _ = img.shape
    _raise(_lib.load().fgh_save_image(os.fsencode(path), C.c_void_p(img.ctypes.data), w, h, fmt.encode() if fmt else None))


def render(params: Params, input_path: str, output_path: str, output_format: str | None = None, roi=None,
           fused: bool = False, device: int = 0, cancel=None):
    """render(params) (src/lib.rs:57-71): image file in, image file out, on the device."""
    info = FghDerived()
    roi4 = (C.c_uint32 * 4)(*roi) if roi is not None else None
    _raise(_lib.load().fgh_render_file(C.byref(params.builder._c()), os.fsencode(input_path), os.fsencode(output_path),
                                       output_format.encode() if output_format else None, roi4, 1 if fused else 0, device,
                                       C.byref(cancel) if cancel is not None else None, C.byref(info)))
    return info
