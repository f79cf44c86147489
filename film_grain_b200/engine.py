"""Thin Python handle on the C-ABI engine (include/fg.h): context lifetime, numpy in/out.

Plumbing only -- every render goes through libfg_b200.so's CUDA kernels; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FgParams, FgStats


class GpuError(RuntimeError):
    """RenderError::Gpu (src/lib.rs:36-37) + the return code of the failing call."""

    def __init__(self, code: int, message: str):
        super().__init__(f"gpu error: {message} [{_lib.load().fg_error_string(code).decode()}]")
        self.code = code
        self.message = message


class Cancelled(RuntimeError):
    """RenderError::Cancelled (src/lib.rs:40-41)."""


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Context:
    """GpuContext (src/wgpu/mod.rs:40-49): owns the device stream and buffer pools."""

    def __init__(self, device: int = 0, devices=None):
        """device: one CUDA ordinal; devices=[d0, d1, ...]: one context over several devices of the box (row bands,
        fg_context_create_multi; device-resident buffers live on d0)."""
        self._lib = _lib.load()
        h = C.c_void_p()
        if devices is not None:
            devices = [int(d) for d in devices]
            arr = (C.c_int * len(devices))(*devices)
            rc = self._lib.fg_context_create_multi(C.byref(h), arr, len(devices))
            device = devices[0] if devices else 0
        else:
            rc = self._lib.fg_context_create(C.byref(h), device)
        if rc != 0:
            raise GpuError(rc, "no usable CUDA device" if rc == _lib.FG_ERR_NO_DEVICE else "context creation failed")
        self._h = h
        self.device = device
        self.devices = devices if devices is not None else [device]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fg_context_destroy(self._h)
            self._h = None

    def leak(self):
        """Give the native context up without destroying it (it lives until the process exits).  For callers
        that wrapped the context's stream in another runtime's stream object: that runtime may still record
        events on the stream while it tears down, after this object is gone."""
        self._h = None

    def device_count(self) -> int:
        return int(self._lib.fg_context_device_count(self._h))

    def eval_kernel_name(self) -> str:
        return self._lib.fg_last_eval_kernel(self._h).decode()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- helpers ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc == 0:
            return
        msg = self._lib.fg_last_error(self._h).decode()
        if rc == _lib.FG_ERR_CANCELLED:
            raise Cancelled()
        raise GpuError(rc, msg)

    def stats(self) -> FgStats:
        s = FgStats()
        self._lib.fg_get_stats(self._h, C.byref(s))
        return s

    @property
    def stream(self) -> int:
        return int(self._lib.fg_context_stream(self._h))

    def synchronize(self):
        self._check(self._lib.fg_context_synchronize(self._h))

    # -- host-buffer entry points ----------------------------------------------------------
    def render_pixelwise(self, p: FgParams, lam: np.ndarray, offsets_input: np.ndarray, out: np.ndarray | None = None):
        lam = np.ascontiguousarray(lam, np.float32)
        off = np.ascontiguousarray(offsets_input, np.float32)
        assert lam.shape == (p.in_h, p.in_w) and off.shape == (p.n_samples, 2)
        if out is None:
            out = np.zeros((p.out_h, p.out_w), np.float32)
        self._check(self._lib.fg_render_pixelwise(self._h, C.byref(p), _ptr(lam), _ptr(off), _ptr(out)))
        return out

    def render_grainwise(self, p: FgParams, lam: np.ndarray, offsets: np.ndarray, out: np.ndarray | None = None):
        lam = np.ascontiguousarray(lam, np.float32)
        off = np.ascontiguousarray(offsets, np.float32)
        assert lam.shape == (p.in_h, p.in_w) and off.shape == (p.n_samples, 2)
        if out is None:
            out = np.zeros((p.out_h, p.out_w), np.float32)
        self._check(self._lib.fg_render_grainwise(self._h, C.byref(p), _ptr(lam), _ptr(off), _ptr(out)))
        return out

    def render_planes(self, p: FgParams, algo: int, lams, offsets: np.ndarray, outs=None):
        lams = [np.ascontiguousarray(a, np.float32) for a in lams]
        off = np.ascontiguousarray(offsets, np.float32)
        n = len(lams)
        if outs is None:
            outs = [np.zeros((p.out_h, p.out_w), np.float32) for _ in range(n)]
        lp = (C.c_void_p * n)(*[a.ctypes.data for a in lams])
        op = (C.c_void_p * n)(*[a.ctypes.data for a in outs])
        self._check(self._lib.fg_render_planes(self._h, C.byref(p), algo, n, lp, _ptr(off), op))
        return outs

    def render_rgb8(self, p: FgParams, algo: int, color_mode: int, rgb: np.ndarray, offsets: np.ndarray,
                    out: np.ndarray | None = None):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        off = np.ascontiguousarray(offsets, np.float32)
        assert rgb.shape == (p.in_h, p.in_w, 3)
        if out is None:
            out = np.zeros((p.out_h, p.out_w, 3), np.uint8)
        self._check(self._lib.fg_render_rgb8(self._h, C.byref(p), algo, color_mode, _ptr(rgb), _ptr(off), _ptr(out)))
        return out

    # -- device-pointer entry points (raw addresses, e.g. torch.Tensor.data_ptr()) ---------
    def render_planes_device(self, p: FgParams, algo: int, n_planes: int, d_lambda: int, d_offsets: int, d_out: int,
                             sync: bool = True):
        self._check(self._lib.fg_render_planes_device(self._h, C.byref(p), algo, n_planes, C.c_void_p(d_lambda),
                                                      C.c_void_p(d_offsets), C.c_void_p(d_out), 1 if sync else 0))

    def render_rgb8_device(self, p: FgParams, algo: int, color_mode: int, d_rgb_in: int, d_offsets: int,
                           d_rgb_out: int, sync: bool = True):
        self._check(self._lib.fg_render_rgb8_device(self._h, C.byref(p), algo, color_mode, C.c_void_p(d_rgb_in),
                                                    C.c_void_p(d_offsets), C.c_void_p(d_rgb_out), 1 if sync else 0))

    # -- debug / measurement -----------------------------------------------------------------
    def dump_cells(self, p: FgParams, stream_kind: int, ij: np.ndarray, lam: np.ndarray, cap: int = 8):
        ij = np.ascontiguousarray(ij, np.int32)
        lam = np.ascontiguousarray(lam, np.float32)
        n = ij.shape[0]
        q = np.zeros(n, np.uint32)
        g = np.zeros((n, cap, 3), np.float32)
        self._check(self._lib.fg_dump_cells(self._h, C.byref(p), stream_kind, _ptr(ij), _ptr(lam), n, cap, _ptr(q), _ptr(g)))
        return q, g

    def measure_issue_peak(self):
        out = (C.c_double * 4)()
        self._check(self._lib.fg_measure_issue_peak(self._h, out))
        return {"ffma": out[0], "imad": out[1], "imad_lop3_mix": out[2], "dfma": out[3]}

    # -- viewer-grade re-render (src/bin/viewer.rs:944-1067) ---------------------
    def set_table_cache(self, enable: bool):
        """Keep the cell table of the last whole-frame pixel-wise render; later renders that differ only in
        n_samples / offsets / zoom evaluate from it (stats().table_reused == 1)."""
        self._lib.fg_set_table_cache(self._h, 1 if enable else 0)

    def refine_planes(self, p: FgParams, algo: int, lams, offsets: np.ndarray, k_begin: int, k_end: int, outs=None,
                      cancel: C.c_int | None = None):
        """Render samples [k_begin, k_end) of the p.n_samples offsets and merge them into the running image: `outs`
        then equals a render of the first k_end samples bit for bit (fg_refine_planes)."""
        lams = [np.ascontiguousarray(a, np.float32) for a in lams]
        off = np.ascontiguousarray(offsets, np.float32)
        assert off.shape == (p.n_samples, 2)
        n = len(lams)
        if outs is None:
            outs = [np.zeros((p.out_h, p.out_w), np.float32) for _ in range(n)]
        lp = (C.c_void_p * n)(*[a.ctypes.data for a in lams])
        op = (C.c_void_p * n)(*[a.ctypes.data for a in outs])
        self._check(self._lib.fg_refine_planes(self._h, C.byref(p), algo, n, lp, _ptr(off), k_begin, k_end, op,
                                               C.byref(cancel) if cancel is not None else None))
        return outs

    def set_cancel_flag(self, flag: C.c_int | None):
        self._cancel = flag  # keep alive
        self._lib.fg_set_cancel_flag(self._h, C.byref(flag) if flag is not None else None)


def device_count() -> int:
    return _lib.load().fg_device_count()
