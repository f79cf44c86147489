#!/usr/bin/env python
"""bench.py -- headline benchmark of the film-grain hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our engine)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path, timed on host cores)

Metric (BASELINE.json): Mpixel*samples/s = W_out*H_out*N / seconds / 1e6 (an RGB pixel counts once)
on BASELINE.json configs[1]: pixel-wise RGB 3840x2160 synthetic noise, radius 0.1, 256 iterations.
A "step" is one full render of the image (3 colour planes).

  value : inputs (lambda planes, offsets) resident in HBM; device time (CUDA events on the engine's
          stream) from the first kernel to the finished image resident on GPU 0; max over ranks.
          N > 1: every rank's kernels store their band rows straight into GPU 0's image over NVLink
          (peer-mapped buffer, one device-side barrier ends the step; film_grain_b200/dist.py
          PeerImage), or -- where peer mapping cannot be set up, or with --bands gather -- the bands
          are gathered with NCCL.
  e2e   : the same render through the reference-facing C-ABI call with HOST buffers
          (fg_render_planes: f32 lambda planes in, f32 planes out; copies inside the timed region).
  roofline : ALU-issue roofline (the path is integer/FP32-ALU bound, not HBM- or tensor-bound):
          achieved = algorithmic lane-ops (SURVEY.md 8(d) op model with oracle-counted n_cell /
          n_test) / strip-kernel time; peak = FFMA issue rate measured in this run.
  cpu_baseline : the CPU oracle (a port of the reference's CPU path) timed on this box's host
          cores on a bounded band of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (w, h, image, color planes, ParamsBuilder kwargs)   -- BASELINE.json configs
    "c1": dict(w=512, h=512, image="gradient", planes=1, radius=0.1, n=64, zoom=1.0, algo="pixel",
               desc="pixel-wise luma 512x512 gradient r=0.1 N=64"),
    "c2": dict(w=3840, h=2160, image="noise", planes=3, radius=0.1, n=256, zoom=1.0, algo="pixel",
               desc="pixel-wise RGB 3840x2160 noise r=0.1 N=256"),
    "c3": dict(w=4096, h=4096, image="noise", planes=1, radius=0.5, n=128, zoom=1.0, algo="grain",
               desc="grain-wise luma 4096x4096 noise r=0.5 N=128"),
    "c4": dict(w=2048, h=2048, image="noise", planes=3, radius=0.05, n=64, zoom=4.0, algo="pixel",
               desc="pixel-wise RGB 2048x2048 zoom 4 (8192x8192 out) r=0.05 N=64"),
    "c5g": dict(w=1024, h=1024, image="noise", planes=1, radius=0.12, n=1024, zoom=1.0, algo="grain",
                desc="grain-wise luma 1024x1024 r=0.12 N=1024 (high-sample sweep point, grain-wise forced)"),
    "c5": dict(w=1024, h=1024, image="noise", planes=1, radius=0.12, n=1024, zoom=1.0, algo="pixel",
               desc="pixel-wise luma 1024x1024 r=0.12 N=1024 (high-sample sweep point)"),
}


def synth_image(kind: str, w: int, h: int) -> np.ndarray:
    """Deterministic 8-bit inputs; same generators as tests/helpers.py."""
    if kind == "gradient":
        row = np.rint(np.tile(np.linspace(0, 1, w), (h, 1)) * 255).astype(np.uint8)
        return np.repeat(row[:, :, None], 3, axis=2)
    return np.random.default_rng(20240611).integers(0, 256, (h, w, 3), dtype=np.uint8)


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md).  NVML in a thread
    of this process, one sample every 10 ms (a `nvidia-smi -lms` child can take longer to print its first row than
    the whole timed region lasts); `nvidia-smi` is the fallback when the NVML binding is missing."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index = index
        self.rows: list[tuple] = []      # (sm MHz, max MHz, reason names)
        self.proc = None
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def _nvml_loop(self, nv, h):
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, mx, tuple(n for b, n in self.REASONS if bits & b)))
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:  # NVML enumerates every GPU of the box, CUDA only the visible ones
                ids = [t.strip() for t in vis.split(",") if t.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    idx = int(ids[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.source = "nvml"
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self._thread.start()
            return
        except Exception:
            self._thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                try:
                    self.rows.append((float(parts[0]), float(parts[1]),
                                      tuple(n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active"))))
                except ValueError:
                    continue

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        rows = list(self.rows)
        sm = [r[0] for r in rows]
        mx = [r[1] for r in rows if r[1] is not None]
        reasons = set()
        for r in rows:
            reasons.update(r[2])
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def oracle_sample(wl: dict, img: np.ndarray, seconds: float, nthreads: int = 0):
    """Time the CPU oracle (port of src/pixelwise.rs / src/grainwise.rs) on a bounded sample of the
    workload.  Returns rate in plane-sample-evals/s plus the work counters of the sample."""
    from oracle import oracle as O
    algo = O.ALGO_PIXEL if wl["algo"] == "pixel" else O.ALGO_GRAIN
    p = O.make_params(radius=wl["radius"], n_samples=wl["n"], zoom=wl["zoom"], algo=algo)
    if wl["algo"] == "pixel":
        d, off, off_in = O.derive_common(p, wl["w"], wl["h"])
        plane = (img[:, :, 0].astype(np.float32) / np.float32(255.0)).astype(np.float32)
        lam = O.lambda_plane(O.normalize_plane(plane), d.inv_e_pi_r2)
        oh = d.output_height
        y0 = oh // 2
        c = O.Counters()
        t = time.perf_counter()
        O.render_pixelwise(lam, p, d, off_in, y0, y0 + 1, nthreads, c)  # calibration row
        dt = max(time.perf_counter() - t, 1e-4)
        threads = O.lib().fgo_max_threads() if nthreads <= 0 else nthreads
        # one row ran on one thread: all threads do ~threads rows in the same time
        rows = int(max(threads, min(oh - y0, threads * seconds / dt)))
        c = O.Counters()
        t = time.perf_counter()
        ref = O.render_pixelwise(lam, p, d, off_in, y0, y0 + rows, nthreads, c)
        dt = time.perf_counter() - t
        evals = c.sample_evals
        return dict(rate=evals / dt, seconds=dt, threads=threads, rows=rows, evals=evals,
                    ref_rows=(y0, y0 + rows, ref[y0:y0 + rows].copy()),  # plane 0, for the bitwise check against the GPU image
                    n_cell=c.cell_visits / max(1, evals), n_test=c.grain_tests / max(1, evals),
                    sample=f"{rows} output rows x {d.output_width} px x {wl['n']} samples of plane 0, {threads} threads")
    # grain-wise: a centred square crop sized for the budget
    threads = O.lib().fgo_max_threads() if nthreads <= 0 else nthreads
    side = 256
    while True:
        crop = np.ascontiguousarray(img[:side, :side, 0])
        d, off, off_in = O.derive_common(p, side, side)
        plane = (crop.astype(np.float32) / np.float32(255.0)).astype(np.float32)
        lam = O.lambda_plane(O.normalize_plane(plane), d.inv_e_pi_r2)
        c = O.Counters()
        t = time.perf_counter()
        ref = O.render_grainwise(lam, p, d, off, nthreads, c)
        dt = time.perf_counter() - t
        if dt > seconds / 4 or side * 2 > min(wl["w"], wl["h"]):
            break
        side *= 2
    evals = d.output_width * d.output_height * wl["n"]
    return dict(rate=evals / dt, seconds=dt, threads=threads, rows=side, evals=evals, n_cell=0.0, n_test=0.0,
                ref_crop=(side, ref),  # the crop's own render: equals the full render away from the crop's right/bottom edge
                sample=f"{side}x{side} crop x {wl['n']} samples, {threads} threads")


def algorithmic_ops(wl: dict, d_block, off_in: np.ndarray, n_cell: float, n_test: float) -> dict:
    """ops = S*(A_setup + n_cell*A_cell + n_test*A_test) + G*A_gen (lane-ops)."""
    A_SETUP, A_CELL, A_TEST, A_GEN = 18, 3, 6, 200
    S = float(wl["planes"]) * d_block.out_w * d_block.out_h * wl["n"]
    inv_zoom = 1.0 / wl["zoom"]
    span_x = (d_block.out_w * inv_zoom + float(off_in[:, 0].max() - off_in[:, 0].min()) + 2 * d_block.rm) / d_block.delta
    span_y = (d_block.out_h * inv_zoom + float(off_in[:, 1].max() - off_in[:, 1].min()) + 2 * d_block.rm) / d_block.delta
    G = float(wl["planes"]) * span_x * span_y
    ops_eval = S * (A_SETUP + n_cell * A_CELL + n_test * A_TEST)  # evaluation: the strip kernel
    ops_gen = G * A_GEN                                            # generation: bitmap + cell-table kernels
    ops = ops_eval + ops_gen
    return dict(ops=ops, ops_eval=ops_eval, ops_gen=ops_gen, S=S, G=G, ops_per_eval=ops / S)


def host_threads() -> int:
    """All the host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would silently turn the CPU arm into a single-thread run: the thread count is passed explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be
    built in this image (no cargo/rustc), so this is the oracle port (kind 'port'), all host threads."""
    if rank != 0:
        return
    img = synth_image(wl["image"], wl["w"], wl["h"])
    per_step = args.cpu_seconds if args.cpu_seconds else max(2.0, min(12.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        oracle_sample(wl, img, per_step / 4, host_threads())
    rates, last = [], None
    for _ in range(args.steps):
        last = oracle_sample(wl, img, per_step, host_threads())
        rates.append(last["rate"])
    rate = float(np.mean(rates))
    planes = wl["planes"]
    value = rate / planes / 1e6
    S_full = planes * (wl["w"] * wl["zoom"]) * (wl["h"] * wl["zoom"]) * wl["n"]
    line = {
        "impl": "reference", "metric": "Mpixel*samples/s", "value": value, "unit": "Mpixel*samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": S_full / rate * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": last["sample"],
                   "note": "rate measured on the bounded sample; ms_per_step extrapolated to the full workload"},
        "cpu_baseline": {"value": value, "unit": "Mpixel*samples/s", "cores": last["threads"], "kind": "port",
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": "Mpixel*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line, default=float), flush=True)


class Prepared:
    """Host-side derivation of one workload, exactly as the reference host does it (ParamsBuilder -> Derived)."""

    def __init__(self, name: str):
        import film_grain_b200 as fg
        from film_grain_b200 import host as H
        self.name = name
        wl = self.wl = WORKLOADS[name]
        self.img = synth_image(wl["image"], wl["w"], wl["h"])
        self.params = H.ParamsBuilder(radius_mean=wl["radius"], n_samples=wl["n"], zoom=wl["zoom"],
                                      algo=H.Algo.Pixel if wl["algo"] == "pixel" else H.Algo.Grain,
                                      color_mode=H.ColorMode.Rgb if wl["planes"] == 3 else H.ColorMode.Luma).build()
        d = self.d = H.derive_common(self.params, (wl["w"], wl["h"]))
        self.planes = wl["planes"]
        self.lam_host = [H.lambda_plane((self.img[:, :, c].astype(np.float32) / np.float32(255.0)).astype(np.float32), d.inv_e_pi_r2)
                         for c in range(self.planes)]
        self.offsets = d.offsets_input if wl["algo"] == "pixel" else d.offsets
        self.algo = fg.FG_ALGO_PIXEL if wl["algo"] == "pixel" else fg.FG_ALGO_GRAIN
        self.out_w, self.out_h = d.output_width, d.output_height

    def block(self, rows=None):
        from film_grain_b200 import host as H
        return H._band(self.d.block, rows)

    def mpx_samples(self, ms: float) -> float:
        return self.out_w * self.out_h * self.wl["n"] / (ms * 1e-3) / 1e6


def check_against_oracle(prep: Prepared, sm: dict, image0) -> dict:
    """Bitwise comparison of plane 0 of a finished GPU image (numpy [out_h, out_w]) with what the CPU oracle
    rendered for the cpu_baseline leg (rows of the same workload / a top-left crop for grain-wise)."""
    if "ref_rows" in sm:
        a, b, ref = sm["ref_rows"]
        ok = bool(np.array_equal(ref, image0[a:b]))
        return {"against": "oracle", "rows": [int(a), int(b)], "plane": 0, "pixels": int(ref.size),
                "result": "bitwise-equal" if ok else "MISMATCH"}
    side, ref = sm["ref_crop"]
    m = side - 16  # grains outside the crop reach its last rows / columns only
    ok = bool(np.array_equal(ref[:m, :m], image0[:m, :m]))
    return {"against": "oracle", "crop": [0, 0, int(m), int(m)], "plane": 0, "pixels": int(m * m),
            "result": "bitwise-equal" if ok else "MISMATCH"}


def quick_workload(ctx, name: str, dev, stream, flush, cpu_seconds: float, steps: int = 3, warmup: int = 3) -> dict:
    """The other BASELINE configs, driver-visible: device-resident value with its own clock record, a short CPU
    baseline and a bitwise check of the GPU image against the oracle's sample (rank 0, one GPU)."""
    import torch
    prep = Prepared(name)
    blk = prep.block()
    d_lam = torch.from_numpy(np.stack(prep.lam_host)).to(dev)
    d_off = torch.from_numpy(np.ascontiguousarray(prep.offsets)).to(dev)
    d_out = torch.zeros((prep.planes, prep.out_h, prep.out_w), dtype=torch.float32, device=dev)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            ctx.render_planes_device(blk, prep.algo, prep.planes, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
        stream.synchronize()
        sampler = ClockSampler(dev.index or 0)
        sampler.start()
        total, strip, table = 0.0, 0.0, 0.0
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.render_planes_device(blk, prep.algo, prep.planes, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
            b.record(stream)
            stream.synchronize()
            total += a.elapsed_time(b)
            st = ctx.stats()
            strip += float(st.strip_ms)
            table += float(st.table_ms)
        # short renders: keep the sampler alive for a few of its periods so that it sees the GPU under load
        t_end = time.perf_counter() + 0.6
        while time.perf_counter() < t_end:
            ctx.render_planes_device(blk, prep.algo, prep.planes, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
            stream.synchronize()
        clocks = sampler.stop()
    st = ctx.stats()
    ms = total / steps
    res = {"workload": prep.wl["desc"], "value": prep.mpx_samples(ms), "unit": "Mpixel*samples/s", "ms_per_step": ms,
           "steps": steps, "warmup": warmup, "clocks": clocks, "gpu_launches_per_step": int(st.launches),
           "tiles": int(st.tiles_total), "tiles_fallback": int(st.tiles_fallback),
           "kernel_ms": {"evaluate_or_rasterise": strip / steps, "generate": table / steps}}
    sm = oracle_sample(prep.wl, prep.img, cpu_seconds, host_threads())
    res["cpu_baseline"] = {"value": sm["rate"] / prep.planes / 1e6, "unit": "Mpixel*samples/s", "cores": sm["threads"],
                           "kind": "port", "sample": sm["sample"], "seconds": sm["seconds"]}
    res["image_check"] = check_against_oracle(prep, sm, d_out[0].cpu().numpy())
    del d_lam, d_off, d_out
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="budget of the CPU sample (default: 12 s for cpu_baseline; ~90 s over all steps of --impl reference)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads block (C1/C3/C4/C5 beside the headline)")
    ap.add_argument("--bands", default="auto", choices=["auto", "peer", "gather"],
                    help="N > 1: how the finished bands reach GPU 0 (peer = the kernels store straight into GPU 0's image "
                         "over NVLink; gather = NCCL gather; auto = peer when it can be set up)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist

    import film_grain_b200 as fg
    from film_grain_b200.dist import PeerImage, band_rows, gather_bands

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    # stdout carries exactly one JSON line: whatever native libraries print while the job runs (NCCL's version
    # banner, ...) goes to stderr; the descriptor is restored right before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side waits that must not put a spinning kernel on a GPU
    dev = torch.device("cuda", local_rank)

    prep = Prepared(args.workload)
    img, d, planes, lam_host, offsets, algo = prep.img, prep.d, prep.planes, prep.lam_host, prep.offsets, prep.algo
    out_w, out_h = prep.out_w, prep.out_h
    rb, re = band_rows(out_h, rank, world)
    blk = prep.block((rb, re) if world > 1 else None)

    ctx = fg.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d_lam = torch.from_numpy(np.stack(lam_host)).to(dev)
    d_off = torch.from_numpy(np.ascontiguousarray(offsets)).to(dev)
    d_out = torch.zeros((planes, out_h, out_w), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # N > 1: GPU 0's image mapped into every rank (NVLink peer stores), else the NCCL gather below
    peer = None
    if world > 1 and args.bands != "gather":
        peer = PeerImage.create((planes, out_h, out_w), torch.float32, dev, rank, world)
        if peer is None and args.bands == "peer":
            raise SystemExit("--bands peer: the peer image could not be set up on every rank")
    torch.cuda.synchronize()

    def step_device():
        """One step; returns the finished image on rank 0 as [planes, out_h, out_w] (None elsewhere)."""
        if peer is not None:
            # every rank's kernels write their band rows straight into GPU 0's image; one device-side
            # barrier on the engine's stream ends the step
            ctx.render_planes_device(blk, algo, planes, d_lam.data_ptr(), d_off.data_ptr(), peer.target.data_ptr(), sync=False)
            peer.finish()
            return peer.local if rank == 0 else None
        ctx.render_planes_device(blk, algo, planes, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=False)
        if world > 1:  # finished image resident on GPU 0: gather the row bands (NCCL over NVLink)
            band = d_out[:, rb:re, :].permute(1, 0, 2).contiguous()
            full = gather_bands(band, out_h, rank, world)
            return full.permute(1, 0, 2) if full is not None else None
        return d_out

    launches = 0
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        stream.synchronize()
        st = ctx.stats()
        launches_per_step = int(st.launches)
        tiles_total, tiles_fb = int(st.tiles_total), int(st.tiles_fallback)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        strip_ms, table_ms = [], []
        for a, b in evs:
            flush.zero_()  # L2 flush between timed iterations (outside the timed events)
            a.record(stream)
            step_device()
            b.record(stream)
            stream.synchronize()
            st_i = ctx.stats()
            strip_ms.append(float(st_i.strip_ms))
            table_ms.append(float(st_i.table_ms))
            launches += launches_per_step
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if rank == 0 else None
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
    tt = torch.tensor([total_ms, float(np.sum(strip_ms)), float(np.sum(table_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, strip_total_ms, table_total_ms = float(tt[0]), float(tt[1]), float(tt[2])
    strip_launches = max(1, int(ctx.stats().strip_launches)) if algo == fg.FG_ALGO_PIXEL else 1
    ms_per_step = total_ms / args.steps
    value = prep.mpx_samples(ms_per_step)

    # ---- image check, outside the timed region: the image the timed steps produce, bit for bit ----
    # N > 1: the image assembled from the ranks' bands against a single-GPU render of the whole image, computed by
    # rank 0 in this process.  N = 1: against the CPU oracle's rows (below, with cpu_baseline) and the host-buffer call.
    image_check = None
    with torch.cuda.stream(stream):
        assembled = step_device()
        stream.synchronize()
        if world > 1:
            dist.barrier()
            if rank == 0:
                single = torch.empty((planes, out_h, out_w), dtype=torch.float32, device=dev)
                ctx.render_planes_device(prep.block(None), algo, planes, d_lam.data_ptr(), d_off.data_ptr(), single.data_ptr(), sync=True)
                same = bool(torch.equal(assembled, single))
                image_check = {"against": "single-GPU render of the whole image in the same process (rank 0)",
                               "pixels": int(single.numel()), "result": "bitwise-equal" if same else "MISMATCH"}
                del single
            dist.barrier()
        device_image0 = assembled[0].cpu().numpy() if rank == 0 else None

    # ---- e2e: the reference-facing C-ABI call with HOST buffers (copies inside the timed region) ----
    e2e = None
    if world == 1:
        # host buffers of the call live in pinned (page-locked) memory, as the contract asks
        pin_lam = torch.from_numpy(np.stack(lam_host)).pin_memory()
        pin_res = torch.empty((planes, out_h, out_w), dtype=torch.float32).pin_memory()
        lam_pinned = [pin_lam[c].numpy() for c in range(planes)]
        host_outs = [pin_res[c].numpy() for c in range(planes)]
        for _ in range(2):
            ctx.render_planes(blk, algo, lam_pinned, offsets, host_outs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.render_planes(blk, algo, lam_pinned, offsets, host_outs)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        s = ctx.stats()
        e2e = {"value": prep.mpx_samples(e2e_ms), "unit": "Mpixel*samples/s",
               "h2d_bytes_per_step": int(s.h2d_bytes), "d2h_bytes_per_step": int(s.d2h_bytes),
               "ms_per_step": e2e_ms, "h2d_ms": float(s.h2d_ms), "kernel_ms": float(s.kernel_ms), "d2h_ms": float(s.d2h_ms),
               "equals_device_image": bool(np.array_equal(host_outs[0], device_image0)),
               "call": "fg_render_planes (pinned host f32 lambda planes in, f32 planes out; the output planes are one page-locked "
                       "block, which the kernels write in place over PCIe -- d2h_ms is the staged copy, 0 when in place)"}
        # The seam's own call pattern (src/color.rs:56-60, src/lib.rs:154-158): one fg_render_pixelwise / _grainwise per
        # colour plane, sequentially, on PAGEABLE buffers (a Rust Vec<f32>): staged copies both ways, per-plane set-up.
        seam_out = [np.empty((out_h, out_w), np.float32) for _ in range(planes)]
        seam_lam = [np.array(l, copy=True) for l in lam_host]
        one = ctx.render_pixelwise if algo == fg.FG_ALGO_PIXEL else ctx.render_grainwise

        def step_seam():
            for c in range(planes):
                one(blk, seam_lam[c], offsets, out=seam_out[c])
        step_seam()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_seam()
        seam_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e["seam"] = {"value": prep.mpx_samples(seam_ms), "unit": "Mpixel*samples/s", "ms_per_step": seam_ms,
                       "h2d_bytes_per_step": int(sum(l.nbytes for l in seam_lam) + planes * offsets.nbytes),
                       "d2h_bytes_per_step": int(sum(o.nbytes for o in seam_out)),
                       "equals_device_image": bool(np.array_equal(seam_out[0], device_image0)),
                       "call": f"{planes} x fg_render_{'pixelwise' if algo == fg.FG_ALGO_PIXEL else 'grainwise'} in sequence on pageable "
                               "numpy buffers (what render_from_workspace_impl does per plane, src/lib.rs:146-166)"}
        del seam_out, seam_lam
        # the fused u8 entry point: 8-bit RGB over PCIe instead of f32 planes
        if planes == 3 and wl["algo"] == "pixel":
            pin_img = torch.from_numpy(np.ascontiguousarray(img)).pin_memory()
            pin_out8 = torch.empty((out_h, out_w, 3), dtype=torch.uint8).pin_memory()
            img8, out8 = pin_img.numpy(), pin_out8.numpy()
            for _ in range(2):
                ctx.render_rgb8(blk, algo, fg.FG_COLOR_RGB, img8, offsets, out8)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.render_rgb8(blk, algo, fg.FG_COLOR_RGB, img8, offsets, out8)
            f_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            s8 = ctx.stats()
            e2e["fused_rgb8"] = {"value": prep.mpx_samples(f_ms), "ms_per_step": f_ms,
                                 "h2d_bytes_per_step": int(s8.h2d_bytes), "d2h_bytes_per_step": int(s8.d2h_bytes),
                                 "call": "fg_render_rgb8 (host u8 RGB in/out, colour fused on device)"}
            del pin_img, pin_out8, img8, out8
        del pin_lam, pin_res, lam_pinned, host_outs
    else:
        from film_grain_b200.dist import SharedHostImage
        shared = SharedHostImage.create((planes, out_h, out_w), rank, world, dev) if args.bands != "gather" else None
        if shared is not None:
            # The reference-facing host call on every rank: fg_render_planes with the rank's row band, pinned
            # lambda planes in (the engine uploads only the input rows the band reads) and the planes of ONE
            # page-locked host image shared by all ranks out (the kernels store the band rows straight into
            # it, each GPU over its own PCIe link); a barrier ends the step with the whole image in host memory.
            pin_lam = torch.from_numpy(np.stack(lam_host)).pin_memory()
            lam_pinned = [pin_lam[c].numpy() for c in range(planes)]
            host_outs = [shared.array[c] for c in range(planes)]

            def step_e2e():
                ctx.render_planes(blk, algo, lam_pinned, offsets, host_outs)
                dist.barrier()
            for _ in range(2):
                step_e2e()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            s = ctx.stats()
            tb = torch.tensor([e2e_ms, float(s.h2d_bytes), float(s.d2h_bytes)], dtype=torch.float64, device=dev)
            tmax = tb.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
            e2e_ms = float(tmax[0])
            e2e = {"value": prep.mpx_samples(e2e_ms), "unit": "Mpixel*samples/s",
                   "h2d_bytes_per_step": int(tb[1]), "d2h_bytes_per_step": int(tb[2]), "ms_per_step": e2e_ms,
                   "equals_device_image": bool(np.array_equal(shared.array[0], device_image0)) if rank == 0 else None,
                   "call": "per rank: fg_render_planes on its row band (pinned host lambda in, band-restricted upload; output = one "
                           "page-locked host image shared by all ranks, written in place by the kernels) + barrier; bytes summed over ranks"}
            del host_outs, lam_pinned, pin_lam
            shared.close()
        else:
            # each rank uploads only the lambda rows its band can see: band / zoom +- (max |offset| + rm) plus slack
            reach = float(np.abs(d.offsets_input[:, 1]).max()) + float(d.rm) + 2.0
            in_r0 = max(0, int(np.floor(rb / wl["zoom"] - reach)))
            in_r1 = min(wl["h"], int(np.ceil(re / wl["zoom"] + reach)) + 1)
            pin_in = torch.from_numpy(np.stack(lam_host)[:, in_r0:in_r1, :].copy()).pin_memory()
            pin_out = torch.empty((planes, out_h, out_w), dtype=torch.float32).pin_memory() if rank == 0 else None
            with torch.cuda.stream(stream):
                def step_e2e():
                    if peer is not None:
                        peer.finish()  # GPU 0 has read the previous image out before anyone overwrites it
                    d_lam[:, in_r0:in_r1, :].copy_(pin_in, non_blocking=True)
                    full = step_device()
                    if rank == 0:
                        pin_out.copy_(full, non_blocking=True)
                step_e2e()
                stream.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    step_e2e()
                    stream.synchronize()
                dist.barrier()
                e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            e2e_ms = float(t2[0])
            e2e = {"value": prep.mpx_samples(e2e_ms), "unit": "Mpixel*samples/s",
                   "h2d_bytes_per_step": int(pin_in.numel() * 4), "d2h_bytes_per_step": int(planes * out_w * out_h * 4),
                   "ms_per_step": e2e_ms,
                   "call": "per rank: pinned H2D of the lambda rows its band sees + band render + "
                           + ("peer stores into GPU 0's image + barrier" if peer is not None else "NCCL band gather") + "; rank 0: D2H of the image"}
            del pin_in, pin_out

    # ---- N > 1: the same job from ONE process through the C ABI (fg_context_create_multi): what a single-process caller such
    # as the reference's render() gets from the box.  Rank 0 drives all N GPUs (one host thread + stream per device inside the
    # library); the other ranks wait at a barrier.  Pinned host planes in and out, copies inside the timed region.
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)  # gloo: the waiting ranks leave their GPUs idle (an NCCL barrier spins ON the GPU rank 0 is about to use from this process, and the two processes would time-slice it)
        if rank == 0:
            try:
                mctx = fg.Context(devices=list(range(world)))
                pin_lam = torch.from_numpy(np.stack(lam_host)).pin_memory()
                pin_res = torch.empty((planes, out_h, out_w), dtype=torch.float32).pin_memory()
                lam_pinned = [pin_lam[c].numpy() for c in range(planes)]
                host_outs = [pin_res[c].numpy() for c in range(planes)]
                full_blk = prep.block(None)
                for _ in range(2):
                    mctx.render_planes(full_blk, algo, lam_pinned, offsets, host_outs)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    mctx.render_planes(full_blk, algo, lam_pinned, offsets, host_outs)
                sp_ms = (time.perf_counter() - t0) * 1e3 / args.steps
                s = mctx.stats()
                e2e["single_process"] = {
                    "value": prep.mpx_samples(sp_ms), "unit": "Mpixel*samples/s", "ms_per_step": sp_ms, "devices": world,
                    "h2d_bytes_per_step": int(s.h2d_bytes), "d2h_bytes_per_step": int(s.d2h_bytes),
                    "equals_device_image": bool(np.array_equal(host_outs[0], device_image0)),
                    "call": "one process, fg_context_create_multi over all N devices + fg_render_planes (pinned host planes in / out; "
                            "row bands, one host thread and stream per device inside the library)"}
                del pin_lam, pin_res, lam_pinned, host_outs
                mctx.close()
            except Exception as e:  # a side measurement never costs the headline line
                e2e["single_process"] = {"error": repr(e)[:300]}
        dist.barrier(group=cpu_group)

    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        issue = ctx.measure_issue_peak()
        cpu, roof = None, None
        n_cell, n_test = 6.6, 1.6
        if world == 1 and not args.no_cpu:
            sm = oracle_sample(wl, img, args.cpu_seconds or 12.0, host_threads())
            n_cell, n_test = (sm["n_cell"], sm["n_test"]) if wl["algo"] == "pixel" else (n_cell, n_test)
            cpu = {"value": sm["rate"] / planes / 1e6, "unit": "Mpixel*samples/s", "cores": sm["threads"], "kind": "port",
                   "sample": sm["sample"], "seconds": sm["seconds"],
                   "note": "the port skips Poisson::new's ln / sqrt / log_gamma for lambda' < 12 (only exp(-lambda') is used on that "
                           "branch), i.e. it does a little LESS work per cell visit than the Rust: the GPU/CPU ratio is conservative"}
            image_check = check_against_oracle(prep, sm, device_image0)
        if wl["algo"] == "pixel":
            ao = algorithmic_ops(wl, d.block, d.offsets_input, n_cell, n_test)
            # dominant kernel: the evaluation kernel (k_pixelwise_tri / k_pixelwise_strip).  Its algorithmic work is the
            # evaluation term of SURVEY 8(d); the generation term belongs to the bitmap + cell-table kernels that
            # run before it and is reported beside it ("table") and in the whole-step figure ("pipeline").
            kernel_ms = (strip_total_ms / args.steps) if strip_total_ms > 0 else ms_per_step
            tab_ms = table_total_ms / args.steps
            achieved = ao["ops_eval"] / strip_launches / (kernel_ms * 1e-3) / 1e12
            traffic, issue_active, prof_src = None, None, None
            prof = os.path.join(ROOT, "profiles", "eval_kernel_static.json")
            if os.path.exists(prof):
                try:
                    pj = json.load(open(prof)).get(args.workload, {}).get(ctx.eval_kernel_name(), {})
                    traffic, issue_active, prof_src = pj.get("traffic"), pj.get("issue_active"), pj.get("source")
                except Exception:
                    traffic = None
            io_bytes = planes * (wl["w"] * wl["h"] * 16 + out_w * out_h * 4)  # thr+e planes in, f32 plane out
            peak = issue["ffma"] / 1e3 * world  # aggregate over the GPUs that shared the launch's work
            roof = {"bound": "alu", "achieved": achieved, "peak": peak, "unit": "Tlane-op/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "issue_slots_filled": issue_active,
                    "static": {"traffic": True, "issue_slots_filled": True, "source": prof_src,
                               "note": "traffic (dram__bytes_read + write) and issue_slots_filled (smsp__issue_active) are NOT measured in this "
                                       "run: they are read from the committed ncu capture of the same kernel and workload"},
                    "kernel": ctx.eval_kernel_name(), "kernel_ms": kernel_ms, "share_of_step": kernel_ms / ms_per_step,
                    "launches_per_step": strip_launches,
                    "table": {"kernels": "k_thresholds + k_first_draw_bitmap + k_row_expect + k_row_bases + k_gen_rows",
                              "ms": tab_ms, "share_of_step": tab_ms / ms_per_step,
                              "achieved": (ao["ops_gen"] / (tab_ms * 1e-3) / 1e12) if tab_ms > 0 else None,
                              "frac": (ao["ops_gen"] / (tab_ms * 1e-3) / 1e12 / peak) if tab_ms > 0 else None},
                    "pipeline": {"achieved": ao["ops"] / (ms_per_step * 1e-3) / 1e12,
                                 "frac": ao["ops"] / (ms_per_step * 1e-3) / 1e12 / peak,
                                 "note": "all SURVEY 8(d) lane-ops of the step / whole step time"},
                    "ops_model": {"A_setup": 18, "A_cell": 3, "A_test": 6, "A_gen": 200, "n_cell": n_cell, "n_test": n_test,
                                  "sample_evals": ao["S"], "cells": ao["G"], "ops_per_sample_eval": ao["ops_per_eval"],
                                  "ops_eval": ao["ops_eval"], "ops_gen": ao["ops_gen"],
                                  "counted_by": "oracle sample" if cpu else "default"},
                    "peak_source": "FFMA issue rate measured in this run (fg_measure_issue_peak); IMAD is half rate",
                    "issue_peaks_glaneops": issue,
                    "hbm": {"achieved_gbs": io_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "frac": io_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if os.path.exists(peaks_file) else "fallback"}}
        if wl["algo"] == "grain":
            # unit = one (grain, sample) disk rasterisation, 14 lane-ops of set-up + 5 per box pixel
            # (a box holds (2R)^2 pixel centres on average), plus A_gen = 200 per input pixel for the generation.
            # Grain count = the Poisson means' sum (the realised count differs by < 1e-3 relative).
            grains = float(sum(float(np.asarray(l, np.float64).sum()) for l in lam_host))
            r_out = float(min(wl["radius"], float(d.rm))) * wl["zoom"]
            pairs = grains * wl["n"]
            ops_r = pairs * (14.0 + 5.0 * (2.0 * r_out) ** 2)
            ops_g = 200.0 * planes * wl["w"] * wl["h"]
            kernel_ms = (strip_total_ms / args.steps) if strip_total_ms > 0 else ms_per_step
            tab_ms = table_total_ms / args.steps
            peak = issue["ffma"] / 1e3 * world
            achieved = ops_r / (kernel_ms * 1e-3) / 1e12
            roof = {"bound": "alu", "achieved": achieved, "peak": peak, "unit": "Tlane-op/s", "frac": achieved / peak,
                    "traffic": None, "kernel": ctx.eval_kernel_name() + " (last plane)", "kernel_ms": kernel_ms,
                    "share_of_step": kernel_ms * planes / ms_per_step,
                    "launches_per_step": planes,
                    "generation": {"kernels": "k_gw_count + scan + k_gw_fill", "ms": tab_ms,
                                   "achieved": (ops_g / planes / (tab_ms * 1e-3) / 1e12) if tab_ms > 0 else None},
                    "pipeline": {"achieved": (ops_r + ops_g) / (ms_per_step * 1e-3) / 1e12,
                                 "frac": (ops_r + ops_g) / (ms_per_step * 1e-3) / 1e12 / peak},
                    "ops_model": {"A_setup": 14, "A_pixel": 5, "A_gen": 200, "grains": grains, "pairs": pairs,
                                  "box_pixels": (2.0 * r_out) ** 2, "ops_raster": ops_r, "ops_gen": ops_g},
                    "peak_source": "FFMA issue rate measured in this run (fg_measure_issue_peak)",
                    "issue_peaks_glaneops": issue}
        others = None
        if world == 1 and args.workload == "c2" and not args.no_others and not args.no_cpu:
            # the other BASELINE configs beside the headline, each with its own clocks, CPU sample and oracle check
            others = {}
            for name in ("c1", "c3", "c4", "c5"):
                try:
                    others[name] = quick_workload(ctx, name, dev, stream, flush, cpu_seconds=3.0)
                except Exception as e:  # never lose the headline line to a side measurement
                    others[name] = {"error": repr(e)[:300]}
        line = {
            "metric": "Mpixel*samples/s", "value": value, "unit": "Mpixel*samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "seed": 5489, "sigma_px": 0.8, "algo": wl["algo"],
                       "parallelism": f"row-bands x{world}" if world > 1 else "single GPU",
                       "bands": (None if world == 1 else (f"NVLink peer stores into GPU 0's image ({peer.mode}) + device barrier"
                                                          if peer is not None else "NCCL gather to GPU 0")),
                       "l2": "flushed between timed iterations (256 MiB memset outside the events)",
                       "tiles": tiles_total, "tiles_fallback": tiles_fb},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "image_check": image_check,
            "roofline": roof, "cpu_baseline": cpu, "other_workloads": others,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line, default=float), flush=True)
    # Teardown in dependency order, then a NORMAL interpreter exit (atexit hooks and finalizers run).  torch's caching
    # allocators record an event on every stream a block was used on when the block is released -- including the
    # engine's stream -- so every tensor goes first and the caches are emptied while that stream is still alive; the
    # engine context (which owns the stream) is closed last.
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    del d_lam, d_off, d_out, flush
    peer = None
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()  # pinned-host cache (torch >= 2.5); absent: the blocks are event-free by now
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()
    del stream
    # the context is deliberately NOT destroyed: process exit releases it, and any block torch still holds can
    # record its event on a live stream
    ctx.leak()
    sys.stdout.flush()
    sys.stderr.flush()


if __name__ == "__main__":
    main()
