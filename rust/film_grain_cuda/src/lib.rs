//! film_grain_cuda -- the B200 engine behind the reference's `--device gpu` seam.
//!
//! `sys` mirrors `include/fg.h` one to one (every exported symbol, both structs, every enum value the callers use).
//! The safe layer mirrors the public surface of the reference's `pub mod wgpu` (src/wgpu/mod.rs:84-86, 336-345,
//! 473-482): `context()`, `render_pixelwise_gpu`, `render_grainwise_gpu`, plus what the engine adds beyond the seam
//! (batched planes, in-launch cancel, the viewer's table cache and progressive refinement, multi-device contexts).
//! It works on plain slices so that it does not depend on the reference crate; README.md shows the ten-line glue
//! that maps `Params` / `Derived` / `Plane` onto it inside `src/lib.rs`.

use std::ffi::{c_char, c_int, CStr};
use std::sync::atomic::AtomicI32;
use std::sync::{Arc, Mutex, OnceLock};

pub mod sys {
    use std::ffi::{c_char, c_int};

    /// struct fg_params (include/fg.h): Params (src/params.rs:45-68) + Derived (src/model.rs:167-179) reduced to
    /// what the integrators read; filled where build_uniforms fills Uniforms today (src/wgpu/mod.rs:661-692).
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct FgParams {
        pub struct_size: u32,
        pub in_w: u32,
        pub in_h: u32,
        pub out_w: u32,
        pub out_h: u32,
        pub n_samples: u32,
        pub dist_kind: u32,
        pub seeding: u32,
        pub seed: u64,
        pub zoom: f32,
        pub delta: f32,
        pub rm: f32,
        pub inv_e_pi_r2: f32,
        pub radius_mean: f32,
        pub has_log: u32,
        pub radius_log_mu: f64,
        pub radius_log_sigma: f64,
        pub row_begin: u32,
        pub row_end: u32,
        pub path: u32,
        pub reserved: u32,
    }

    /// struct fg_stats (include/fg.h)
    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct FgStats {
        pub kernel_ms: f32,
        pub h2d_ms: f32,
        pub d2h_ms: f32,
        pub launches: u32,
        pub tiles_total: u32,
        pub tiles_fallback: u32,
        pub h2d_bytes: u64,
        pub d2h_bytes: u64,
        pub strip_ms: f32,
        pub strip_launches: u32,
        pub table_ms: f32,
        pub table_reused: u32,
    }

    #[repr(C)]
    pub struct FgCtx {
        _private: [u8; 0],
    }

    pub const FG_OK: c_int = 0;
    pub const FG_ERR_INVALID: c_int = -1;
    pub const FG_ERR_OOM: c_int = -2;
    pub const FG_ERR_CUDA_STICKY: c_int = -3;
    pub const FG_ERR_NO_DEVICE: c_int = -4;
    pub const FG_ERR_CANCELLED: c_int = -5;
    pub const FG_ERR_CUDA: c_int = -6;

    pub const FG_DIST_CONST: u32 = 0;
    pub const FG_DIST_LOGNORM: u32 = 1;
    pub const FG_COLOR_LUMA: c_int = 0;
    pub const FG_COLOR_RGB: c_int = 1;
    pub const FG_ALGO_GRAIN: c_int = 1;
    pub const FG_ALGO_PIXEL: c_int = 2;
    pub const FG_SEEDING_RAND_0_8: u32 = 0;
    pub const FG_SEEDING_RAND_0_9: u32 = 1;
    pub const FG_PATH_AUTO: u32 = 0;

    extern "C" {
        pub fn fg_abi_version() -> c_int;
        pub fn fg_device_count() -> c_int;
        pub fn fg_error_string(code: c_int) -> *const c_char;
        pub fn fg_context_create(out: *mut *mut FgCtx, device: c_int) -> c_int;
        pub fn fg_context_create_multi(out: *mut *mut FgCtx, devices: *const c_int, n_devices: c_int) -> c_int;
        pub fn fg_context_device_count(ctx: *const FgCtx) -> c_int;
        pub fn fg_context_destroy(ctx: *mut FgCtx);
        pub fn fg_last_error(ctx: *const FgCtx) -> *const c_char;
        pub fn fg_last_eval_kernel(ctx: *const FgCtx) -> *const c_char;
        pub fn fg_set_cancel_flag(ctx: *mut FgCtx, flag: *const c_int);
        pub fn fg_get_stats(ctx: *const FgCtx, out: *mut FgStats);
        pub fn fg_render_pixelwise(ctx: *mut FgCtx, p: *const FgParams, lambda: *const f32, offsets_input: *const f32, out: *mut f32) -> c_int;
        pub fn fg_render_grainwise(ctx: *mut FgCtx, p: *const FgParams, lambda: *const f32, offsets: *const f32, out: *mut f32) -> c_int;
        pub fn fg_render_planes(ctx: *mut FgCtx, p: *const FgParams, algo: c_int, n_planes: c_int, lambda: *const *const f32,
                                offsets: *const f32, out: *const *mut f32) -> c_int;
        pub fn fg_render_planes_cancelable(ctx: *mut FgCtx, p: *const FgParams, algo: c_int, n_planes: c_int, lambda: *const *const f32,
                                           offsets: *const f32, out: *const *mut f32, cancel: *const c_int) -> c_int;
        pub fn fg_set_table_cache(ctx: *mut FgCtx, enable: c_int);
        pub fn fg_refine_planes(ctx: *mut FgCtx, p: *const FgParams, algo: c_int, n_planes: c_int, lambda: *const *const f32,
                                offsets: *const f32, k_begin: u32, k_end: u32, out: *const *mut f32, cancel: *const c_int) -> c_int;
        pub fn fg_render_planes_device(ctx: *mut FgCtx, p: *const FgParams, algo: c_int, n_planes: c_int, d_lambda: *const f32,
                                       d_offsets: *const f32, d_out: *mut f32, stream_sync: c_int) -> c_int;
        pub fn fg_context_stream(ctx: *const FgCtx) -> u64;
        pub fn fg_context_synchronize(ctx: *mut FgCtx) -> c_int;
        pub fn fg_render_rgb8(ctx: *mut FgCtx, p: *const FgParams, algo: c_int, color_mode: c_int, rgb_in: *const u8,
                              offsets: *const f32, rgb_out: *mut u8) -> c_int;
        pub fn fg_render_rgb8_device(ctx: *mut FgCtx, p: *const FgParams, algo: c_int, color_mode: c_int, d_rgb_in: *const u8,
                                     d_offsets: *const f32, d_rgb_out: *mut u8, stream_sync: c_int) -> c_int;
        pub fn fg_dump_cells(ctx: *mut FgCtx, p: *const FgParams, stream_kind: c_int, ij: *const i32, lambda_cell: *const f32,
                             n: usize, cap: u32, counts: *mut u32, grains: *mut f32) -> c_int;
        pub fn fg_measure_issue_peak(ctx: *mut FgCtx, out4: *mut f64) -> c_int;
    }
}

pub use sys::{FgParams, FgStats};

/// RenderError::{Gpu, Cancelled} of the reference (src/lib.rs:32-44), without depending on that crate.
#[derive(Debug)]
pub enum GpuError {
    /// RenderError::Cancelled
    Cancelled,
    /// RenderError::Gpu(String); `fatal`: the context was dropped (handle_gpu_error, src/wgpu/mod.rs:727-752)
    Gpu { message: String, fatal: bool },
}

impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            GpuError::Cancelled => write!(f, "cancelled"),
            GpuError::Gpu { message, .. } => write!(f, "gpu: {message}"),
        }
    }
}
impl std::error::Error for GpuError {}

pub struct GpuContext {
    raw: *mut sys::FgCtx,
}
unsafe impl Send for GpuContext {} // calls are serialised inside the library (one mutex per context)
unsafe impl Sync for GpuContext {}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { sys::fg_context_destroy(self.raw) }
    }
}

static GPU_CONTEXT: OnceLock<Mutex<Option<Arc<GpuContext>>>> = OnceLock::new();

fn cstr(p: *const c_char) -> String {
    if p.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }
}

/// wgpu::context() (src/wgpu/mod.rs:84-86): the process-wide context, created on first use.  `FG_B200_DEVICES`
/// ("0,1,2,3") makes it a multi-device context: every render is split into row bands, one per device.
pub fn context() -> Result<Arc<GpuContext>, GpuError> {
    let cell = GPU_CONTEXT.get_or_init(|| Mutex::new(None));
    let mut guard = cell.lock().unwrap_or_else(|e| e.into_inner());
    if let Some(ctx) = guard.as_ref() {
        return Ok(ctx.clone());
    }
    let devices: Vec<c_int> = std::env::var("FG_B200_DEVICES")
        .ok()
        .map(|s| s.split(',').filter_map(|t| t.trim().parse().ok()).collect())
        .unwrap_or_default();
    let ctx = Arc::new(if devices.len() > 1 { GpuContext::new_multi(&devices)? } else { GpuContext::new(devices.first().copied().unwrap_or(0))? });
    *guard = Some(ctx.clone());
    Ok(ctx)
}

/// invalidate_context() (src/wgpu/mod.rs:88-92)
pub fn invalidate_context() {
    if let Some(cell) = GPU_CONTEXT.get() {
        cell.lock().unwrap_or_else(|e| e.into_inner()).take();
    }
}

impl GpuContext {
    pub fn new(device: c_int) -> Result<Self, GpuError> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::fg_context_create(&mut raw, device) };
        if rc != sys::FG_OK {
            // no CPU fallback: the viewer greys out the GPU option (src/bin/viewer.rs:784-790)
            return Err(GpuError::Gpu { message: cstr(unsafe { sys::fg_error_string(rc) }), fatal: true });
        }
        Ok(Self { raw })
    }

    pub fn new_multi(devices: &[c_int]) -> Result<Self, GpuError> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::fg_context_create_multi(&mut raw, devices.as_ptr(), devices.len() as c_int) };
        if rc != sys::FG_OK {
            return Err(GpuError::Gpu { message: cstr(unsafe { sys::fg_error_string(rc) }), fatal: true });
        }
        Ok(Self { raw })
    }

    pub fn device_count(&self) -> usize {
        unsafe { sys::fg_context_device_count(self.raw) as usize }
    }

    pub fn stats(&self) -> FgStats {
        let mut s = FgStats::default();
        unsafe { sys::fg_get_stats(self.raw, &mut s) };
        s
    }

    /// Keep the cell table across renders that change only n_samples / sigma / zoom (the viewer's sliders).
    pub fn set_table_cache(&self, enable: bool) {
        unsafe { sys::fg_set_table_cache(self.raw, enable as c_int) }
    }

    fn check(&self, rc: c_int, label: &str) -> Result<(), GpuError> {
        if rc == sys::FG_OK {
            return Ok(());
        }
        if rc == sys::FG_ERR_CANCELLED {
            return Err(GpuError::Cancelled);
        }
        let fatal = rc == sys::FG_ERR_OOM || rc == sys::FG_ERR_CUDA_STICKY;
        if fatal {
            invalidate_context();
        }
        Err(GpuError::Gpu { message: format!("{label}: {}", cstr(unsafe { sys::fg_last_error(self.raw) })), fatal })
    }

    fn check_sizes(p: &FgParams, lambda: &[f32], offsets: &[[f32; 2]], out: &[f32]) -> Result<(), GpuError> {
        let bad = |m: &str| Err(GpuError::Gpu { message: m.to_string(), fatal: false });
        if lambda.len() != p.in_w as usize * p.in_h as usize {
            return bad("lambda plane size does not match in_w * in_h");
        }
        if out.len() != p.out_w as usize * p.out_h as usize {
            return bad("output plane size does not match out_w * out_h");
        }
        if offsets.len() != p.n_samples as usize {
            return bad("offset count does not match sample count"); // src/wgpu/mod.rs:353-358, 490-495
        }
        Ok(())
    }

    /// wgpu::render_pixelwise_gpu (src/wgpu/mod.rs:336-345); `offsets_input` = Derived.offsets_input
    pub fn render_pixelwise(&self, p: &FgParams, lambda: &[f32], offsets_input: &[[f32; 2]], out: &mut [f32]) -> Result<(), GpuError> {
        Self::check_sizes(p, lambda, offsets_input, out)?;
        let rc = unsafe { sys::fg_render_pixelwise(self.raw, p, lambda.as_ptr(), offsets_input.as_ptr() as *const f32, out.as_mut_ptr()) };
        self.check(rc, "Pixel renderer")
    }

    /// wgpu::render_grainwise_gpu (src/wgpu/mod.rs:473-482); `offsets` = Derived.offsets
    pub fn render_grainwise(&self, p: &FgParams, lambda: &[f32], offsets: &[[f32; 2]], out: &mut [f32]) -> Result<(), GpuError> {
        Self::check_sizes(p, lambda, offsets, out)?;
        let rc = unsafe { sys::fg_render_grainwise(self.raw, p, lambda.as_ptr(), offsets.as_ptr() as *const f32, out.as_mut_ptr()) };
        self.check(rc, "Grain renderer")
    }

    /// All planes of Workspace::for_each_plane (src/color.rs:47-64) in one call; `cancel` is the viewer's
    /// CancelToken as a flag (non-zero = cancelled), honoured inside the kernel launches.
    pub fn render_planes(&self, p: &FgParams, pixelwise: bool, lambda: &[&[f32]], offsets: &[[f32; 2]], out: &mut [&mut [f32]],
                         cancel: Option<&AtomicI32>) -> Result<(), GpuError> {
        if lambda.len() != out.len() || lambda.is_empty() {
            return Err(GpuError::Gpu { message: "plane count mismatch".into(), fatal: false });
        }
        for (l, o) in lambda.iter().zip(out.iter()) {
            Self::check_sizes(p, l, offsets, o)?;
        }
        let lp: Vec<*const f32> = lambda.iter().map(|l| l.as_ptr()).collect();
        let op: Vec<*mut f32> = out.iter_mut().map(|o| o.as_mut_ptr()).collect();
        let algo = if pixelwise { sys::FG_ALGO_PIXEL } else { sys::FG_ALGO_GRAIN };
        let flag = cancel.map_or(std::ptr::null(), |c| c.as_ptr() as *const c_int);
        let rc = unsafe {
            sys::fg_render_planes_cancelable(self.raw, p, algo, lp.len() as c_int, lp.as_ptr(), offsets.as_ptr() as *const f32, op.as_ptr(), flag)
        };
        self.check(rc, if pixelwise { "Pixel renderer" } else { "Grain renderer" })
    }

    /// Progressive refinement: after the call `out` holds the render of samples [0, k_end) of `offsets`.
    #[allow(clippy::too_many_arguments)]
    pub fn refine_planes(&self, p: &FgParams, pixelwise: bool, lambda: &[&[f32]], offsets: &[[f32; 2]], k_begin: u32, k_end: u32,
                         out: &mut [&mut [f32]], cancel: Option<&AtomicI32>) -> Result<(), GpuError> {
        if lambda.len() != out.len() || lambda.is_empty() {
            return Err(GpuError::Gpu { message: "plane count mismatch".into(), fatal: false });
        }
        for (l, o) in lambda.iter().zip(out.iter()) {
            Self::check_sizes(p, l, offsets, o)?;
        }
        let lp: Vec<*const f32> = lambda.iter().map(|l| l.as_ptr()).collect();
        let op: Vec<*mut f32> = out.iter_mut().map(|o| o.as_mut_ptr()).collect();
        let algo = if pixelwise { sys::FG_ALGO_PIXEL } else { sys::FG_ALGO_GRAIN };
        let flag = cancel.map_or(std::ptr::null(), |c| c.as_ptr() as *const c_int);
        let rc = unsafe {
            sys::fg_refine_planes(self.raw, p, algo, lp.len() as c_int, lp.as_ptr(), offsets.as_ptr() as *const f32, k_begin, k_end, op.as_ptr(), flag)
        };
        self.check(rc, "Refinement")
    }

    /// u8 RGB in, u8 RGB out: load / lambda / store on the device (src/color.rs:116-170, src/model.rs:228-265, src/color.rs:66-114)
    pub fn render_rgb8(&self, p: &FgParams, pixelwise: bool, rgb_mode: bool, rgb_in: &[u8], offsets: &[[f32; 2]], rgb_out: &mut [u8]) -> Result<(), GpuError> {
        if rgb_in.len() != 3 * p.in_w as usize * p.in_h as usize || rgb_out.len() != 3 * p.out_w as usize * p.out_h as usize || offsets.len() != p.n_samples as usize {
            return Err(GpuError::Gpu { message: "buffer sizes do not match the parameter block".into(), fatal: false });
        }
        let algo = if pixelwise { sys::FG_ALGO_PIXEL } else { sys::FG_ALGO_GRAIN };
        let mode = if rgb_mode { sys::FG_COLOR_RGB } else { sys::FG_COLOR_LUMA };
        let rc = unsafe { sys::fg_render_rgb8(self.raw, p, algo, mode, rgb_in.as_ptr(), offsets.as_ptr() as *const f32, rgb_out.as_mut_ptr()) };
        self.check(rc, "RGB renderer")
    }
}

/// build_uniforms (src/wgpu/mod.rs:661-692) for callers that hold the reference's values as plain numbers.
#[allow(clippy::too_many_arguments)]
pub fn params_block(in_w: usize, in_h: usize, out_w: usize, out_h: usize, n_samples: u32, lognorm: bool, seed: u64, zoom: f32, delta: f32,
                    rm: f32, inv_e_pi_r2: f32, radius_mean: f32, log_mu: Option<f32>, log_sigma: Option<f32>) -> FgParams {
    FgParams {
        struct_size: std::mem::size_of::<FgParams>() as u32,
        in_w: in_w as u32,
        in_h: in_h as u32,
        out_w: out_w as u32,
        out_h: out_h as u32,
        n_samples,
        dist_kind: lognorm as u32,
        seeding: sys::FG_SEEDING_RAND_0_8, // rand 0.8.5, pinned by the reference's Cargo.lock:2478
        seed,                              // the full u64 (the wgpu path truncates, src/wgpu/mod.rs:676)
        zoom,
        delta,
        rm,
        inv_e_pi_r2,
        radius_mean,
        has_log: (lognorm && log_mu.is_some() && log_sigma.is_some()) as u32,
        radius_log_mu: log_mu.unwrap_or(0.0) as f64, // RadiusProfile keeps the f64 widenings (src/model.rs:112-113)
        radius_log_sigma: log_sigma.unwrap_or(0.0) as f64,
        row_begin: 0,
        row_end: 0,
        path: sys::FG_PATH_AUTO,
        reserved: 0,
    }
}
