/* fg.h -- C ABI of the B200-native film-grain engine (libfg_b200.so).
 *
 * This is the drop-in boundary behind the reference's `--device gpu` switch: every entry
 * point replaces one item of the reference's wgpu backend (paths relative to the
 * joseph-wardle/film_grain repository root):
 *
 *   fg_context_create / fg_context_destroy   <- wgpu::context()            src/wgpu/mod.rs:84-86
 *                                               (+ invalidate_context      src/wgpu/mod.rs:88-92)
 *   fg_render_pixelwise                      <- render_pixelwise_gpu       src/wgpu/mod.rs:336-345
 *   fg_render_grainwise                      <- render_grainwise_gpu       src/wgpu/mod.rs:473-482
 *   fg_params                                <- Uniforms / build_uniforms  src/wgpu/mod.rs:17-38, 661-692
 *   fg_last_error / return codes             <- RenderError::Gpu + handle_gpu_error  src/wgpu/mod.rs:727-752
 *   call sites on the reference side:           src/lib.rs:141-144, 154-163; src/bin/viewer.rs:973
 *
 * Only plain C types cross the boundary.  All `const T*` arguments of the host entry
 * points are HOST pointers owned by the caller for the duration of the call; nothing is
 * retained.  Calls block until the output buffer is complete.  A context may be used from
 * any thread, one call at a time (calls are serialised on an internal mutex).
 *
 * There is no CPU fallback: without a usable CUDA device fg_context_create fails with
 * FG_ERR_NO_DEVICE and the caller reports "gpu unavailable" (src/bin/viewer.rs:784-790).
 */
#ifndef FG_H
#define FG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_ABI_VERSION 1

/* return codes (0 = success, negative = failure; text via fg_last_error) */
enum {
    FG_OK = 0,
    FG_ERR_INVALID = -1,     /* bad argument / parameter block: context stays valid (wgpu Validation) */
    FG_ERR_OOM = -2,         /* device allocation failed: caller drops the context (wgpu OutOfMemory) */
    FG_ERR_CUDA_STICKY = -3, /* unrecoverable CUDA error: caller drops the context (wgpu Internal) */
    FG_ERR_NO_DEVICE = -4,   /* no CUDA device / device index out of range */
    FG_ERR_CANCELLED = -5,   /* cancel flag observed (RenderError::Cancelled, src/lib.rs:116-123) */
    FG_ERR_CUDA = -6         /* recoverable CUDA error (e.g. bad launch config): context stays valid */
};

enum { FG_DIST_CONST = 0, FG_DIST_LOGNORM = 1 };          /* RadiusDist, src/params.rs:6-10 */
enum { FG_STREAM_CELL = 1, FG_STREAM_PIXEL = 2 };         /* CELL_STREAM / PIXEL_STREAM, src/rng.rs:6-7 */
enum { FG_COLOR_LUMA = 0, FG_COLOR_RGB = 1 };             /* ColorMode, src/params.rs:19-23 */
enum { FG_ALGO_GRAIN = 1, FG_ALGO_PIXEL = 2 };            /* Algo (resolved), src/params.rs:12-17 */
/* SmallRng::seed_from_u64 flavour: rand 0.8.5 (pinned by Cargo.lock) uses rand_core's PCG32
 * fill; rand >= 0.9 forwards to xoshiro's SplitMix64 seeding. */
enum { FG_SEEDING_RAND_0_8 = 0, FG_SEEDING_RAND_0_9 = 1 };

/* Which kernel family serves fg_render_pixelwise (diagnostics / tests; AUTO in production).
 * Grain-wise: FG_PATH_DIRECT = the global-mask rasteriser, FG_PATH_TILED / FG_PATH_STAGED = the
 * shared-memory tile rasteriser, FG_PATH_AUTO = whichever the cost model expects to be faster. */
enum {
    FG_PATH_AUTO = 0,   /* staged (falling back to tiled, then direct); direct when there are too few samples per cell for the table to pay */
    FG_PATH_DIRECT = 1, /* per-sample regeneration (the reference's own structure) */
    FG_PATH_TILED = 2,  /* strip kernel generating its cell windows in shared memory */
    FG_PATH_STAGED = 3  /* cell table generated once per band in HBM, strip kernel loads windows */
};

/* Parameter block of one plane render = Params (src/params.rs:45-68) + Derived
 * (src/model.rs:167-179) reduced to what the integrators read; the Rust side fills it
 * exactly where build_uniforms does today (src/wgpu/mod.rs:661-692).  Differences from the
 * wgpu Uniforms: the seed is the full u64 (the wgpu path truncates, src/wgpu/mod.rs:676);
 * log_mu/log_sigma are the f64 widenings RadiusProfile keeps (src/model.rs:112-113);
 * `lanes` is internal. */
typedef struct fg_params {
    uint32_t struct_size;     /* = sizeof(fg_params); guards ABI drift */
    uint32_t in_w, in_h;      /* Derived.input_width/height  (lambda plane size) */
    uint32_t out_w, out_h;    /* Derived.output_width/height */
    uint32_t n_samples;       /* Params.n_samples (>= 1) */
    uint32_t dist_kind;       /* FG_DIST_* */
    uint32_t seeding;         /* FG_SEEDING_*; 0 for the pinned rand 0.8.5 */
    uint64_t seed;            /* Params.seed */
    float zoom;               /* Params.zoom */
    float delta;              /* Derived.delta */
    float rm;                 /* Derived.rm */
    float inv_e_pi_r2;        /* Derived.inv_e_pi_r2 (used only by the fused u8 entry point) */
    float radius_mean;        /* Params.radius_mean (RadiusProfile.mean_linear) */
    uint32_t has_log;         /* RadiusProfile.lognormal is Some */
    double radius_log_mu;     /* RadiusProfile.log_mu   (f32 widened to f64) */
    double radius_log_sigma;  /* RadiusProfile.log_sigma */
    uint32_t row_begin;       /* output rows [row_begin,row_end) to render; 0,0 = whole plane. */
    uint32_t row_end;         /*   rows outside the band are not written (multi-GPU row bands) */
    uint32_t path;            /* FG_PATH_*; 0 in production */
    uint32_t reserved;
} fg_params;

typedef struct fg_ctx fg_ctx;

/* per-call statistics (optional; zero-cost when not requested) */
typedef struct fg_stats {
    float kernel_ms;          /* device time of the compute kernels (CUDA events on the ctx stream) */
    float h2d_ms, d2h_ms;     /* host<->device copy time inside the call (host entry points) */
    uint32_t launches;        /* kernels launched by the call */
    uint32_t tiles_total;     /* pixel-wise tiled path: tiles rendered */
    uint32_t tiles_fallback;  /*   of which re-rendered by the direct kernel (capacity / lambda>=12) */
    uint64_t h2d_bytes, d2h_bytes;
    float strip_ms;           /* pixel-wise tiled path: device time of the (last) strip-kernel launch; grain-wise: of the rasterisation (last plane) */
    uint32_t strip_launches;  /*   strip-kernel launches of the call (row sub-bands; normally 1) */
    float table_ms;           /*   device time of the (last) band's thresholds + bitmap + cell-table kernels; grain-wise: grain generation */
    uint32_t table_reused;    /* 1: the render evaluated from the context's cached cell table (fg_set_table_cache), no table pass */
} fg_stats;

int fg_abi_version(void);
int fg_device_count(void);               /* 0 when no driver / no device */
const char* fg_error_string(int code);

/* Create a context on CUDA device `device` (ordinal).  Allocates a stream and lazily-grown
 * device buffer pools.  Replaces wgpu::context() (src/wgpu/mod.rs:84-86). */
int fg_context_create(fg_ctx** out, int device);
/* One context over SEVERAL devices of the box, for the reference's single-process caller (src/lib.rs:141-163,
 * src/main.rs:60): every render call on it splits the requested output rows into n_devices contiguous row bands, one
 * host thread and one stream per device; a band's margin cells are regenerated locally (no halo exchange, no
 * collective).  Host entry points: each device uploads the input rows its band reads and delivers its band into the
 * caller's buffer over its own PCIe link.  fg_render_planes_device: inputs and output live on devices[0]; the other
 * devices pull their lambda rows and store their band rows into devices[0]'s image over NVLink (peer access is
 * enabled here).  Results are bit-identical to a single-device context.  fg_render_rgb8_device, fg_dump_cells and
 * fg_measure_issue_peak run on devices[0]. */
int fg_context_create_multi(fg_ctx** out, const int* devices, int n_devices);
int fg_context_device_count(const fg_ctx* ctx); /* devices a context renders on (1 for fg_context_create) */
void fg_context_destroy(fg_ctx* ctx);
/* Last error text of this context (valid until the next call on it); "" if none. */
const char* fg_last_error(const fg_ctx* ctx);
/* Name of the kernel that evaluated (pixel-wise) / rasterised (grain-wise) the last render of this context:
 * "k_pixelwise_skew", "k_pixelwise_strip", "k_pixelwise_direct", "k_gw_tile", "k_gw_splat + k_gw_reduce"
 * (measurement labels; static storage). */
const char* fg_last_eval_kernel(const fg_ctx* ctx);
/* Optional cooperative cancel: *flag != 0 is polled by the host thread that waits for the render and forwarded to a
 * device word every CTA tests when it starts, so a render stops INSIDE a kernel launch (a 4K frame within ~2 ms), not
 * only between stages -> FG_ERR_CANCELLED; the output is then undefined. NULL disables (and removes the test). */
void fg_set_cancel_flag(fg_ctx* ctx, const volatile int* flag);
/* Statistics of the last render call on this context. */
void fg_get_stats(const fg_ctx* ctx, fg_stats* out);

/* ---- the two integrators, HOST buffers (the drop-in calls) ---------------------------
 * lambda: in_w*in_h f32 row-major (Plane.data, src/model.rs:14-18)
 * offsets_input: n_samples x [f32;2] = Derived.offsets_input (offsets / zoom, src/model.rs:209-212)
 * offsets:       n_samples x [f32;2] = Derived.offsets       (output pixels)
 * out: out_w*out_h f32 row-major, values k/N.  Mirrors render_pixelwise (src/pixelwise.rs:11-45)
 * and render_grainwise (src/grainwise.rs:12-124) bit for bit. */
int fg_render_pixelwise(fg_ctx* ctx, const fg_params* p, const float* lambda,
                        const float* offsets_input, float* out);
int fg_render_grainwise(fg_ctx* ctx, const fg_params* p, const float* lambda,
                        const float* offsets, float* out);

/* n_planes planes in one call with shared parameters (RGB renders 3 planes with the same
 * seed and offsets, src/color.rs:56-60): one upload, one batched launch, one download.
 * algo = FG_ALGO_PIXEL (offsets = offsets_input) or FG_ALGO_GRAIN (offsets = offsets).
 * When the output planes are one contiguous block of page-locked host memory (cudaHostAlloc /
 * cudaHostRegister; out[pl] == out[0] + pl*out_w*out_h) the kernels store their results straight
 * into it over PCIe while they run, instead of a staged device->host copy afterwards; pageable or
 * scattered planes take the staged copy.  Same results either way. */
int fg_render_planes(fg_ctx* ctx, const fg_params* p, int algo, int n_planes,
                     const float* const* lambda, const float* offsets, float* const* out);

/* Same call with a PER-CALL cancel flag (render_with_input_image_cancelable, src/lib.rs:116-132): *cancel != 0
 * is polled between the stages and row sub-bands of the render and inside its launches (see fg_set_cancel_flag) -> FG_ERR_CANCELLED.  The pointer is used only
 * for the duration of this call (unlike fg_set_cancel_flag, which stays registered), so concurrent callers of
 * a shared context cannot overwrite or outlive each other's flags.  cancel == NULL: identical to fg_render_planes. */
int fg_render_planes_cancelable(fg_ctx* ctx, const fg_params* p, int algo, int n_planes,
                                const float* const* lambda, const float* offsets, float* const* out,
                                const volatile int* cancel);

/* ---- Viewer-grade re-render (the reference's interactive caller is src/bin/viewer.rs:944-1067: a worker
 * that re-renders on every parameter change, latest job wins, CancelToken polled inside the integrators) -------------
 *
 * fg_set_table_cache(ctx, 1): the context keeps the cell table of its last whole-frame pixel-wise render.  A later
 * render whose table identity is the same -- seed, seeding, delta, radius model, plane count and the CONTENT of the
 * lambda planes (128-bit hash taken on the device) -- and whose cell rectangle lies inside the cached one (the table is
 * built with a margin) skips the table pass: n_samples, the offsets (sigma) and the zoom are free to change.
 * fg_stats.table_reused reports it.  Off by default (a batch render pays the 0.1 ms hash for nothing). */
void fg_set_table_cache(fg_ctx* ctx, int enable);
/* Progressive refinement: render samples [k_begin, k_end) of the n_samples offsets and merge them into the context's
 * running image, so a preview can show N = 16 at once and refine to the full N without evaluating a sample twice.
 * After the call `out` (and the context) hold the render of samples [0, k_end): bit-identical to fg_render_planes with
 * n_samples = k_end and the first k_end offsets (sample counts are merged as integers).  k_begin = 0 starts a
 * refinement; k_begin > 0 must equal the k_end of the previous fg_refine_planes call on this context with the same
 * geometry, else FG_ERR_INVALID.  p->n_samples = the total N (its offsets size the cell table once for all slices).
 * `cancel` as in fg_render_planes_cancelable (may be NULL).  Single frame, one device (a multi-device context uses
 * its first device). */
int fg_refine_planes(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda,
                     const float* offsets, uint32_t k_begin, uint32_t k_end, float* const* out,
                     const volatile int* cancel);

/* ---- DEVICE-pointer variants (inputs already resident in HBM; asynchronous on the
 * context stream unless stream_sync != 0).  d_lambda / d_out hold n_planes planes
 * back to back (plane stride in_w*in_h resp. out_w*out_h floats). */
int fg_render_planes_device(fg_ctx* ctx, const fg_params* p, int algo, int n_planes,
                            const float* d_lambda, const float* d_offsets, float* d_out,
                            int stream_sync);
/* CUDA stream handle (cudaStream_t) the context launches on, as an integer. */
uint64_t fg_context_stream(const fg_ctx* ctx);
int fg_context_synchronize(fg_ctx* ctx);

/* ---- fused colour path: decoded 8-bit interleaved RGB in, 8-bit
 * interleaved RGB out; load (src/color.rs:158-213), normalize + lambda (src/model.rs:228-265)
 * and store (src/color.rs:66-114, 233-239) run on the device.  lambda uses a 256-entry
 * table computed on the HOST with the host libm (bit-identical to what the Rust host
 * computes) in RGB mode. */
int fg_render_rgb8(fg_ctx* ctx, const fg_params* p, int algo, int color_mode,
                   const uint8_t* rgb_in, const float* offsets, uint8_t* rgb_out);
int fg_render_rgb8_device(fg_ctx* ctx, const fg_params* p, int algo, int color_mode,
                          const uint8_t* d_rgb_in, const float* d_offsets, uint8_t* d_rgb_out,
                          int stream_sync);

/* ---- debug / parity: grain realisation of individual cells, exactly as the integrators
 * draw them (counts, centres, clamped radii).  ij: n x [i32;2]; lambda_cell: n f32 (the
 * lambda the cell sees); stream_kind FG_STREAM_CELL (pixel-wise cells, mean lambda*delta^2)
 * or FG_STREAM_PIXEL (grain-wise unit cells, mean lambda).  q_out: n counts;
 * grains_out: n x cap x [cx, cy, r] (first `cap` grains of each cell). */
int fg_dump_cells(fg_ctx* ctx, const fg_params* p, int stream_kind, const int32_t* ij,
                  const float* lambda_cell, size_t n, uint32_t cap, uint32_t* q_out,
                  float* grains_out);

/* Measured issue-rate microbenchmark (FFMA / IMAD / LOP3 dependent chains over all SMs):
 * lane-instructions per second for each pipe mix; used as the ALU roofline denominator.
 * out[0]=FFMA, out[1]=IMAD(u32), out[2]=LOP3/IADD mix, out[3]=DFMA.  Units: 1e9 lane-ops/s. */
int fg_measure_issue_peak(fg_ctx* ctx, double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* FG_H */
