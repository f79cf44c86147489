/* fg_host.h -- C entry points of the HOST-side mirror of the reference's library API
 * (film_grain_b200/host/film_grain.hpp: Params / ParamsBuilder::build, derive_common,
 * make_offsets, choose_algorithm, normalize_plane + lambda_plane, render_with_input_image),
 * exported from libfg_b200.so so that non-C++ callers (the Python tests, bench.py) drive the
 * same host code a C++/Rust caller would.  Everything here runs above the engine ABI (fg.h).
 *
 *   fgh_params                     <- ParamsBuilder                 src/params.rs:70-91
 *   fgh_derive                     <- ParamsBuilder::build + derive_common + choose_algorithm
 *                                     src/params.rs:141-180, src/model.rs:181-226, src/choose.rs:4-26
 *   fgh_lambda_from_plane          <- normalize_plane + lambda_plane src/model.rs:228-265
 *   fgh_render_with_input_image    <- render_with_input_image(_cancelable) with Device::Gpu
 *                                     src/lib.rs:78-93, 134-173
 *   fgh_context / fgh_invalidate_context <- wgpu::context / invalidate_context src/wgpu/mod.rs:84-92
 *   fgh_render_file / fgh_load_image / fgh_save_image <- render(params), image::open, save_with_format
 *                                     src/lib.rs:57-71, src/color.rs:26-29
 */
#ifndef FG_HOST_H
#define FG_HOST_H
#include "fg.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { FGH_OK = 0, FGH_ERR_PARAMS = -101, FGH_ERR_GPU = -102, FGH_ERR_CANCELLED = -103, FGH_ERR_MESSAGE = -104 };

typedef struct fgh_params {      /* ParamsBuilder (src/params.rs:70-91); defaults = CLI defaults */
    int32_t radius_dist;         /* 0 const, 1 lognorm */
    float radius_mean;
    float radius_stddev;
    float zoom;
    float sigma_px;
    uint32_t n_samples;
    int32_t algo;                /* 0 auto, 1 grain, 2 pixel */
    int32_t max_radius_kind;     /* 0 absolute, 1 quantile */
    float max_radius_value;
    int32_t has_cell_delta;
    float cell_delta;
    int32_t color_mode;          /* 0 luma, 1 rgb */
    int32_t has_size;
    uint32_t size_w;
    int32_t has_size_h;
    uint32_t size_h;
    uint64_t seed;
} fgh_params;

typedef struct fgh_derived {     /* Derived (src/model.rs:167-179) + the resolved algorithm */
    uint64_t input_width, input_height, output_width, output_height;
    float inv_e_pi_r2, rm, delta;
    float radius_stddev;         /* Params.radius_stddev as derived */
    int32_t has_log;
    int32_t algorithm;           /* 1 grain, 2 pixel */
    double log_mu, log_sigma;
    fg_params block;             /* the engine parameter block build_uniforms would fill */
} fgh_derived;

/* text of the last failure on this thread ("" if none) */
const char* fgh_last_error(void);
/* build + derive: offsets / offsets_input receive n_samples x [f32;2] (may be NULL) */
int fgh_derive(const fgh_params* p, uint64_t in_w, uint64_t in_h, fgh_derived* out, float* offsets,
               float* offsets_input);
int fgh_lambda_from_plane(const float* plane, uint64_t w, uint64_t h, float inv_e_pi_r2, float* lambda_out);
/* rgb: decoded 8-bit interleaved RGB, w*h*3 bytes; rgb_out: out_w*out_h*3 bytes.
 * fused = 0: host load/lambda/store + device integrator (the reference's data flow);
 * fused = 1: load/lambda/store on the device (fg_render_rgb8). */
int fgh_render_with_input_image(const fgh_params* p, const uint8_t* rgb, uint64_t w, uint64_t h, int fused,
                                int device, const volatile int* cancel, uint8_t* rgb_out,
                                uint64_t out_capacity, fgh_derived* info);
/* process-global cached context; NULL on failure.  The returned pointer stays valid until the process ends
 * (a context handed out here is not destroyed by fgh_invalidate_context or by a device switch). */
fg_ctx* fgh_context(int device);
void fgh_invalidate_context(void);
/* ---- image files ----------------------------------------------------------------------------------
 * fgh_render_file <- render(params) src/lib.rs:57-71: image::open (src/color.rs:26-29), ROI crop (src/color.rs:215-231),
 * render on the device, fs::create_dir_all of the output's parent, save_with_format with the format resolved from
 * `format_token` or the output extension, PNG when there is none (src/lib.rs:188-203).  This build reads and writes PNG
 * (8-bit and 1/2/4-bit grey / palette, non-interlaced; alpha dropped like to_rgb32f) and binary PNM (P5 / P6, maxval
 * 255); other known formats -> FGH_ERR_MESSAGE "not built into this engine", unknown tokens -> the reference's message.
 * roi4 = {x0, y0, x1, y1} (exclusive end) or NULL. */
int fgh_render_file(const fgh_params* p, const char* input_path, const char* output_path, const char* format_token,
                    const uint32_t* roi4, int fused, int device, const volatile int* cancel, fgh_derived* info);
/* image::open -> 8-bit interleaved RGB in a malloc'ed buffer (release with fgh_free). */
int fgh_load_image(const char* path, uint8_t** rgb, uint64_t* w, uint64_t* h);
void fgh_free(void* p);
/* save_with_format: format from `format_token`, else from the extension of `path`, else PNG. */
int fgh_save_image(const char* path, const uint8_t* rgb, uint64_t w, uint64_t h, const char* format_token);

/* The libm logf restatement the fused luma path runs on the device (csrc/fg_logf.h), host build:
 * lets the CPU tests compare it with the platform's logf. */
void fgh_logf_restated(const float* x, uint64_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
