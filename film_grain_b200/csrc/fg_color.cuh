// fg_color.cuh -- colour load/store fused around the integrators (src/color.rs), u8 in/out.
//
//   RGB mode : u8 -> lambda through a 256-entry table built on the HOST with the host libm
//              (c/255 -> clamp01 -> normalize_plane -> lambda_plane; src/color.rs:158-179,
//              src/model.rs:228-265), so lambda is bit-identical to what the Rust host computes;
//              3 planes rendered in one batched launch; store = to_u8 (src/color.rs:98-112, 237-239).
//   Luma mode: Y/Cb/Cr split (src/color.rs:181-213) with un-fused f32 arithmetic on the device;
//              lambda(Y) needs logf per pixel: fg_logf.h restates the libm algorithm the Rust host
//              calls (glibc / musl logf; checked against libm over every positive normal float),
//              so lambda is bit-identical here too;
//              chroma nearest-resize (src/model.rs:77-98) + Y'CbCr -> RGB (src/color.rs:68-97).
#pragma once
#include "fg_ctx.cuh"
#include "fg_kernels.cuh"
#include "fg_logf.h"

namespace {

using namespace fg;

int render_planes_device_locked(fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int algo, int n_planes,
                                const float* d_lambda, const float* d_offsets, float* d_out);
RenderConsts make_consts(const fg_params* p, const float* offsets_host);
int check_offsets(fg_ctx* ctx, const fg_params* p, const float* offsets_host);

__device__ __forceinline__ float clamp01_dev(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
__device__ __forceinline__ uint8_t to_u8_dev(float v) { // src/color.rs:237-239
    return (uint8_t)floorf(__fadd_rn(__fmul_rn(clamp01_dev(v), 255.0f), 0.5f));
}

// RGB: interleaved u8 -> 3 lambda planes through the host-built LUT
__global__ void __launch_bounds__(256) k_load_rgb_lut(const uint8_t* __restrict__ rgb, size_t npix,
                                                       const float* __restrict__ lut, float* __restrict__ lambda) {
    __shared__ float s_lut[256];
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < npix; t += (size_t)gridDim.x * 256) {
        lambda[t] = s_lut[rgb[3 * t + 0]];
        lambda[npix + t] = s_lut[rgb[3 * t + 1]];
        lambda[2 * npix + t] = s_lut[rgb[3 * t + 2]];
    }
}

// Luma: interleaved u8 -> lambda(Y), Cb, Cr planes (src/color.rs:181-213, src/model.rs:228-265)
__global__ void __launch_bounds__(256) k_load_luma(const uint8_t* __restrict__ rgb, size_t npix, float inv_e_pi_r2,
                                                    float* __restrict__ lambda, float* __restrict__ cb,
                                                    float* __restrict__ cr) {
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < npix; t += (size_t)gridDim.x * 256) {
        float r = clamp01_dev(__fdiv_rn((float)rgb[3 * t + 0], 255.0f));
        float g = clamp01_dev(__fdiv_rn((float)rgb[3 * t + 1], 255.0f));
        float b = clamp01_dev(__fdiv_rn((float)rgb[3 * t + 2], 255.0f));
        float luma = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, r), __fmul_rn(0.7152f, g)), __fmul_rn(0.0722f, b));
        cb[t] = __fdiv_rn(__fsub_rn(b, luma), 1.8556f);
        cr[t] = __fdiv_rn(__fsub_rn(r, luma), 1.5748f);
        float y = clamp01_dev(luma);
        const float EPS = 1e-6f;
        float hi = __fsub_rn(1.0f, EPS);
        float clamped = y < 0.0f ? 0.0f : (y > hi ? hi : y); // normalize_plane (max <= 1+eps branch) + lambda_plane clamp
        float safe = fmaxf(__fsub_rn(1.0f, clamped), EPS);
        float ln = logf_libm(safe); // the host libm's logf, restated (fg_logf.h): bit-identical lambda
        float activity = __fmul_rn(-inv_e_pi_r2, ln);
        lambda[t] = fminf(activity, 1.0e6f);
    }
}

__global__ void __launch_bounds__(256) k_store_rgb(const float* __restrict__ planes, size_t npix_plane, int out_w,
                                                    int row_begin, int row_end, uint8_t* __restrict__ rgb) {
    size_t n = (size_t)(row_end - row_begin) * out_w, base = (size_t)row_begin * out_w;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (size_t)gridDim.x * 256) {
        size_t i = base + t;
        rgb[3 * i + 0] = to_u8_dev(planes[i]);
        rgb[3 * i + 1] = to_u8_dev(planes[npix_plane + i]);
        rgb[3 * i + 2] = to_u8_dev(planes[2 * npix_plane + i]);
    }
}

// Plane::resize_nearest source index (src/model.rs:85-92)
__device__ __forceinline__ int nearest_src(int x, float scale, int n_src) {
    float s = __fsub_rn(__fmul_rn(__fadd_rn((float)x, 0.5f), scale), 0.5f);
    float hi = (float)(n_src - 1);
    s = s < 0.0f ? 0.0f : (s > hi ? hi : s);
    return (int)roundf(s);
}

__global__ void __launch_bounds__(256) k_store_luma(const float* __restrict__ yplane, const float* __restrict__ cb,
                                                     const float* __restrict__ cr, int in_w, int in_h, int out_w,
                                                     int out_h, int row_begin, int row_end, uint8_t* __restrict__ rgb) {
    size_t n = (size_t)(row_end - row_begin) * out_w, base = (size_t)row_begin * out_w;
    const bool same = (in_w == out_w && in_h == out_h);
    float scale_x = __fdiv_rn((float)in_w, (float)out_w), scale_y = __fdiv_rn((float)in_h, (float)out_h);
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (size_t)gridDim.x * 256) {
        size_t i = base + t;
        int x = (int)(i % out_w), y = (int)(i / out_w);
        size_t src = same ? i : (size_t)nearest_src(y, scale_y, in_h) * in_w + nearest_src(x, scale_x, in_w);
        float y_val = yplane[i], cb_val = cb[src], cr_val = cr[src];
        float r = clamp01_dev(__fadd_rn(y_val, __fmul_rn(1.5748f, cr_val)));
        float b = clamp01_dev(__fadd_rn(y_val, __fmul_rn(1.8556f, cb_val)));
        float g_un = __fdiv_rn(__fsub_rn(__fsub_rn(y_val, __fmul_rn(0.2126f, r)), __fmul_rn(0.0722f, b)), 0.7152f);
        float g = clamp01_dev(g_un);
        rgb[3 * i + 0] = to_u8_dev(r);
        rgb[3 * i + 1] = to_u8_dev(g);
        rgb[3 * i + 2] = to_u8_dev(b);
    }
}

// host-built u8 -> lambda table (RGB mode): exactly the reference's host arithmetic
void build_lambda_lut(float inv_e_pi_r2, float lut[256]) {
    const float EPS = 1e-6f;
    for (int v = 0; v < 256; ++v) {
        volatile float c = (float)v / 255.0f;                 // image 0.25.8 to_rgb32f
        float cl = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);   // clamp01, src/color.rs:171-173
        float hi = 1.0f - EPS;
        float n = cl < 0.0f ? 0.0f : (cl > hi ? hi : cl);     // normalize_plane (already normalised), :247
        volatile float one_minus = 1.0f - n;
        float safe = one_minus > EPS ? one_minus : EPS;       // lambda_plane, src/model.rs:259-262
        volatile float ln = logf(safe);
        volatile float activity = -inv_e_pi_r2 * ln;
        lut[v] = activity < 1.0e6f ? activity : 1.0e6f;
    }
}

int color_render_device_impl(fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int algo, int color_mode,
                             const uint8_t* d_rgb_in, const float* d_offsets, uint8_t* d_rgb_out) {
    if (algo != FG_ALGO_PIXEL && algo != FG_ALGO_GRAIN) return set_err(ctx, FG_ERR_INVALID, "algo must be FG_ALGO_PIXEL or FG_ALGO_GRAIN");
    if (color_mode != FG_COLOR_LUMA && color_mode != FG_COLOR_RGB) return set_err(ctx, FG_ERR_INVALID, "unknown color_mode");
    if (!(std::isfinite(p->inv_e_pi_r2) && p->inv_e_pi_r2 > 0.0f)) return set_err(ctx, FG_ERR_INVALID, "inv_e_pi_r2 must be finite and > 0");
    const size_t in_pix = (size_t)p->in_w * p->in_h, out_pix = (size_t)p->out_w * p->out_h;
    const int n_planes = color_mode == FG_COLOR_RGB ? 3 : 1;
    int rc;
    if ((rc = ensure(ctx, ctx->lambda, in_pix * n_planes * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->out, out_pix * n_planes * sizeof(float)))) return rc;
    const unsigned lblocks = (unsigned)std::min<size_t>((in_pix + 255) / 256, (size_t)ctx->sm_count * 16);
    const size_t band_pix = (size_t)(c.row_end - c.row_begin) * p->out_w;
    const unsigned sblocks = (unsigned)std::min<size_t>((band_pix + 255) / 256, (size_t)ctx->sm_count * 16);
    if (color_mode == FG_COLOR_RGB) {
        float lut[256];
        build_lambda_lut(p->inv_e_pi_r2, lut);
        if ((rc = ensure(ctx, ctx->lut, sizeof lut))) return rc;
        FG_CUDA(ctx, cudaMemcpyAsync(ctx->lut.p, lut, sizeof lut, cudaMemcpyHostToDevice, ctx->stream));
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // lut is a stack buffer
        k_load_rgb_lut<<<lblocks, 256, 0, ctx->stream>>>(d_rgb_in, in_pix, (const float*)ctx->lut.p, (float*)ctx->lambda.p);
    } else {
        if ((rc = ensure(ctx, ctx->chroma, in_pix * 2 * sizeof(float)))) return rc;
        k_load_luma<<<lblocks, 256, 0, ctx->stream>>>(d_rgb_in, in_pix, p->inv_e_pi_r2, (float*)ctx->lambda.p,
                                                      (float*)ctx->chroma.p, (float*)ctx->chroma.p + in_pix);
    }
    FG_CUDA(ctx, cudaGetLastError());
    ctx->stats.launches += 1;
    rc = render_planes_device_locked(ctx, p, c, algo, n_planes, (const float*)ctx->lambda.p, d_offsets, (float*)ctx->out.p);
    if (rc) return rc;
    if (color_mode == FG_COLOR_RGB)
        k_store_rgb<<<sblocks, 256, 0, ctx->stream>>>((const float*)ctx->out.p, out_pix, (int)p->out_w, c.row_begin, c.row_end, d_rgb_out);
    else
        k_store_luma<<<sblocks, 256, 0, ctx->stream>>>((const float*)ctx->out.p, (const float*)ctx->chroma.p,
                                                       (const float*)ctx->chroma.p + in_pix, (int)p->in_w, (int)p->in_h,
                                                       (int)p->out_w, (int)p->out_h, c.row_begin, c.row_end, d_rgb_out);
    FG_CUDA(ctx, cudaGetLastError());
    ctx->stats.launches += 1;
    return FG_OK;
}

int color_render_device(fg_ctx* ctx, const fg_params* p, int algo, int color_mode, const uint8_t* d_rgb_in,
                        const float* d_offsets, uint8_t* d_rgb_out, int stream_sync) {
    std::vector<float> off((size_t)p->n_samples * 2);
    FG_CUDA(ctx, cudaMemcpyAsync(off.data(), d_offsets, off.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = check_offsets(ctx, p, off.data());
    if (rc) return rc;
    RenderConsts c = make_consts(p, off.data());
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    rc = color_render_device_impl(ctx, p, c, algo, color_mode, d_rgb_in, d_offsets, d_rgb_out);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (stream_sync) {
        FG_CUDA(ctx, wait_stream(ctx));
        cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev[1], ctx->ev[2]);
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    }
    return FG_OK;
}

int color_render_host(fg_ctx* ctx, const fg_params* p, int algo, int color_mode, const uint8_t* rgb_in,
                      const float* offsets, uint8_t* rgb_out) {
    int rc = check_offsets(ctx, p, offsets);
    if (rc) return rc;
    RenderConsts c = make_consts(p, offsets);
    const size_t in_bytes = (size_t)p->in_w * p->in_h * 3, out_bytes = (size_t)p->out_w * p->out_h * 3;
    if ((rc = ensure(ctx, ctx->rgb_in, in_bytes))) return rc;
    if ((rc = ensure(ctx, ctx->rgb_out, out_bytes))) return rc;
    if ((rc = ensure(ctx, ctx->offsets, (size_t)p->n_samples * 2 * sizeof(float)))) return rc;
    cudaStream_t s = ctx->stream;
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->rgb_in.p, rgb_in, in_bytes, cudaMemcpyHostToDevice, s));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->offsets.p, offsets, (size_t)p->n_samples * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
    rc = color_render_device_impl(ctx, p, c, algo, color_mode, (const uint8_t*)ctx->rgb_in.p, (const float*)ctx->offsets.p,
                                  (uint8_t*)ctx->rgb_out.p);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[2], s));
    if (cancel_armed(ctx)) { // the copy into pageable memory would block the host until the kernels are done
        FG_CUDA(ctx, wait_stream(ctx));
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    }
    const size_t band_off = (size_t)c.row_begin * p->out_w * 3, band_bytes = (size_t)(c.row_end - c.row_begin) * p->out_w * 3;
    FG_CUDA(ctx, cudaMemcpyAsync(rgb_out + band_off, (uint8_t*)ctx->rgb_out.p + band_off, band_bytes, cudaMemcpyDeviceToHost, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
    FG_CUDA(ctx, wait_stream(ctx));
    cudaEventElapsedTime(&ctx->stats.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->stats.d2h_ms, ctx->ev[2], ctx->ev[3]);
    ctx->stats.h2d_bytes = in_bytes + (size_t)p->n_samples * 8;
    ctx->stats.d2h_bytes = band_bytes;
    if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    return FG_OK;
}

} // namespace
