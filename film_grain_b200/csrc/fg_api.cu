// fg_api.cu -- C ABI (include/fg.h) over the sm_100a kernels.  Host-side plumbing only:
// validation, device buffer pools, uploads/downloads, launches, error mapping.
//
// There is no CPU fallback anywhere in this file: every render either runs the CUDA
// kernels or returns an error code.
#include "../../include/fg.h"

#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstdlib>
#include "fg_ctx.cuh"
#include "fg_kernels.cuh"
#include "fg_pixel_host.cuh"
#include "fg_gw_tile.cuh"
#include "fg_color.cuh"
#include "fg_zig_tables.h"
#include <algorithm>

using namespace fg;

namespace {

float uniform_scale(float low, float high) { // rand 0.8.5 UniformFloat<f32>::new
    uint32_t mr = (0xFFFFFFFFu >> 9) | 0x3F800000u;
    float max_rand;
    std::memcpy(&max_rand, &mr, 4);
    max_rand = max_rand - 1.0f;
    float scale = high - low;
    for (;;) {
        volatile float t = scale * max_rand;
        volatile float u = t + low;
        if (!(u >= high)) break;
        uint32_t b;
        std::memcpy(&b, &scale, 4);
        b -= 1;
        std::memcpy(&scale, &b, 4);
    }
    return scale;
}

int validate(fg_ctx* ctx, const fg_params* p) {
    if (!p) return set_err(ctx, FG_ERR_INVALID, "params is NULL");
    if (p->struct_size != sizeof(fg_params))
        return set_err(ctx, FG_ERR_INVALID, "fg_params.struct_size does not match this library (ABI drift)");
    if (p->in_w == 0 || p->in_h == 0) return set_err(ctx, FG_ERR_INVALID, "input image is empty");
    if (p->out_w == 0 || p->out_h == 0) return set_err(ctx, FG_ERR_INVALID, "output dimensions must be positive");
    if (p->in_w > (1u << 30) || p->in_h > (1u << 30) || p->out_w > (1u << 30) || p->out_h > (1u << 30))
        return set_err(ctx, FG_ERR_INVALID, "image dimensions too large");
    if (p->n_samples == 0) return set_err(ctx, FG_ERR_INVALID, "n_samples must be >= 1");
    if (!(std::isfinite(p->zoom) && p->zoom > 0.0f)) return set_err(ctx, FG_ERR_INVALID, "zoom must be finite and > 0");
    if (!(std::isfinite(p->delta) && p->delta > 0.0f)) return set_err(ctx, FG_ERR_INVALID, "delta must be finite and > 0");
    if (!(std::isfinite(p->rm) && p->rm > 0.0f)) return set_err(ctx, FG_ERR_INVALID, "rm must be finite and > 0");
    if (!std::isfinite(p->radius_mean)) return set_err(ctx, FG_ERR_INVALID, "radius_mean must be finite");
    if (p->dist_kind > FG_DIST_LOGNORM) return set_err(ctx, FG_ERR_INVALID, "unknown dist_kind");
    if (p->dist_kind == FG_DIST_LOGNORM && p->has_log &&
        !(std::isfinite(p->radius_log_mu) && std::isfinite(p->radius_log_sigma) && p->radius_log_sigma >= 0.0))
        return set_err(ctx, FG_ERR_INVALID, "log-normal parameters must be finite, sigma >= 0");
    if (p->row_end != 0 || p->row_begin != 0) {
        if (p->row_begin >= p->row_end || p->row_end > p->out_h)
            return set_err(ctx, FG_ERR_INVALID, "row band must satisfy row_begin < row_end <= out_h");
    }
    if (p->path > FG_PATH_STAGED) return set_err(ctx, FG_ERR_INVALID, "unknown path");
    return FG_OK;
}

RenderConsts make_consts(const fg_params* p, const float* offsets_host) {
    RenderConsts c{};
    c.seed_cell = p->seed ^ 0xA24B1C30BEBCCF59ULL;
    c.seed_pixel = p->seed ^ 0x6935FA5C55F65F1BULL;
    c.seeding = p->seeding;
    c.in_w = (int)p->in_w; c.in_h = (int)p->in_h; c.out_w = (int)p->out_w; c.out_h = (int)p->out_h;
    c.n = p->n_samples;
    c.zoom = p->zoom;
    c.inv_zoom = 1.0f / p->zoom;
    c.delta = p->delta;
    c.uscale_cell = uniform_scale(0.0f, p->delta);
    c.uscale_unit = uniform_scale(0.0f, 1.0f);
    c.inv_samples = 1.0f / (float)(p->n_samples < 1 ? 1 : p->n_samples);
    c.rad.lognorm = (p->dist_kind == FG_DIST_LOGNORM && p->has_log) ? 1u : 0u;
    c.rad.mean_linear = p->radius_mean;
    c.rad.rm = p->rm;
    c.rad.mu = p->radius_log_mu;
    c.rad.sigma = p->radius_log_sigma;
    if (p->row_begin == 0 && p->row_end == 0) { c.row_begin = 0; c.row_end = (int)p->out_h; }
    else { c.row_begin = (int)p->row_begin; c.row_end = (int)p->row_end; }
    float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (uint32_t k = 0; k < p->n_samples; ++k) {
        float ox = offsets_host[2 * k], oy = offsets_host[2 * k + 1];
        mnx = fminf(mnx, ox); mxx = fmaxf(mxx, ox); mny = fminf(mny, oy); mxy = fmaxf(mxy, oy);
    }
    c.off_min_x = mnx; c.off_max_x = mxx; c.off_min_y = mny; c.off_max_y = mxy;
    return c;
}

int check_offsets(fg_ctx* ctx, const fg_params* p, const float* offsets_host) {
    for (uint32_t k = 0; k < 2 * p->n_samples; ++k)
        if (!std::isfinite(offsets_host[k])) return set_err(ctx, FG_ERR_INVALID, "offsets must be finite");
    return FG_OK;
}

int init_tables(fg_ctx* ctx) {
    if (ctx->tables_ready) return FG_OK;
    FG_CUDA(ctx, cudaMemcpyToSymbolAsync(kZigX, FG_ZIG_NORM_X_INIT, sizeof(double) * 257, 0, cudaMemcpyHostToDevice, ctx->stream));
    FG_CUDA(ctx, cudaMemcpyToSymbolAsync(kZigF, FG_ZIG_NORM_F_INIT, sizeof(double) * 257, 0, cudaMemcpyHostToDevice, ctx->stream));
    ctx->tables_ready = true;
    return FG_OK;
}

struct ToU64 {
    __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; }
};

__global__ void k_total(const uint32_t* counts, const uint64_t* excl, size_t n, uint64_t* total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *total = n ? excl[n - 1] + counts[n - 1] : 0;
}

// ---- pixel-wise on device-resident planes ---------------------------------------------
int pixelwise_device(fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int n_planes, const float* d_lambda,
                     const float* d_offsets, float* d_out) {
    const size_t in_stride = (size_t)p->in_w * p->in_h, out_stride = (size_t)p->out_w * p->out_h;
    const int band_rows = c.row_end - c.row_begin;
    uint32_t path = p->path;
    if (p->path != FG_PATH_AUTO && p->path != FG_PATH_STAGED) FG_CUDA(ctx, flush_upload(ctx));
    if (path == FG_PATH_AUTO) {
        // The cell table costs ~ (6 / planes + 6.5) ps per Boolean-model cell of the band whatever N is (the
        // first-draw bitmap is shared by the planes); evaluating from it costs ~6 ps per sample, regenerating
        // the visited cells per sample (the reference's structure, k_pixelwise_direct) ~150 ps for a 3 x 3
        // cell window.  Measured on a B200 (tools/path_probe.py, 4K plane): direct wins up to N = 8 at
        // delta = 0.1 and up to N = 32 at delta = 0.05.  Few samples per cell -> regenerate.
        const double per_axis = 2.0 * (double)p->rm / (double)p->delta + 1.0;
        const double direct_ps = 150.0 * per_axis * per_axis / 9.0 - 6.0;
        const double samples_per_cell = (double)p->n_samples * (double)p->zoom * (double)p->zoom * (double)p->delta * (double)p->delta;
        path = (samples_per_cell * direct_ps < 6.0 / n_planes + 6.5) ? FG_PATH_DIRECT : FG_PATH_STAGED;
    }
    if (path != FG_PATH_STAGED) FG_CUDA(ctx, flush_upload(ctx));
    if (path == FG_PATH_TILED || path == FG_PATH_STAGED) {
        int rc = tile_render(ctx, p, c, n_planes, d_lambda, d_offsets, d_out, path);
        if (rc != 1) return rc; // 1 = "tiled path not applicable, use direct"
    }
    FG_CUDA(ctx, flush_upload(ctx));
    // grid.y is limited to 65 535 CTAs of 8 rows: taller bands (validate() admits 2^30 rows) go in row chunks
    const int rows_per_launch = 65535 * 8;
    for (int y = c.row_begin; y < c.row_end; y += rows_per_launch) {
        RenderConsts cb = c;
        cb.row_begin = y;
        cb.row_end = std::min(c.row_end, y + rows_per_launch);
        dim3 grid((p->out_w + 31) / 32, (unsigned)((cb.row_end - cb.row_begin + 7) / 8), n_planes), block(32, 8);
        k_pixelwise_direct<<<grid, block, 0, ctx->stream>>>(d_lambda, in_stride, (const float2*)d_offsets, d_out, out_stride, cb);
        ctx->eval_kernel = "k_pixelwise_direct";
        ctx->stats.launches += 1;
        FG_CUDA(ctx, cudaGetLastError());
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    }
    (void)band_rows;
    return FG_OK;
}

// ---- grain-wise on a device-resident plane ----------------------------------------------
int grainwise_device_plane(fg_ctx* ctx, const fg_params* p, const RenderConsts& c, const float* d_lambda,
                           const float* d_offsets, float* d_out) {
    // input rows whose grains can reach the output band (all rows for a full-plane render)
    int iy0 = 0, iy1 = (int)p->in_h;
    if (c.row_begin > 0 || c.row_end < (int)p->out_h) {
        const double reach = (double)p->rm * p->zoom + 2.0;
        double lo = ((double)c.row_begin - reach - (double)c.off_max_y) / p->zoom - 2.0;
        double hi = ((double)c.row_end + reach - (double)c.off_min_y) / p->zoom + 2.0;
        iy0 = lo > 0 ? (int)lo : 0;
        iy1 = hi < (double)p->in_h ? (int)hi : (int)p->in_h;
        if (iy1 < iy0) iy1 = iy0;
    }
    const size_t npix_in = (size_t)(iy1 - iy0) * p->in_w;
    const uint32_t lanes32 = (p->n_samples + 31) / 32;
    const size_t band_pix = (size_t)(c.row_end - c.row_begin) * p->out_w;
    // Tiled rasteriser (coverage masks in shared memory, fg_gw_tile.cuh) whenever a tile's input footprint is
    // bounded; fg_params.path == FG_PATH_DIRECT (tests) or an extreme zoom-out / radius selects the global-mask kernels.
    const double rmax = (double)p->rm * p->zoom;
    const double foot_rows = ((double)FG_GT_H + 2.0 + ((double)c.off_max_y - (double)c.off_min_y) + 2.0 * rmax) / p->zoom + 8.0;
    bool tiled = p->path != FG_PATH_DIRECT && foot_rows <= (double)FG_GT_MAXROWS && rmax < 4096.0;
    if (tiled && p->path == FG_PATH_AUTO) {
        // Which rasteriser is faster (measured on a B200, C3 and the benchmarks/ sweep): the global-mask splat
        // costs max(1.55 ps of instructions, 4.1 ps per covered pixel: one L2 atomic each) per (grain, sample)
        // and spreads the grains evenly over the whole GPU; the tile kernel costs (1.45 + 0.25 per covered
        // pixel) ps per pair, processes the margin grains of every tile again and keeps only as many SMs busy
        // as there are tiles.  Small disks (few covered pixels per pair) or few tiles -> global mask.
        const double r_out = (double)(p->radius_mean < p->rm ? p->radius_mean : p->rm) * p->zoom;
        const double cover = 3.141592653589793 * r_out * r_out;
        const double mx = std::fmax(std::fabs((double)c.off_min_x), std::fabs((double)c.off_max_x)) + rmax + 1.0;
        const double my = std::fmax(std::fabs((double)c.off_min_y), std::fabs((double)c.off_max_y)) + rmax + 1.0;
        const double halo = ((double)FG_GT_W + 2.0 * mx) * ((double)FG_GT_H + 2.0 * my) / ((double)FG_GT_W * FG_GT_H);
        const double n_tiles = std::ceil((double)p->out_w / FG_GT_W) * std::ceil((double)(c.row_end - c.row_begin) / FG_GT_H);
        // one CTA per SM: with fewer than ~3 waves of tiles the slowest tile (bright content holds many times
        // the average number of grains) sets the time, while the global splat is balanced by construction
        const double util = std::fmin(1.0, n_tiles / (3.0 * (double)ctx->sm_count));
        tiled = std::fmax(1.55, 4.1 * cover) * util > (1.45 + 0.25 * cover) * halo;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
    uint64_t* d_total = (uint64_t*)ctx->misc.p;
    uint64_t total = 0;
    bool ev4 = false, bits_ready = false;
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream)); // ev[6]..ev[4]: grain generation, ev[4]..ev[5]: rasterisation
    if (npix_in > 0) {
        if ((rc = ensure(ctx, ctx->counts, npix_in * sizeof(uint32_t)))) return rc;
        if ((rc = ensure(ctx, ctx->scan_out, npix_in * sizeof(uint64_t)))) return rc;
        if ((rc = ensure(ctx, ctx->gw_states, npix_in * sizeof(ulonglong4)))) return rc;
        const unsigned blocks = (unsigned)((npix_in + 255) / 256);
        k_gw_count<<<blocks, 256, 0, ctx->stream>>>(d_lambda, iy0, iy1, (uint32_t*)ctx->counts.p, (ulonglong4*)ctx->gw_states.p, c);
        FG_CUDA(ctx, cudaGetLastError());
        cub::TransformInputIterator<uint64_t, ToU64, const uint32_t*> it((const uint32_t*)ctx->counts.p, ToU64());
        size_t tmp_bytes = 0;
        FG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, it, (uint64_t*)ctx->scan_out.p, npix_in, ctx->stream));
        if ((rc = ensure(ctx, ctx->scan_tmp, tmp_bytes))) return rc;
        FG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp_bytes, it, (uint64_t*)ctx->scan_out.p, npix_in, ctx->stream));
        k_total<<<1, 32, 0, ctx->stream>>>((const uint32_t*)ctx->counts.p, (const uint64_t*)ctx->scan_out.p, npix_in, d_total);
        FG_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, wait_stream(ctx));
        total = ctx->h_pin[0];
        ctx->stats.launches += 4;
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
        if (total >> 32) tiled = false; // the tile kernel indexes a tile's grains with 32 bits
        if (total > 0) {
            if ((rc = ensure(ctx, ctx->grains, total * sizeof(GrainRec)))) return rc;
            k_gw_fill<<<blocks, 256, 0, ctx->stream>>>(iy0, iy1, (const uint32_t*)ctx->counts.p, (const ulonglong4*)ctx->gw_states.p,
                                                       (const uint64_t*)ctx->scan_out.p, (GrainRec*)ctx->grains.p, c);
            FG_CUDA(ctx, cudaGetLastError());
            ctx->stats.launches += 1;
            FG_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
            ev4 = true;
            if (!tiled) {
                if ((rc = ensure(ctx, ctx->bits, band_pix * lanes32 * sizeof(uint32_t)))) return rc;
                FG_CUDA(ctx, cudaMemsetAsync(ctx->bits.p, 0, band_pix * lanes32 * sizeof(uint32_t), ctx->stream));
                bits_ready = true;
                uint64_t want_blocks = (total + 255) / 256;
                const uint64_t max_blocks = (uint64_t)ctx->sm_count * 64;
                unsigned sblocks = (unsigned)(want_blocks < max_blocks ? want_blocks : max_blocks);
                const bool idx32 = (size_t)p->out_w * p->out_h * lanes32 < ((size_t)1 << 32); // word indices relative to row 0 fit 32 bits
                static const bool sparse_env = !(std::getenv("FG_B200_GW_SPARSE") && std::atoi(std::getenv("FG_B200_GW_SPARSE")) == 0); // experiments
                const bool sparse = sparse_env && 2.0 * rmax < 0.6 && p->out_w < (1u << 21) && p->out_h < (1u << 21);                          // at most ~1/3 of the boxes are non-empty
#define FG_SPLAT(IDX, SP)                                                                                                    \
    k_gw_splat<IDX, SP><<<sblocks, 256, 0, ctx->stream>>>((const GrainRec*)ctx->grains.p, d_total, (const float2*)d_offsets, \
                                                          (uint32_t*)ctx->bits.p, lanes32, c)
                if (idx32) { if (sparse) FG_SPLAT(uint32_t, true); else FG_SPLAT(uint32_t, false); }
                else { if (sparse) FG_SPLAT(size_t, true); else FG_SPLAT(size_t, false); }
#undef FG_SPLAT
                FG_CUDA(ctx, cudaGetLastError());
                ctx->stats.launches += 1;
            }
        }
    } else {
        FG_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint64_t), ctx->stream));
    }
    if (!ev4) FG_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    ctx->fb_count_host() = 0; // fg_get_stats: strip_ms = rasterisation, table_ms = grain generation of the last plane
    ctx->strip_launches = 1;
    ctx->fb_pending = true;
    if (tiled) {
        const int tiles_x = (int)((p->out_w + FG_GT_W - 1) / FG_GT_W), tiles_y = (c.row_end - c.row_begin + FG_GT_H - 1) / FG_GT_H;
        if (tiles_x > 0 && tiles_y > 0) {
            ctx->eval_kernel = "k_gw_tile";
            k_gw_tile<<<(unsigned)tiles_x * (unsigned)tiles_y, FG_GT_THREADS, sizeof(GwTileSmem), ctx->stream>>>(
                (const GrainRec*)ctx->grains.p, (const uint64_t*)ctx->scan_out.p, npix_in, d_total, iy0, iy1, (const float2*)d_offsets,
                d_out, tiles_x, c);
            FG_CUDA(ctx, cudaGetLastError());
            ctx->stats.launches += 1;
        }
        FG_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
        return FG_OK;
    }
    if (!bits_ready) { // no grains at all: an empty mask
        if ((rc = ensure(ctx, ctx->bits, band_pix * lanes32 * sizeof(uint32_t)))) return rc;
        FG_CUDA(ctx, cudaMemsetAsync(ctx->bits.p, 0, band_pix * lanes32 * sizeof(uint32_t), ctx->stream));
    }
    ctx->eval_kernel = "k_gw_splat + k_gw_reduce";
    k_gw_reduce<<<(unsigned)((band_pix + 255) / 256), 256, 0, ctx->stream>>>((const uint32_t*)ctx->bits.p, lanes32, d_out, c);
    FG_CUDA(ctx, cudaGetLastError());
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->stats.launches += 1;
    return FG_OK;
}

int render_planes_device_locked(fg_ctx* ctx, const fg_params* p, const RenderConsts& c_in, int algo, int n_planes,
                                const float* d_lambda, const float* d_offsets, float* d_out) {
    int rc = init_tables(ctx);
    if (rc) return rc;
    if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    RenderConsts c = c_in;
    if (ctx->abort_sent) { // the previous render on this context was stopped inside a launch: lower the word again
        FG_CUDA(ctx, cudaStreamSynchronize(ctx->abort_stream));
        FG_CUDA(ctx, cudaMemsetAsync(ctx->d_abort, 0, sizeof(int), ctx->stream));
        ctx->abort_sent = false;
    }
    c.abort = cancel_armed(ctx) ? ctx->d_abort : nullptr; // no flag armed: the kernels do not even look
    if (algo == FG_ALGO_PIXEL) return pixelwise_device(ctx, p, c, n_planes, d_lambda, d_offsets, d_out);
    const size_t in_stride = (size_t)p->in_w * p->in_h, out_stride = (size_t)p->out_w * p->out_h;
    for (int pl = 0; pl < n_planes; ++pl) {
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
        rc = grainwise_device_plane(ctx, p, c, d_lambda + in_stride * pl, d_offsets, d_out + out_stride * pl);
        if (rc) return rc;
    }
    return FG_OK;
}

// Input rows [r0, r1) a render of the band [c.row_begin, c.row_end) can read (everything for a full render).
// Pixel-wise: the cells within rm of the band's sample points map to input rows floor(j * delta), clamped
// (src/pixelwise.rs:55-73); the cell table adds two cells and a row of slack on each side.  Grain-wise: the
// rows whose grains can reach the band (the same expression as grainwise_device_plane).
void input_rows_of_band(const fg_params* p, const RenderConsts& c, int algo, int& r0, int& r1) {
    r0 = 0; r1 = (int)p->in_h;
    if (c.row_begin <= 0 && c.row_end >= (int)p->out_h) return;
    double lo, hi;
    if (algo == FG_ALGO_PIXEL) {
        const double slack = 2.0 * (double)p->delta + 4.0;
        lo = ((double)c.row_begin + 0.5) / p->zoom - (double)c.off_max_y - (double)p->rm - slack;
        hi = ((double)c.row_end - 0.5) / p->zoom - (double)c.off_min_y + (double)p->rm + slack + 1.0;
    } else {
        const double reach = (double)p->rm * p->zoom + 2.0;
        lo = ((double)c.row_begin - reach - (double)c.off_max_y) / p->zoom - 3.0;
        hi = ((double)c.row_end + reach - (double)c.off_min_y) / p->zoom + 3.0;
    }
    if (lo > 0.0) r0 = lo < (double)p->in_h ? (int)lo : (int)p->in_h - 1; // clamped lookups read the edge rows
    if (hi < (double)p->in_h) r1 = hi > 1.0 ? (int)hi : 1;
    if (r1 <= r0) { r0 = 0; r1 = (int)p->in_h; }
}

// Output planes that form ONE contiguous block of page-locked host memory mapped into the device's address
// space (cudaHostAlloc / cudaHostRegister; unified addressing): the kernels can store their results
// straight into it, so the device->host transfer of the image happens during the render instead of after
// it.  Returns the device-side pointer of the block, or nullptr (pageable or scattered planes: staged copy).
float* mapped_host_planes(float* const* out, int n_planes, size_t out_elems) {
    static const bool enabled = !(std::getenv("FG_B200_ZEROCOPY") && std::atoi(std::getenv("FG_B200_ZEROCOPY")) == 0);
    if (!enabled) return nullptr;
    for (int pl = 1; pl < n_planes; ++pl)
        if (out[pl] != out[0] + out_elems * pl) return nullptr;
    cudaPointerAttributes first{}, last{};
    if (cudaPointerGetAttributes(&first, out[0]) != cudaSuccess ||
        cudaPointerGetAttributes(&last, (const char*)out[0] + out_elems * n_planes * sizeof(float) - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (first.type != cudaMemoryTypeHost || last.type != cudaMemoryTypeHost || !first.devicePointer || !last.devicePointer) return nullptr;
    if ((const char*)last.devicePointer - (const char*)first.devicePointer != (ptrdiff_t)(out_elems * n_planes * sizeof(float) - 1)) return nullptr;
    return (float*)first.devicePointer;
}

int render_planes_host(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda,
                       const float* offsets, float* const* out, const volatile int* cancel = nullptr) {
    if (!ctx) return FG_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedCallCancel call_cancel(ctx, cancel);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    ctx->stats = fg_stats{};
    ctx->fb_pending = false;
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (algo != FG_ALGO_PIXEL && algo != FG_ALGO_GRAIN) return set_err(ctx, FG_ERR_INVALID, "algo must be FG_ALGO_PIXEL or FG_ALGO_GRAIN");
    if (n_planes < 1 || n_planes > 16) return set_err(ctx, FG_ERR_INVALID, "n_planes must be in 1..16");
    if (!lambda || !offsets || !out) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    for (int pl = 0; pl < n_planes; ++pl)
        if (!lambda[pl] || !out[pl]) return set_err(ctx, FG_ERR_INVALID, "NULL plane pointer");
    if ((rc = check_offsets(ctx, p, offsets))) return rc;
    RenderConsts c = make_consts(p, offsets);
    const size_t in_elems = (size_t)p->in_w * p->in_h, out_elems = (size_t)p->out_w * p->out_h;
    if ((rc = ensure(ctx, ctx->lambda, in_elems * n_planes * sizeof(float)))) return rc;
    float* const mapped = mapped_host_planes(out, n_planes, out_elems);
    if (!mapped && (rc = ensure(ctx, ctx->out, out_elems * n_planes * sizeof(float)))) return rc;
    float* const d_dst = mapped ? mapped : (float*)ctx->out.p;
    if ((rc = ensure(ctx, ctx->offsets, (size_t)p->n_samples * 2 * sizeof(float)))) return rc;
    cudaStream_t s = ctx->stream;
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
    // a row band uploads only the input rows it can read
    int in_r0, in_r1;
    input_rows_of_band(p, c, algo, in_r0, in_r1);
    static const bool poison = std::getenv("FG_B200_POISON") && std::atoi(std::getenv("FG_B200_POISON")) != 0; // tests: NaN outside the uploaded rows
    if (poison) FG_CUDA(ctx, cudaMemsetAsync(ctx->lambda.p, 0xFF, in_elems * n_planes * sizeof(float), s));
    const size_t up_off = (size_t)in_r0 * p->in_w, up_elems = (size_t)(in_r1 - in_r0) * p->in_w;
    // A whole-frame pixel-wise render of a few MB or more leaves the upload to the pipeline, which overlaps it with the
    // thresholds and the first-draw bitmap row chunk by row chunk (fg_pixel_host.cuh); everything else uploads here.
    static const bool chunk_env = !(std::getenv("FG_B200_CHUNKED_UPLOAD") && std::atoi(std::getenv("FG_B200_CHUNKED_UPLOAD")) == 0);
    const bool defer = chunk_env && algo == FG_ALGO_PIXEL && in_r0 == 0 && in_r1 == (int)p->in_h && p->in_h >= 64 && ctx->copy_stream &&
                       in_elems * n_planes * sizeof(float) >= ((size_t)4 << 20) && !ctx->tcache.enabled;
    ctx->up.pending = false;
    if (defer) {
        ctx->up.pending = true;
        ctx->up.host = lambda; ctx->up.n_planes = n_planes; ctx->up.in_w = p->in_w; ctx->up.in_h = p->in_h; ctx->up.dev = (float*)ctx->lambda.p;
    } else {
        for (int pl = 0; pl < n_planes; ++pl)
            FG_CUDA(ctx, cudaMemcpyAsync((float*)ctx->lambda.p + in_elems * pl + up_off, lambda[pl] + up_off, up_elems * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->offsets.p, offsets, (size_t)p->n_samples * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
    static const bool slice_env = !(std::getenv("FG_B200_SLICED_OUTPUT") && std::atoi(std::getenv("FG_B200_SLICED_OUTPUT")) == 0);
    ctx->outp.want = slice_env && !mapped && algo == FG_ALGO_PIXEL && ctx->copy_stream && !cancel_armed(ctx) &&
                     out_elems * n_planes * sizeof(float) >= ((size_t)8 << 20);
    ctx->outp.n = 0;
    rc = render_planes_device_locked(ctx, p, c, algo, n_planes, (const float*)ctx->lambda.p, (const float*)ctx->offsets.p, d_dst);
    ctx->outp.want = false;
    ctx->up.pending = false; // (consumed by now; the caller's plane array is not referenced past this call)
    if (rc) { cudaStreamSynchronize(s); if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream); return rc; }
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[2], s));
    if (cancel_armed(ctx)) { // a copy into pageable memory blocks the host until the kernels are done: watch the flag first
        FG_CUDA(ctx, wait_stream(ctx));
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    }
    const size_t band_off = (size_t)c.row_begin * p->out_w, band_elems = (size_t)(c.row_end - c.row_begin) * p->out_w;
    if (!mapped && ctx->outp.n > 1) { // the evaluation ran in row slices: copy slice e while slice e + 1 is being evaluated
        int y0 = c.row_begin;
        for (int e = 0; e < ctx->outp.n; ++e) {
            const int y1 = ctx->outp.row_end[e];
            FG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->outp.ev[e], 0));
            const size_t o = (size_t)y0 * p->out_w, n = (size_t)(y1 - y0) * p->out_w;
            for (int pl = 0; pl < n_planes && n; ++pl)
                FG_CUDA(ctx, cudaMemcpyAsync(out[pl] + o, (float*)ctx->out.p + out_elems * pl + o, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_stream));
            y0 = y1;
        }
        FG_CUDA(ctx, cudaEventRecord(ctx->up_ev[4], ctx->copy_stream));
        FG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->up_ev[4], 0)); // the wait below covers the copies
        ctx->outp.n = 0;
    } else
    for (int pl = 0; pl < n_planes && !mapped; ++pl) // mapped: the kernels' own stores were the transfer
        FG_CUDA(ctx, cudaMemcpyAsync(out[pl] + band_off, (float*)ctx->out.p + out_elems * pl + band_off, band_elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
    FG_CUDA(ctx, wait_stream(ctx));
    cudaEventElapsedTime(&ctx->stats.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->stats.d2h_ms, ctx->ev[2], ctx->ev[3]);
    ctx->stats.h2d_bytes = (up_elems * n_planes + (size_t)p->n_samples * 2) * sizeof(float);
    ctx->stats.d2h_bytes = band_elems * n_planes * sizeof(float);
    if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    return FG_OK;
}


// ---- progressive refinement (fg_refine_planes) ---------------------------------------------------------------------
// out = (count_a + count_b) * (1 / (n_a + n_b)) with count = rint(mean * n): exact for the sample counts a render can
// hold (f32 integers), so the merged image equals one render of all n_a + n_b samples bit for bit.
__global__ void __launch_bounds__(256) k_refine_merge(float* __restrict__ acc, const float* __restrict__ part, size_t out_elems, size_t band_off,
                                                       size_t band_elems, int n_planes, float n_acc, float n_part, float inv_total) {
    const size_t n = band_elems * (size_t)n_planes;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (size_t)gridDim.x * 256) {
        const size_t pl = t / band_elems, idx = pl * out_elems + band_off + (t - pl * band_elems);
        const float a = rintf(__fmul_rn(acc[idx], n_acc)), b = rintf(__fmul_rn(part[idx], n_part));
        acc[idx] = __fmul_rn(__fadd_rn(a, b), inv_total);
    }
}

int refine_planes_host(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda, const float* offsets,
                       uint32_t k_begin, uint32_t k_end, float* const* out, const volatile int* cancel) {
    if (!ctx) return FG_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedCallCancel call_cancel(ctx, cancel);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    ctx->stats = fg_stats{};
    ctx->fb_pending = false;
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (algo != FG_ALGO_PIXEL && algo != FG_ALGO_GRAIN) return set_err(ctx, FG_ERR_INVALID, "algo must be FG_ALGO_PIXEL or FG_ALGO_GRAIN");
    if (n_planes < 1 || n_planes > 16) return set_err(ctx, FG_ERR_INVALID, "n_planes must be in 1..16");
    if (!lambda || !offsets || !out) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    for (int pl = 0; pl < n_planes; ++pl)
        if (!lambda[pl] || !out[pl]) return set_err(ctx, FG_ERR_INVALID, "NULL plane pointer");
    if (!(k_begin < k_end && k_end <= p->n_samples)) return set_err(ctx, FG_ERR_INVALID, "sample slice must satisfy k_begin < k_end <= n_samples");
    if ((rc = check_offsets(ctx, p, offsets))) return rc;
    // the extents of ALL n_samples offsets size the cell rectangle, so every slice of a refinement asks for the same table
    RenderConsts c = make_consts(p, offsets);
    fg_params pp = *p;
    pp.n_samples = k_end - k_begin;
    c.n = pp.n_samples;
    c.inv_samples = 1.0f / (float)pp.n_samples;
    auto& pg = ctx->prog;
    if (k_begin > 0 && !(pg.valid && pg.k == k_begin && pg.algo == algo && pg.n_planes == (uint32_t)n_planes && pg.out_w == p->out_w &&
                         pg.out_h == p->out_h && pg.row_begin == (uint32_t)c.row_begin && pg.row_end == (uint32_t)c.row_end))
        return set_err(ctx, FG_ERR_INVALID, "fg_refine_planes: k_begin > 0 continues the previous refinement of this context, which ended elsewhere");
    pg.valid = false;
    const size_t in_elems = (size_t)p->in_w * p->in_h, out_elems = (size_t)p->out_w * p->out_h;
    if ((rc = ensure(ctx, ctx->lambda, in_elems * n_planes * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->acc, out_elems * n_planes * sizeof(float)))) return rc;
    if (k_begin && (rc = ensure(ctx, ctx->part, out_elems * n_planes * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->offsets, (size_t)pp.n_samples * 2 * sizeof(float)))) return rc;
    cudaStream_t s = ctx->stream;
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[0], s));
    int in_r0, in_r1;
    input_rows_of_band(p, c, algo, in_r0, in_r1);
    const size_t up_off = (size_t)in_r0 * p->in_w, up_elems = (size_t)(in_r1 - in_r0) * p->in_w;
    for (int pl = 0; pl < n_planes; ++pl)
        FG_CUDA(ctx, cudaMemcpyAsync((float*)ctx->lambda.p + in_elems * pl + up_off, lambda[pl] + up_off, up_elems * sizeof(float), cudaMemcpyHostToDevice, s));
    FG_CUDA(ctx, cudaMemcpyAsync(ctx->offsets.p, offsets + 2 * (size_t)k_begin, (size_t)pp.n_samples * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[1], s));
    float* const d_dst = k_begin ? (float*)ctx->part.p : (float*)ctx->acc.p;
    rc = render_planes_device_locked(ctx, &pp, c, algo, n_planes, (const float*)ctx->lambda.p, (const float*)ctx->offsets.p, d_dst);
    if (rc) { cudaStreamSynchronize(s); return rc; }
    const size_t band_off = (size_t)c.row_begin * p->out_w, band_elems = (size_t)(c.row_end - c.row_begin) * p->out_w;
    if (k_begin) {
        const unsigned blocks = (unsigned)std::min<size_t>((band_elems * n_planes + 255) / 256, (size_t)ctx->sm_count * 16);
        k_refine_merge<<<blocks, 256, 0, s>>>((float*)ctx->acc.p, (const float*)ctx->part.p, out_elems, band_off, band_elems, n_planes,
                                              (float)k_begin, (float)pp.n_samples, 1.0f / (float)k_end);
        FG_CUDA(ctx, cudaGetLastError());
        ctx->stats.launches += 1;
    }
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[2], s));
    if (cancel_armed(ctx)) {
        FG_CUDA(ctx, wait_stream(ctx));
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    }
    for (int pl = 0; pl < n_planes; ++pl)
        FG_CUDA(ctx, cudaMemcpyAsync(out[pl] + band_off, (float*)ctx->acc.p + out_elems * pl + band_off, band_elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[3], s));
    FG_CUDA(ctx, wait_stream(ctx));
    cudaEventElapsedTime(&ctx->stats.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->stats.d2h_ms, ctx->ev[2], ctx->ev[3]);
    ctx->stats.h2d_bytes = (up_elems * n_planes + (size_t)pp.n_samples * 2) * sizeof(float);
    ctx->stats.d2h_bytes = band_elems * n_planes * sizeof(float);
    if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    pg.valid = true; pg.k = k_end; pg.algo = algo; pg.n_planes = (uint32_t)n_planes; pg.out_w = p->out_w; pg.out_h = p->out_h;
    pg.row_begin = (uint32_t)c.row_begin; pg.row_end = (uint32_t)c.row_end;
    return FG_OK;
}


// ---- single-process multi-device rendering (fg_context_create_multi) --------------------------------------------
// The path shards into independent output row bands: device g renders the g-th band of the requested
// rows with the margin cells regenerated locally -- no halo exchange, no collective.  One host thread per device drives
// that device's own context; the bands meet in the caller's buffer (host entry points: every device copies its band
// there itself, over its own PCIe link) or in device 0's image (device entry point: peer stores over NVLink).
void band_of(int g, int n, int r0, int r1, int& b0, int& b1) {
    const int rows = r1 - r0, base = rows / n, rem = rows % n;
    b0 = r0 + g * base + std::min(g, rem);
    b1 = b0 + base + (g < rem ? 1 : 0);
}

void request_rows(const fg_params* p, int& r0, int& r1) {
    if (p->row_begin == 0 && p->row_end == 0) { r0 = 0; r1 = (int)p->out_h; }
    else { r0 = (int)p->row_begin; r1 = (int)p->row_end; }
}

// run f(g) for every device of the context, one thread each (the calling thread takes device 0); first failure wins
template <class F>
int multi_run(fg_ctx* ctx, F&& f) {
    const int n = (int)ctx->subs.size();
    std::vector<int> rcs((size_t)n, FG_OK);
    try {
        std::vector<std::thread> th;
        for (int g = 1; g < n; ++g) th.emplace_back([&rcs, &f, g] { rcs[(size_t)g] = f(g); });
        rcs[0] = f(0);
        for (auto& t : th) t.join();
    } catch (const std::exception& e) {
        return set_err(ctx, FG_ERR_CUDA, std::string("multi-device render: ") + e.what());
    }
    fg_stats agg{};
    for (int g = 0; g < n; ++g) {
        fg_stats st{};
        fg_get_stats(ctx->subs[(size_t)g], &st);
        agg.kernel_ms = std::max(agg.kernel_ms, st.kernel_ms);
        agg.h2d_ms = std::max(agg.h2d_ms, st.h2d_ms);
        agg.d2h_ms = std::max(agg.d2h_ms, st.d2h_ms);
        agg.strip_ms = std::max(agg.strip_ms, st.strip_ms);
        agg.table_ms = std::max(agg.table_ms, st.table_ms);
        agg.strip_launches = std::max(agg.strip_launches, st.strip_launches);
        agg.launches += st.launches;
        agg.tiles_total += st.tiles_total;
        agg.tiles_fallback += st.tiles_fallback;
        agg.h2d_bytes += st.h2d_bytes;
        agg.d2h_bytes += st.d2h_bytes;
    }
    ctx->stats = agg;
    for (int g = 0; g < n; ++g)
        if (rcs[(size_t)g]) {
            ctx->err = "device " + std::to_string(ctx->subs[(size_t)g]->device) + ": " + ctx->subs[(size_t)g]->err;
            return rcs[(size_t)g];
        }
    return FG_OK;
}

int multi_render_planes_host(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda,
                             const float* offsets, float* const* out, const volatile int* cancel) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    int rc = validate(ctx, p);
    if (rc) return rc;
    int r0, r1;
    request_rows(p, r0, r1);
    const int n = (int)ctx->subs.size();
    return multi_run(ctx, [&](int g) -> int {
        fg_params pb = *p;
        int b0, b1;
        band_of(g, n, r0, r1, b0, b1);
        if (b0 >= b1) return FG_OK; // more devices than rows
        pb.row_begin = (uint32_t)b0;
        pb.row_end = (uint32_t)b1;
        return render_planes_host(ctx->subs[(size_t)g], &pb, algo, n_planes, lambda, offsets, out,
                                  cancel ? cancel : ctx->cancel);
    });
}

int multi_render_rgb8_host(fg_ctx* ctx, const fg_params* p, int algo, int color_mode, const uint8_t* rgb_in,
                           const float* offsets, uint8_t* rgb_out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    int rc = validate(ctx, p);
    if (rc) return rc;
    int r0, r1;
    request_rows(p, r0, r1);
    const int n = (int)ctx->subs.size();
    return multi_run(ctx, [&](int g) -> int {
        fg_params pb = *p;
        int b0, b1;
        band_of(g, n, r0, r1, b0, b1);
        if (b0 >= b1) return FG_OK;
        pb.row_begin = (uint32_t)b0;
        pb.row_end = (uint32_t)b1;
        return fg_render_rgb8(ctx->subs[(size_t)g], &pb, algo, color_mode, rgb_in, offsets, rgb_out);
    });
}

// Device-resident planes: d_lambda / d_offsets / d_out live on device 0 (subs[0]).  Device g > 0 pulls the lambda rows its
// band reads over NVLink (cudaMemcpyPeerAsync into its own pool) and its kernels store their band rows straight into
// device 0's image through the peer mapping enabled at context creation (a staged peer copy where that is unavailable).
// subs[0]'s stream waits for every other device, so "the image is complete" is an ordinary stream-order fact for the caller.
int multi_render_planes_device(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* d_lambda,
                               const float* d_offsets, float* d_out, int stream_sync) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (algo != FG_ALGO_PIXEL && algo != FG_ALGO_GRAIN) return set_err(ctx, FG_ERR_INVALID, "algo must be FG_ALGO_PIXEL or FG_ALGO_GRAIN");
    if (n_planes < 1 || n_planes > 16) return set_err(ctx, FG_ERR_INVALID, "n_planes must be in 1..16");
    if (!d_lambda || !d_offsets || !d_out) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    fg_ctx* c0 = ctx->subs[0];
    std::vector<float> off((size_t)p->n_samples * 2);
    {
        ScopedDevice dev(c0->device);
        FG_CUDA(ctx, cudaMemcpyAsync(off.data(), d_offsets, off.size() * sizeof(float), cudaMemcpyDeviceToHost, c0->stream));
        FG_CUDA(ctx, cudaEventRecord(c0->ev[7], c0->stream)); // the inputs are ready on device 0 from here on
        FG_CUDA(ctx, cudaStreamSynchronize(c0->stream));
    }
    if ((rc = check_offsets(ctx, p, off.data()))) return rc;
    int r0, r1;
    request_rows(p, r0, r1);
    const int n = (int)ctx->subs.size();
    const size_t in_elems = (size_t)p->in_w * p->in_h, out_elems = (size_t)p->out_w * p->out_h;
    rc = multi_run(ctx, [&](int g) -> int {
        fg_ctx* sc = ctx->subs[(size_t)g];
        std::lock_guard<std::mutex> sub_lock(sc->mu);
        ScopedDevice dev(sc->device);
        sc->err.clear();
        sc->stats = fg_stats{};
        sc->fb_pending = false;
        fg_params pb = *p;
        int b0, b1;
        band_of(g, n, r0, r1, b0, b1);
        if (b0 >= b1) return FG_OK;
        pb.row_begin = (uint32_t)b0;
        pb.row_end = (uint32_t)b1;
        RenderConsts c = make_consts(&pb, off.data());
        cudaStream_t s = sc->stream;
        const float* lam = d_lambda;
        const float* offs = d_offsets;
        float* dst = d_out;
        int rcg;
        if (g > 0) {
            FG_CUDA(sc, cudaStreamWaitEvent(s, c0->ev[7], 0));
            if ((rcg = ensure(sc, sc->lambda, in_elems * n_planes * sizeof(float)))) return rcg;
            if ((rcg = ensure(sc, sc->offsets, off.size() * sizeof(float)))) return rcg;
            int in_r0, in_r1;
            input_rows_of_band(&pb, c, algo, in_r0, in_r1);
            const size_t up_off = (size_t)in_r0 * p->in_w, up_elems = (size_t)(in_r1 - in_r0) * p->in_w;
            for (int pl = 0; pl < n_planes; ++pl)
                FG_CUDA(sc, cudaMemcpyPeerAsync((float*)sc->lambda.p + in_elems * pl + up_off, sc->device, d_lambda + in_elems * pl + up_off,
                                                c0->device, up_elems * sizeof(float), s));
            FG_CUDA(sc, cudaMemcpyAsync(sc->offsets.p, off.data(), off.size() * sizeof(float), cudaMemcpyHostToDevice, s));
            lam = (const float*)sc->lambda.p;
            offs = (const float*)sc->offsets.p;
            if (!ctx->peer[(size_t)g]) { // no peer mapping: render locally, then a peer copy of the band
                if ((rcg = ensure(sc, sc->out, out_elems * n_planes * sizeof(float)))) return rcg;
                dst = (float*)sc->out.p;
            }
        }
        FG_CUDA(sc, cudaEventRecord(sc->ev[1], s));
        rcg = render_planes_device_locked(sc, &pb, c, algo, n_planes, lam, offs, dst);
        if (rcg) { cudaStreamSynchronize(s); return rcg; }
        FG_CUDA(sc, cudaEventRecord(sc->ev[2], s));
        if (g > 0 && !ctx->peer[(size_t)g]) {
            const size_t band_off = (size_t)b0 * p->out_w, band_elems = (size_t)(b1 - b0) * p->out_w;
            for (int pl = 0; pl < n_planes; ++pl)
                FG_CUDA(sc, cudaMemcpyPeerAsync(d_out + out_elems * pl + band_off, c0->device, dst + out_elems * pl + band_off, sc->device,
                                                band_elems * sizeof(float), s));
        }
        FG_CUDA(sc, cudaEventRecord(sc->ev[8], s));
        return FG_OK;
    });
    if (rc) return rc;
    {
        ScopedDevice dev(c0->device);
        for (int g = 1; g < n; ++g) FG_CUDA(ctx, cudaStreamWaitEvent(c0->stream, ctx->subs[(size_t)g]->ev[8], 0));
        if (stream_sync) {
            FG_CUDA(ctx, cudaStreamSynchronize(c0->stream));
            for (int g = 0; g < n; ++g) {
                float ms = 0.0f;
                if (cudaEventElapsedTime(&ms, ctx->subs[(size_t)g]->ev[1], ctx->subs[(size_t)g]->ev[2]) == cudaSuccess)
                    ctx->stats.kernel_ms = std::max(ctx->stats.kernel_ms, ms);
                else cudaGetLastError();
            }
        }
    }
    return FG_OK;
}

} // namespace

// ------------------------------------------------------------------------------- ABI
extern "C" {

int fg_abi_version(void) { return FG_ABI_VERSION; }

int fg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* fg_error_string(int code) {
    switch (code) {
    case FG_OK: return "ok";
    case FG_ERR_INVALID: return "invalid argument";
    case FG_ERR_OOM: return "out of device memory";
    case FG_ERR_CUDA_STICKY: return "unrecoverable CUDA error (drop the context)";
    case FG_ERR_NO_DEVICE: return "no usable CUDA device";
    case FG_ERR_CANCELLED: return "cancelled";
    case FG_ERR_CUDA: return "CUDA error";
    default: return "unknown error";
    }
}

int fg_context_create(fg_ctx** out, int device) {
    if (!out) return FG_ERR_INVALID;
    *out = nullptr;
    int n = fg_device_count();
    if (n <= 0 || device < 0 || device >= n) return FG_ERR_NO_DEVICE;
    fg_ctx* ctx = new (std::nothrow) fg_ctx();
    if (!ctx) return FG_ERR_OOM;
    ctx->device = device;
    ScopedDevice dev(device);
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        cudaGetLastError();
        delete ctx;
        return FG_ERR_NO_DEVICE; // kernels are sm_100a only
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    {
        // cell-table budget: at most half of the device memory, 48 GiB by default
        ctx->table_max = std::min<size_t>(ctx->table_max, prop.totalGlobalMem / 2);
        if (const char* v = std::getenv("FG_B200_TABLE_MAX_BYTES")) {
            const unsigned long long b = std::strtoull(v, nullptr, 10);
            if (b >= (1ULL << 20)) ctx->table_max = (size_t)b;
        }
        if (const char* v = std::getenv("FG_B200_TABLE_SLACK_SIGMA")) ctx->table_slack_sigma = std::atof(v);
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return FG_ERR_CUDA_STICKY;
    }
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    if (cudaHostAlloc((void**)&ctx->h_pin, 16 * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fg_context_destroy(ctx);
        return FG_ERR_OOM;
    }
    std::memset(ctx->h_pin, 0, 16 * sizeof(uint64_t));
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ctx->copy_stream = nullptr; }
    for (auto& e : ctx->up_ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); e = nullptr; }
    for (auto& e : ctx->outp.ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); e = nullptr; }
    // in-launch cancel (fg_ctx.cuh: wait_stream); without these the flag is still honoured between the stages
    if (cudaStreamCreateWithFlags(&ctx->abort_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_wait, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_abort, sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_abort, 0, sizeof(int)) != cudaSuccess ||
        cudaHostAlloc((void**)&ctx->h_one, sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        if (ctx->d_abort) cudaFree(ctx->d_abort);
        ctx->d_abort = nullptr;
    }
    if (ctx->h_one) *ctx->h_one = 1;
    int rc = tile_setup(ctx);
    if (!rc && cudaFuncSetAttribute(k_gw_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GwTileSmem)) != cudaSuccess) {
        cudaGetLastError();
        rc = FG_ERR_CUDA_STICKY;
    }
    if (rc) { fg_context_destroy(ctx); return rc; }
    *out = ctx;
    return FG_OK;
}

int fg_context_create_multi(fg_ctx** out, const int* devices, int n_devices) {
    if (!out) return FG_ERR_INVALID;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return FG_ERR_INVALID;
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b]) return FG_ERR_INVALID;
    fg_ctx* ctx = new (std::nothrow) fg_ctx();
    if (!ctx) return FG_ERR_OOM;
    ctx->device = devices[0];
    for (int g = 0; g < n_devices; ++g) {
        fg_ctx* sc = nullptr;
        const int rc = fg_context_create(&sc, devices[g]);
        if (rc) { fg_context_destroy(ctx); return rc; }
        ctx->subs.push_back(sc);
        char ok = 1;
        if (g > 0) { // let device g store into device 0's memory (NVLink / NVSwitch peer mapping)
            ScopedDevice dev(devices[g]);
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[g], devices[0]) != cudaSuccess || !can) ok = 0;
            else {
                const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
            }
            cudaGetLastError();
        }
        ctx->peer.push_back(ok);
    }
    ctx->sm_count = ctx->subs[0]->sm_count;
    *out = ctx;
    return FG_OK;
}

void fg_context_destroy(fg_ctx* ctx) {
    if (!ctx) return;
    if (!ctx->subs.empty()) { // multi-device router: owns nothing but its sub-contexts
        for (fg_ctx* sc : ctx->subs) fg_context_destroy(sc);
        delete ctx;
        return;
    }
    {
        ScopedDevice dev(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        for (DevBuf* b : {&ctx->lambda, &ctx->out, &ctx->offsets, &ctx->bits, &ctx->counts, &ctx->scan_out, &ctx->scan_tmp,
                          &ctx->grains, &ctx->misc, &ctx->tiles, &ctx->thr, &ctx->bitmap, &ctx->rowinfo, &ctx->ptab, &ctx->gtab, &ctx->fbtotal, &ctx->rgb_in, &ctx->rgb_out, &ctx->chroma, &ctx->lut, &ctx->gw_states, &ctx->acc, &ctx->part})
            release(*b);
        for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
        if (ctx->abort_stream) { cudaStreamSynchronize(ctx->abort_stream); cudaStreamDestroy(ctx->abort_stream); }
        if (ctx->ev_wait) cudaEventDestroy(ctx->ev_wait);
        if (ctx->d_abort) cudaFree(ctx->d_abort);
        if (ctx->h_one) cudaFreeHost(ctx->h_one);
        if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
        if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
        for (auto& e : ctx->up_ev) if (e) cudaEventDestroy(e);
        for (auto& e : ctx->outp.ev) if (e) cudaEventDestroy(e);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        cudaGetLastError();
    }
    delete ctx;
}

const char* fg_last_error(const fg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
const char* fg_last_eval_kernel(const fg_ctx* ctx) { return ctx ? (ctx->subs.empty() ? ctx->eval_kernel : ctx->subs[0]->eval_kernel) : ""; }
int fg_context_device_count(const fg_ctx* ctx) { return ctx ? (ctx->subs.empty() ? 1 : (int)ctx->subs.size()) : 0; }
void fg_set_cancel_flag(fg_ctx* ctx, const volatile int* flag) {
    if (!ctx) return;
    ctx->cancel = flag;
    for (fg_ctx* sc : ctx->subs) sc->cancel = flag;
}
void fg_get_stats(const fg_ctx* ctx, fg_stats* out) {
    if (!ctx || !out) return;
    *out = ctx->stats;
    if (!ctx->subs.empty()) return; // aggregated over the devices by the render call
    if (ctx->fb_pending) { // meaningful once the stream is synchronised
        out->tiles_fallback = ctx->fb_count_host();
        out->strip_launches = ctx->strip_launches;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]) == cudaSuccess) out->strip_ms = ms;
        else cudaGetLastError();
        if (cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[4]) == cudaSuccess) out->table_ms = ms;
        else cudaGetLastError();
    }
}
uint64_t fg_context_stream(const fg_ctx* ctx) {
    if (!ctx) return 0;
    return (uint64_t)(uintptr_t)(ctx->subs.empty() ? ctx->stream : ctx->subs[0]->stream); // multi: device 0's stream
}

int fg_context_synchronize(fg_ctx* ctx) {
    if (!ctx) return FG_ERR_INVALID;
    if (!ctx->subs.empty()) {
        for (fg_ctx* sc : ctx->subs) {
            const int rc = fg_context_synchronize(sc);
            if (rc) { ctx->err = sc->err; return rc; }
        }
        return FG_OK;
    }
    ScopedDevice dev(ctx->device);
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FG_OK;
}

int fg_render_pixelwise(fg_ctx* ctx, const fg_params* p, const float* lambda, const float* offsets_input, float* out) {
    const float* lp[1] = {lambda};
    float* op[1] = {out};
    return fg_render_planes(ctx, p, FG_ALGO_PIXEL, 1, lp, offsets_input, op);
}

int fg_render_grainwise(fg_ctx* ctx, const fg_params* p, const float* lambda, const float* offsets, float* out) {
    const float* lp[1] = {lambda};
    float* op[1] = {out};
    return fg_render_planes(ctx, p, FG_ALGO_GRAIN, 1, lp, offsets, op);
}

int fg_render_planes(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda,
                     const float* offsets, float* const* out) {
    if (ctx && !ctx->subs.empty()) return multi_render_planes_host(ctx, p, algo, n_planes, lambda, offsets, out, nullptr);
    return render_planes_host(ctx, p, algo, n_planes, lambda, offsets, out);
}

int fg_refine_planes(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda, const float* offsets,
                     uint32_t k_begin, uint32_t k_end, float* const* out, const volatile int* cancel) {
    // a multi-device context refines on its first device (a preview path: one frame, one GPU)
    fg_ctx* c1 = (ctx && !ctx->subs.empty()) ? ctx->subs[0] : ctx;
    const int rc = refine_planes_host(c1, p, algo, n_planes, lambda, offsets, k_begin, k_end, out, cancel ? cancel : (ctx ? ctx->cancel : nullptr));
    if (c1 != ctx && rc) ctx->err = c1->err;
    return rc;
}

void fg_set_table_cache(fg_ctx* ctx, int enable) {
    if (!ctx) return;
    std::lock_guard<std::mutex> lock(ctx->mu);
    for (fg_ctx* sc : ctx->subs) fg_set_table_cache(sc, enable);
    ctx->tcache.enabled = enable != 0;
    if (!enable) ctx->tcache.valid = false;
}

int fg_render_planes_cancelable(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* const* lambda,
                                const float* offsets, float* const* out, const volatile int* cancel) {
    if (ctx && !ctx->subs.empty()) return multi_render_planes_host(ctx, p, algo, n_planes, lambda, offsets, out, cancel);
    return render_planes_host(ctx, p, algo, n_planes, lambda, offsets, out, cancel);
}

int fg_render_planes_device(fg_ctx* ctx, const fg_params* p, int algo, int n_planes, const float* d_lambda,
                            const float* d_offsets, float* d_out, int stream_sync) {
    if (!ctx) return FG_ERR_INVALID;
    if (!ctx->subs.empty()) return multi_render_planes_device(ctx, p, algo, n_planes, d_lambda, d_offsets, d_out, stream_sync);
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    ctx->stats = fg_stats{};
    ctx->fb_pending = false;
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (algo != FG_ALGO_PIXEL && algo != FG_ALGO_GRAIN) return set_err(ctx, FG_ERR_INVALID, "algo must be FG_ALGO_PIXEL or FG_ALGO_GRAIN");
    if (n_planes < 1 || n_planes > 16) return set_err(ctx, FG_ERR_INVALID, "n_planes must be in 1..16");
    if (!d_lambda || !d_offsets || !d_out) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    // the offset extremes size the tile windows: read the (tiny) offset array back once
    std::vector<float> off((size_t)p->n_samples * 2);
    FG_CUDA(ctx, cudaMemcpyAsync(off.data(), d_offsets, off.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = check_offsets(ctx, p, off.data()))) return rc;
    RenderConsts c = make_consts(p, off.data());
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    rc = render_planes_device_locked(ctx, p, c, algo, n_planes, d_lambda, d_offsets, d_out);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (stream_sync) {
        FG_CUDA(ctx, wait_stream(ctx));
        cudaEventElapsedTime(&ctx->stats.kernel_ms, ctx->ev[1], ctx->ev[2]);
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
    }
    return FG_OK;
}

int fg_render_rgb8(fg_ctx* ctx, const fg_params* p, int algo, int color_mode, const uint8_t* rgb_in,
                   const float* offsets, uint8_t* rgb_out) {
    if (!ctx) return FG_ERR_INVALID;
    if (!ctx->subs.empty()) return multi_render_rgb8_host(ctx, p, algo, color_mode, rgb_in, offsets, rgb_out);
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    ctx->stats = fg_stats{};
    ctx->fb_pending = false;
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (!rgb_in || !offsets || !rgb_out) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    return color_render_host(ctx, p, algo, color_mode, rgb_in, offsets, rgb_out);
}

int fg_render_rgb8_device(fg_ctx* ctx, const fg_params* p, int algo, int color_mode, const uint8_t* d_rgb_in,
                          const float* d_offsets, uint8_t* d_rgb_out, int stream_sync) {
    if (!ctx) return FG_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    ctx->stats = fg_stats{};
    ctx->fb_pending = false;
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (!d_rgb_in || !d_offsets || !d_rgb_out) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    return color_render_device(ctx, p, algo, color_mode, d_rgb_in, d_offsets, d_rgb_out, stream_sync);
}

int fg_dump_cells(fg_ctx* ctx, const fg_params* p, int stream_kind, const int32_t* ij, const float* lambda_cell,
                  size_t n, uint32_t cap, uint32_t* q_out, float* grains_out) {
    if (!ctx) return FG_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    int rc = validate(ctx, p);
    if (rc) return rc;
    if (stream_kind != FG_STREAM_CELL && stream_kind != FG_STREAM_PIXEL) return set_err(ctx, FG_ERR_INVALID, "unknown stream_kind");
    if (!ij || !lambda_cell || !q_out || (cap && !grains_out)) return set_err(ctx, FG_ERR_INVALID, "NULL buffer");
    if (n == 0) return FG_OK;
    if ((rc = init_tables(ctx))) return rc;
    float zero2[2] = {0.0f, 0.0f};
    fg_params pp = *p;
    pp.n_samples = 1;
    RenderConsts c = make_consts(&pp, zero2);
    DevBuf d_ij, d_lam, d_q, d_g;
    auto cleanup = [&]() { release(d_ij); release(d_lam); release(d_q); release(d_g); };
    if ((rc = ensure(ctx, d_ij, n * 8)) || (rc = ensure(ctx, d_lam, n * 4)) || (rc = ensure(ctx, d_q, n * 4)) ||
        (rc = ensure(ctx, d_g, n * (size_t)(cap ? cap : 1) * 12))) { cleanup(); return rc; }
    cudaStream_t s = ctx->stream;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d_ij.p, ij, n * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_lam.p, lambda_cell, n * 4, cudaMemcpyHostToDevice, s)) != cudaSuccess ||
        (e = cudaMemsetAsync(d_g.p, 0, n * (size_t)(cap ? cap : 1) * 12, s)) != cudaSuccess) { cleanup(); return map_cuda_error(ctx, e, "dump upload"); }
    k_dump_cells<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((const int2*)d_ij.p, (const float*)d_lam.p, n, cap, stream_kind,
                                                               (uint32_t*)d_q.p, (float*)d_g.p, c);
    if ((e = cudaGetLastError()) != cudaSuccess ||
        (e = cudaMemcpyAsync(q_out, d_q.p, n * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
        (cap && (e = cudaMemcpyAsync(grains_out, d_g.p, n * (size_t)cap * 12, cudaMemcpyDeviceToHost, s)) != cudaSuccess) ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess) { cleanup(); return map_cuda_error(ctx, e, "dump cells"); }
    cleanup();
    return FG_OK;
}

int fg_measure_issue_peak(fg_ctx* ctx, double out[4]) {
    if (!ctx || !out) return FG_ERR_INVALID;
    if (!ctx->subs.empty()) ctx = ctx->subs[0];
    std::lock_guard<std::mutex> lock(ctx->mu);
    ScopedDevice dev(ctx->device);
    ctx->err.clear();
    int rc;
    if ((rc = ensure(ctx, ctx->misc, 64))) return rc;
    const uint32_t iters = 4096;
    const unsigned blocks = (unsigned)ctx->sm_count * 2, threads = 1024;
    const double ops_per_thread[4] = {64.0 * iters, 64.0 * iters, 64.0 * iters, 32.0 * iters};
    for (int kind = 0; kind < 4; ++kind) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            FG_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
            switch (kind) {
            case 0: k_issue_peak<0><<<blocks, threads, 0, ctx->stream>>>(iters, (uint32_t*)ctx->misc.p); break;
            case 1: k_issue_peak<1><<<blocks, threads, 0, ctx->stream>>>(iters, (uint32_t*)ctx->misc.p); break;
            case 2: k_issue_peak<2><<<blocks, threads, 0, ctx->stream>>>(iters, (uint32_t*)ctx->misc.p); break;
            default: k_issue_peak<3><<<blocks, threads, 0, ctx->stream>>>(iters, (uint32_t*)ctx->misc.p); break;
            }
            FG_CUDA(ctx, cudaGetLastError());
            FG_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
            FG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
            if (rep > 0 && ms < best) best = ms;
        }
        out[kind] = ops_per_thread[kind] * (double)blocks * threads / (best * 1e-3) / 1e9;
    }
    return FG_OK;
}

} // extern "C"
