// fg_kernels.cuh -- general ("direct") kernels: always correct for every parameter
// combination (both Poisson branches, const/lognormal radii, any delta/rm/zoom).  The
// pixel-wise direct kernel is the fallback of the tiled fast path (fg_tile.cuh); the
// grain-wise kernels are the grain-parallel rasteriser.
#pragma once
#include "fg_rng.cuh"

namespace fg {

// Per-render constants, built on the host from fg_params (fg_api.cu: make_consts()).
struct RenderConsts {
    uint64_t seed_cell;   // Params.seed ^ CELL_STREAM   (src/rng.rs:6, 41)
    uint64_t seed_pixel;  // Params.seed ^ PIXEL_STREAM  (src/rng.rs:7, 41)
    uint32_t seeding;
    int in_w, in_h, out_w, out_h;
    uint32_t n;           // n_samples
    float zoom, inv_zoom; // inv_zoom = 1.0f / zoom        (src/pixelwise.rs:21)
    float delta;
    float uscale_cell;    // Uniform::new(0.0, delta) scale (src/pixelwise.rs:64)
    float uscale_unit;    // Uniform::new(0.0, 1.0) scale   (src/grainwise.rs:47)
    float inv_samples;    // 1.0f / max(n,1) as f32         (src/pixelwise.rs:20)
    RadiusModel rad;
    int row_begin, row_end; // output row band
    // extreme sample offsets (tiled path window sizing)
    float off_min_x, off_max_x, off_min_y, off_max_y;
    // Cooperative cancel INSIDE a launch (render_with_input_image_cancelable, src/lib.rs:116-132; the viewer's
    // latest-job-wins worker, src/bin/viewer.rs:975-1028): a device word the host raises from a second stream while the
    // kernels run; CTAs that start after that return at once.  nullptr (no cancel flag armed): no check at all.
    const int* abort;
};

// CTA-uniform: every thread reads the word, the barrier ORs the verdicts (threads may read either side of the host's store)
__device__ __forceinline__ bool cta_aborted(const RenderConsts& c) {
    return c.abort != nullptr && __syncthreads_or(*(const volatile int*)c.abort != 0) != 0;
}
// warp-uniform variant for kernels whose warps work independently
__device__ __forceinline__ bool warp_aborted(const RenderConsts& c) {
    return c.abort != nullptr && __any_sync(0xFFFFFFFFu, *(const volatile int*)c.abort != 0);
}

// Plane::get_clamped (src/model.rs:60-67) at the cell->pixel mapping of src/pixelwise.rs:69-73
__device__ __forceinline__ float lambda_of_cell(const float* __restrict__ lambda, const RenderConsts& c,
                                                float sample_x, float sample_y) {
    long long ix = floor_i64(sample_x), iy = floor_i64(sample_y);
    ix = ix < 0 ? 0 : (ix > c.in_w - 1 ? c.in_w - 1 : ix);
    iy = iy < 0 ? 0 : (iy > c.in_h - 1 ? c.in_h - 1 : iy);
    return __ldg(lambda + (size_t)iy * (size_t)c.in_w + (size_t)ix);
}

// evaluate_indicator (src/pixelwise.rs:47-106): 1 if any grain of the cells within rm covers (xg,yg)
__device__ inline bool indicator_direct(const float* __restrict__ lambda, const RenderConsts& c, float xg, float yg) {
    const float rm = c.rad.rm, delta = c.delta;
    if (rm <= 0.0f) return false;
    int i0 = floor_i32(__fdiv_rn(__fsub_rn(xg, rm), delta));
    int i1 = floor_i32(__fdiv_rn(__fadd_rn(xg, rm), delta));
    int j0 = floor_i32(__fdiv_rn(__fsub_rn(yg, rm), delta));
    int j1 = floor_i32(__fdiv_rn(__fadd_rn(yg, rm), delta));
    if (i0 > i1 || j0 > j1) return false;
    for (long long i = i0; i <= i1; ++i) {
        uint64_t hcol = mix3_col(c.seed_cell, (int32_t)i);
        float sample_x = __fmul_rn(__int2float_rn((int)i), delta);
        for (long long j = j0; j <= j1; ++j) {
            float sample_y = __fmul_rn(__int2float_rn((int)j), delta);
            float lam = lambda_of_cell(lambda, c, sample_x, sample_y);
            if (lam <= 0.0f) continue;
            float expected = __fmul_rn(__fmul_rn(lam, delta), delta);
            if (expected <= 0.0f) continue;
            Xoshiro rng;
            seed_small_rng(rng, mix3_row(hcol, (int32_t)j), c.seeding);
            uint32_t q = poisson_f64(rng, (double)expected);
            for (uint32_t g = 0; g < q; ++g) {
                float cx = __fadd_rn(sample_x, uniform_f32(rng, c.uscale_cell));
                float cy = __fadd_rn(sample_y, uniform_f32(rng, c.uscale_cell));
                float radius = radius_sample_clamped(c.rad, rng);
                if (radius <= 0.0f) continue;
                float dx = __fsub_rn(xg, cx), dy = __fsub_rn(yg, cy);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= __fmul_rn(radius, radius)) return true;
            }
        }
    }
    return false;
}

// One thread per output pixel; per-sample regeneration exactly like src/pixelwise.rs:31-39.
// grid: (ceil(out_w/32), ceil(band_rows/8), n_planes), block (32,8).
__global__ void __launch_bounds__(256) k_pixelwise_direct(const float* __restrict__ lambda, size_t lambda_stride,
                                                           const float2* __restrict__ offsets_input,
                                                           float* __restrict__ out, size_t out_stride, RenderConsts c) {
    if (cta_aborted(c)) return;
    int x = blockIdx.x * 32 + threadIdx.x;
    int y = c.row_begin + blockIdx.y * 8 + threadIdx.y;
    if (x >= c.out_w || y >= c.row_end) return;
    const float* lam = lambda + lambda_stride * blockIdx.z;
    float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);
    float by = __fmul_rn(__fadd_rn((float)y, 0.5f), c.inv_zoom);
    uint32_t count = 0;
    for (uint32_t k = 0; k < c.n; ++k) {
        float2 o = __ldg(offsets_input + k);
        count += indicator_direct(lam, c, __fsub_rn(bx, o.x), __fsub_rn(by, o.y)) ? 1u : 0u;
    }
    out[out_stride * blockIdx.z + (size_t)y * c.out_w + x] = __fmul_rn((float)count, c.inv_samples);
}

// Same, for an explicit list of tiles (the tiled path's fallback list).  Work item = one
// 32 x 8 pixel chunk of one tile; CTAs stride over (tile, chunk) so that a few tall tiles still
// spread over the machine.  `chunks_per_tile` = ceil(max tile height / 8).
struct TileRef { int x0, y0, w, h, plane; };
__global__ void __launch_bounds__(256) k_pixelwise_direct_tiles(const float* __restrict__ lambda, size_t lambda_stride,
                                                                 const float2* __restrict__ offsets_input,
                                                                 float* __restrict__ out, size_t out_stride,
                                                                 const TileRef* __restrict__ tiles,
                                                                 const uint32_t* __restrict__ n_tiles, uint32_t tile_cap,
                                                                 uint32_t chunks_per_tile, uint32_t* __restrict__ n_total,
                                                                 RenderConsts c) {
    const uint32_t nt = min(*n_tiles, tile_cap);
    if (n_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_total, nt); // running count over the bands of a render
    const uint64_t work = (uint64_t)nt * chunks_per_tile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint64_t wi = blockIdx.x; wi < work; wi += gridDim.x) {
        if (c.abort != nullptr && *(const volatile int*)c.abort != 0) return; // cancelled (no barriers below: threads may leave alone)
        const TileRef t = tiles[wi / chunks_per_tile];
        const int yl = (int)(wi % chunks_per_tile) * 8 + ty;
        if (yl >= t.h || tx >= t.w) continue;
        const int x = t.x0 + tx, y = t.y0 + yl;
        if (x >= c.out_w || y >= c.row_end) continue;
        const float* lam = lambda + lambda_stride * t.plane;
        float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);
        float by = __fmul_rn(__fadd_rn((float)y, 0.5f), c.inv_zoom);
        uint32_t count = 0;
        for (uint32_t k = 0; k < c.n; ++k) {
            float2 o = __ldg(offsets_input + k);
            count += indicator_direct(lam, c, __fsub_rn(bx, o.x), __fsub_rn(by, o.y)) ? 1u : 0u;
        }
        out[out_stride * t.plane + (size_t)y * c.out_w + x] = __fmul_rn((float)count, c.inv_samples);
    }
}

// ---- debug: grain realisation of listed cells (fg_dump_cells) ---------------------------
__global__ void k_dump_cells(const int2* __restrict__ ij, const float* __restrict__ lambda_cell, size_t n,
                             uint32_t cap, int stream_kind, uint32_t* __restrict__ q_out,
                             float* __restrict__ grains, RenderConsts c) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = ij[t].x, j = ij[t].y;
    float lam = lambda_cell[t];
    float ox, oy, scale, mean;
    uint64_t s;
    if (stream_kind == 2) { // grain-wise unit cell (src/grainwise.rs:37-53)
        ox = (float)i; oy = (float)j; scale = c.uscale_unit; mean = lam; s = c.seed_pixel;
    } else {               // pixel-wise delta cell (src/pixelwise.rs:68-88)
        ox = __fmul_rn(__int2float_rn(i), c.delta); oy = __fmul_rn(__int2float_rn(j), c.delta);
        scale = c.uscale_cell; mean = __fmul_rn(__fmul_rn(lam, c.delta), c.delta); s = c.seed_cell;
    }
    uint32_t q = 0;
    if (lam > 0.0f && mean > 0.0f) {
        Xoshiro rng;
        seed_small_rng(rng, mix3_row(mix3_col(s, i), j), c.seeding);
        q = poisson_f64(rng, (double)mean);
        for (uint32_t g = 0; g < q; ++g) {
            float cx = __fadd_rn(ox, uniform_f32(rng, scale));
            float cy = __fadd_rn(oy, uniform_f32(rng, scale));
            float r = radius_sample_clamped(c.rad, rng);
            if (g < cap) {
                float* o = grains + (t * cap + g) * 3;
                o[0] = cx; o[1] = cy; o[2] = r;
            }
        }
    }
    q_out[t] = q;
}

// ---- grain-wise: grain-parallel rasteriser (src/grainwise.rs:12-124) ---------------------
// Pass 1: Poisson count per input pixel of rows [iy0, iy1), keeping the pixel's generator state.  Pass 2
// (after an exclusive scan) continues each pixel's own RNG stream and writes its grains in pixel order,
// so the realisation does not depend on the traversal order.
struct GrainRec { float cxz, cyz, radius_out, radius_sq; }; // centre in OUTPUT px (cx*zoom), R, R^2

__global__ void __launch_bounds__(256) k_gw_count(const float* __restrict__ lambda, int iy0, int iy1,
                                                   uint32_t* __restrict__ counts, ulonglong4* __restrict__ states, RenderConsts c) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t npix = (size_t)(iy1 - iy0) * c.in_w;
    if (t >= npix) return;
    int x = (int)(t % c.in_w), y = iy0 + (int)(t / c.in_w);
    float lam = __ldg(lambda + (size_t)y * c.in_w + x);
    uint32_t q = 0;
    if (lam > 0.0f) {
        Xoshiro rng;
        seed_small_rng(rng, mix3_row(mix3_col(c.seed_pixel, x), y), c.seeding);
        q = poisson_f64(rng, (double)lam);
        // the generator right after the Poisson draw: the fill pass continues from here instead of
        // seeding and sampling the pixel a second time
        if (q) states[t] = make_ulonglong4(rng.s0, rng.s1, rng.s2, rng.s3);
    }
    counts[t] = q;
}

__global__ void __launch_bounds__(256) k_gw_fill(int iy0, int iy1, const uint32_t* __restrict__ counts,
                                                  const ulonglong4* __restrict__ states, const uint64_t* __restrict__ offsets_excl,
                                                  GrainRec* __restrict__ grains, RenderConsts c) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t npix = (size_t)(iy1 - iy0) * c.in_w;
    if (t >= npix) return;
    const uint32_t q = counts[t];
    if (!q) return;
    int x = (int)(t % c.in_w), y = iy0 + (int)(t / c.in_w);
    const ulonglong4 st = states[t];
    Xoshiro rng;
    rng.s0 = st.x; rng.s1 = st.y; rng.s2 = st.z; rng.s3 = st.w;
    GrainRec* dst = grains + offsets_excl[t];
    for (uint32_t g = 0; g < q; ++g) {
        float cx = __fadd_rn((float)x, uniform_f32(rng, c.uscale_unit));
        float cy = __fadd_rn((float)y, uniform_f32(rng, c.uscale_unit));
        float radius = radius_sample_clamped(c.rad, rng);
        GrainRec rec;
        rec.cxz = __fmul_rn(cx, c.zoom);
        rec.cyz = __fmul_rn(cy, c.zoom);
        float radius_out = (radius > 0.0f) ? __fmul_rn(radius, c.zoom) : 0.0f; // skip radius<=0 (:58-65)
        rec.radius_out = radius_out;
        rec.radius_sq = __fmul_rn(radius_out, radius_out);
        dst[g] = rec;
    }
}

// packed f32x2 arithmetic (one FADD2 / FMUL2 per pair; round-to-nearest per element, i.e. exactly the
// two scalar operations of the reference)
struct f32x2 { uint64_t v; };
__device__ __forceinline__ f32x2 f2_make(float x, float y) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void f2_split(f32x2 a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); }
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// One thread per grain, looping over the N sample offsets: rasterise the zoomed disk at
// centre + offset[k] and OR bit k into the per-pixel coverage mask (src/grainwise.rs:66-101).
//   * bounds() (src/grainwise.rs:126-142) without branches: [ceil((c-r)-0.5), floor((c+r)-0.5)] clipped
//     to the image / row band by max/min; the range is empty iff lo > hi afterwards (clipping one end
//     only never turns an empty or outside range into a non-empty one), x and y packed as f32x2;
//   * a box of ONE pixel (every disk of diameter <= 1 output pixel, r * zoom <= 0.5, away from the exact
//     half-pixel case) is one test and one reduction; boxes up to 2 x 2 are four predicated tests;
//     larger boxes take the reference's loops.  The reference's per-row reject `dy_sq > radius_sq` is
//     implied by the pixel test (dx*dx >= 0, round-to-nearest addition is monotone), so the
//     straight-line paths omit it;
//   * the offsets of the pass sit in shared memory and the next one is fetched while the current one is
//     processed; the sample loop runs word by word of the mask (32 samples share a word index).
// Lanes of a warp hold consecutive grains of the same input pixels, so their reductions land on
// neighbouring mask words.  IDX = uint32_t when the whole image's mask has fewer than 2^32 words.
#define FG_GW_OFF_CHUNK 2048 // sample offsets staged in shared memory per pass (16 KB), a multiple of 32
// One (grain, sample) pair of the global-mask rasteriser: bounds() clipped to the image / row band, then
// the pixel tests.  `t` = (tx, ty).  Returns false when the box is empty (nothing was written).
template <typename IDX, bool TEST_ONLY>
__device__ __forceinline__ bool gw_pair_global(f32x2 t, f32x2 rr, float R2, int last_x, int lo_y, int hi_y, uint32_t* bw,
                                               IDX xstep, IDX ystep, uint32_t bit) {
    const f32x2 half2 = f2_make(0.5f, 0.5f), one2 = f2_make(1.0f, 1.0f);
    float lx, ly, hx, hy;
    f2_split(f2_sub(f2_sub(t, rr), half2), lx, ly);
    f2_split(f2_sub(f2_add(t, rr), half2), hx, hy);
    const int x_min = max(__float2int_ru(lx), 0), x_max = min(__float2int_rd(hx), last_x);
    const int y_min = max(__float2int_ru(ly), lo_y), y_max = min(__float2int_rd(hy), hi_y);
    const int wx = x_max - x_min, wy = y_max - y_min;
    if ((wx | wy) < 0) return false; // empty in x or y
    if (TEST_ONLY) return true;
    const IDX i00 = (IDX)y_min * ystep + (IDX)x_min * xstep;
    const f32x2 f0 = f2_add(f2_make((float)x_min, (float)y_min), half2); // pixel centre (ox + 0.5, oy + 0.5)
    float sx0, sy0;
    f2_split(f2_mul(f2_sub(f0, t), f2_sub(f0, t)), sx0, sy0);
    if ((wx | wy) == 0) { // one pixel
        if (__fadd_rn(sx0, sy0) <= R2) atomicOr(bw + i00, bit);
        return true;
    }
    if (wx <= 1 && wy <= 1) {
        // (ox + 1) + 0.5 == (ox + 0.5) + 1 exactly: |ox| < 2^22 after clipping to the image
        const f32x2 f1 = f2_add(f0, one2);
        float sx1, sy1;
        f2_split(f2_mul(f2_sub(f1, t), f2_sub(f1, t)), sx1, sy1);
        const bool x1 = wx > 0, y1 = wy > 0;
        if (__fadd_rn(sx0, sy0) <= R2) atomicOr(bw + i00, bit);
        if (x1 && __fadd_rn(sx1, sy0) <= R2) atomicOr(bw + (i00 + xstep), bit);
        if (y1 && __fadd_rn(sx0, sy1) <= R2) atomicOr(bw + (i00 + ystep), bit);
        if (x1 && y1 && __fadd_rn(sx1, sy1) <= R2) atomicOr(bw + (i00 + ystep + xstep), bit);
        return true;
    }
    float tx, ty;
    f2_split(t, tx, ty);
    uint32_t* p00 = bw + i00;
    for (int oy = y_min; oy <= y_max; ++oy, p00 += ystep) {
        const float dy = __fsub_rn(__fadd_rn((float)oy, 0.5f), ty);
        const float dy_sq = __fmul_rn(dy, dy);
        if (dy_sq > R2) continue;
        uint32_t* p = p00;
        for (int ox = x_min; ox <= x_max; ++ox, p += xstep) {
            const float dx = __fsub_rn(__fadd_rn((float)ox, 0.5f), tx);
            if (__fadd_rn(__fmul_rn(dx, dx), dy_sq) <= R2) atomicOr(p, bit);
        }
    }
    return true;
}

// SPARSE: disks much smaller than a pixel (2R << 1) leave most boxes empty, but in a warp some lane almost
// always has a non-empty one, so the warp would walk the box arithmetic (four F2I on the quarter-rate XU pipe:
// the measured bound of the dense loop on such inputs) and the pixel tests for nearly every sample.  The
// sparse variant first runs a cheap NECESSARY condition for "the box holds a pixel centre" on the 32 samples
// of a mask word -- the distance from tx - 0.5 to the nearest integer (magic-number rounding, FMA pipe only)
// is at most R plus a bound on the f32 rounding of the reference's box arithmetic, in x and in y -- and keeps
// the survivors as a bit mask per lane; then it drains the masks through the exact code.  A pair that fails
// the condition has an empty box (so the reference writes nothing for it); survivors run the same arithmetic
// as in the dense loop.  Valid for image coordinates below 2^21 (checked by the host).
template <typename IDX, bool SPARSE>
__global__ void __launch_bounds__(256) k_gw_splat(const GrainRec* __restrict__ grains, const uint64_t* __restrict__ n_grains_ptr,
                                                   const float2* __restrict__ offsets, uint32_t* __restrict__ bits,
                                                   uint32_t lanes32, RenderConsts c) {
    __shared__ float2 s_off[FG_GW_OFF_CHUNK + 1];
    if (cta_aborted(c)) return;
    const uint64_t total = *n_grains_ptr;
    const int last_x = c.out_w - 1, lo_y = c.row_begin, hi_y = c.row_end - 1;
    const IDX xstep = (IDX)lanes32, ystep = (IDX)c.out_w * (IDX)lanes32;
    const f32x2 half2 = f2_make(0.5f, 0.5f), magic2 = f2_make(12582912.0f, 12582912.0f); // 1.5 * 2^23
    // 8 ulp of the largest coordinate that can still reach the image (box arithmetic: 3 roundings, u: 1, margin x2)
    const float eps_box = (float)(max(c.out_w, c.out_h) + 64) * 9.5367431640625e-7f; // * 2^-20
    for (uint32_t k0 = 0; k0 < c.n; k0 += FG_GW_OFF_CHUNK) {
        const uint32_t kn = min((uint32_t)FG_GW_OFF_CHUNK, c.n - k0);
        if (k0) __syncthreads();
        for (uint32_t t = threadIdx.x; t <= kn; t += 256) s_off[t] = t < kn ? __ldg(offsets + k0 + t) : make_float2(0.0f, 0.0f);
        __syncthreads();
        // mask word of sample k0 in row 0 of the image (the band's mask starts at row_begin)
        uint32_t* const bits0 = bits + (IDX)(k0 >> 5) - (IDX)c.row_begin * ystep;
        for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x) {
            const float4 rv = __ldg((const float4*)(grains + g));
            const float R = rv.z, R2 = rv.w;
            if (!(R > 0.0f)) continue;
            const f32x2 ctr = f2_make(rv.x, rv.y), rr = f2_make(R, R);
            // |(tx - 0.5) - nearest integer| <= R + eps is necessary for a pixel centre in [(tx - R) - 0.5, (tx + R) - 0.5]:
            // eps covers the roundings of those two expressions and of tx - 0.5 (each <= 1 ulp of a coordinate < 2^21 + reach)
            const float Re = __fadd_rn(R, eps_box);
            float2 o = s_off[0];
            for (uint32_t kw = 0; kw < kn; kw += 32) {
                uint32_t* const bw = bits0 + (IDX)(kw >> 5);
                const uint32_t ke = min(32u, kn - kw);
                uint32_t surv = 0;
#pragma unroll 4
                for (uint32_t kb = 0; kb < ke; ++kb) {
                    const f32x2 t = f2_add(ctr, f2_make(o.x, o.y)); // (tx, ty) = (cx*zoom + ox, cy*zoom + oy)
                    o = s_off[kw + kb + 1];                         // next sample's offset (entry kn is padding)
                    if (SPARSE) {
                        const f32x2 u = f2_sub(t, half2);
                        float dx, dy;
                        f2_split(f2_sub(u, f2_sub(f2_add(u, magic2), magic2)), dx, dy); // u - rint(u), exact for |u| < 2^22
                        surv |= (fabsf(dx) <= Re && fabsf(dy) <= Re) ? (1u << kb) : 0u;
                    } else {
                        gw_pair_global<IDX, false>(t, rr, R2, last_x, lo_y, hi_y, bw, xstep, ystep, 1u << kb);
                    }
                }
                while (SPARSE && surv) {
                    const uint32_t kb = (uint32_t)__ffs(surv) - 1u;
                    surv &= surv - 1u;
                    const float2 ok = s_off[kw + kb];
                    gw_pair_global<IDX, false>(f2_add(ctr, f2_make(ok.x, ok.y)), rr, R2, last_x, lo_y, hi_y, bw, xstep, ystep, 1u << kb);
                }
            }
        }
    }
}

// popcount / N epilogue (src/grainwise.rs:114-122)
__global__ void __launch_bounds__(256) k_gw_reduce(const uint32_t* __restrict__ bits, uint32_t lanes32,
                                                    float* __restrict__ out, RenderConsts c) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t npix = (size_t)(c.row_end - c.row_begin) * c.out_w;
    if (t >= npix) return;
    uint32_t count = 0;
    for (uint32_t l = 0; l < lanes32; ++l) count += __popc(bits[t * lanes32 + l]);
    out[(size_t)c.row_begin * c.out_w + t] = __fmul_rn((float)count, c.inv_samples);
}

// ---- issue-rate microbenchmarks (ALU roofline denominator) ------------------------------
template <int KIND>
__global__ void __launch_bounds__(1024) k_issue_peak(uint32_t iters, uint32_t* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) { // FFMA, 8 independent chains
        float a0 = tid * 1e-9f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
        const float m = 1.0000001f, b = 1e-7f;
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
                a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
            }
        }
        float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
        if (s == 123.456f) sink[0] = tid;
    } else if (KIND == 1) { // IMAD u32
        uint32_t a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3, a4 = tid + 4, a5 = tid + 5, a6 = tid + 6, a7 = tid + 7;
        const uint32_t m = 2654435761u + (iters & 2), b = 40503u;
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = a0 * m + b; a1 = a1 * m + b; a2 = a2 * m + b; a3 = a3 * m + b;
                a4 = a4 * m + b; a5 = a5 * m + b; a6 = a6 * m + b; a7 = a7 * m + b;
            }
        }
        uint32_t s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
        if (s == 0x12345678u) sink[0] = tid;
    } else if (KIND == 2) { // interleaved IMAD (fma pipe) + LOP3/SHF (alu pipe): the hashing mix
        uint32_t a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3, b0 = tid + 4, b1 = tid + 5, b2 = tid + 6, b3 = tid + 7;
        const uint32_t m = 2654435761u + (iters & 2), c = 40503u;
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = a0 * m + c; b0 = (b0 ^ (b0 >> 7)) ^ a1;
                a1 = a1 * m + c; b1 = (b1 ^ (b1 >> 9)) ^ a2;
                a2 = a2 * m + c; b2 = (b2 ^ (b2 >> 11)) ^ a3;
                a3 = a3 * m + c; b3 = (b3 ^ (b3 >> 13)) ^ a0;
            }
        }
        uint32_t s = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
        if (s == 0x12345678u) sink[0] = tid;
    } else { // DFMA
        double a0 = tid * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
        const double m = 1.0000001, b = 1e-7;
        for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b); }
        }
        double s = a0 + a1 + a2 + a3;
        if (s == 123.456) sink[0] = tid;
    }
}

} // namespace fg
