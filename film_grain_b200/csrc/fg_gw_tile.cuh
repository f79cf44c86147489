// fg_gw_tile.cuh -- grain-wise integrator, one CTA per OUTPUT TILE (src/grainwise.rs:12-124).
//
// k_gw_splat (fg_kernels.cuh) ORs every covered (pixel, sample) bit into a global mask: on C3 that is
// 2.1e9 reductions resolved in L2 -- one 32-bit atomic per L2 slice and clock, the measured bound of
// that kernel (its instruction count can be halved without changing its time).  Here the coverage
// mask lives in shared memory:
//   * a CTA owns FG_GT_W x FG_GT_H output pixels and a mask of 32*W samples per pixel ([word][pixel], so
//     neighbouring pixels fall into different banks); the sample loop runs in passes of 32*W samples,
//     each pass ends with popcount -> per-pixel counters, and the tile is written once as count * (1/N):
//     the global mask, its memset and the separate popcount kernel disappear;
//   * the grains that can reach the tile are those of the input pixels within (max |offset| + r_max) of
//     it: per input row one contiguous slice of the per-pixel CSR grain array (k_gw_count / k_gw_fill),
//     flattened over the rows by a prefix array so that all threads have grains to work on;
//   * per (grain, sample) the arithmetic is the reference's: tx = cx*zoom + ox, bounds(), pixel test --
//     with the box clipped to the tile instead of the image.  Tiles partition the band, so every
//     (pixel, sample) bit is produced exactly once, by the same f32 operations.
// Grains in the margin are processed by up to four tiles (x1.15 pairs on C3); that is the price of
// replacing 2.1e9 L2 atomics by shared-memory ones.
#pragma once
#include "fg_kernels.cuh"

namespace fg {

#define FG_GT_W 128        // tile width (output pixels)
#define FG_GT_H 64         // tile height
#define FG_GT_PIX (FG_GT_W * FG_GT_H)
#define FG_GT_WORDS 4      // mask words per pixel and pass (128 samples)
#define FG_GT_THREADS 1024
#define FG_GT_MAXROWS 512  // input rows of a tile's footprint

struct GwTileSmem {
    uint32_t mask[FG_GT_WORDS * FG_GT_PIX]; // [word][pixel]
    uint32_t cnt[FG_GT_PIX];
    float2 off[32 * FG_GT_WORDS + 1];
    uint32_t rowpref[FG_GT_MAXROWS + 1];     // grains of footprint rows [0, r)
    uint64_t rowstart[FG_GT_MAXROWS];        // first grain of the row's slice
};

// Input-pixel footprint [lo, hi) along one axis of output range [t0, t1): every grain centre c (input px)
// with some sample's disk reaching the range satisfies  t0 - 1 < c*zoom + off + R  and  c*zoom + off - R < t1,
// widened by one input pixel of slack against rounding (a wider footprint only adds work).
__device__ __forceinline__ void gw_footprint(int t0, int t1, float off_min, float off_max, float rmax, float inv_zoom, int limit,
                                             int& lo, int& hi) {
    const float a = ((float)t0 - 1.0f - off_max - rmax) * inv_zoom - 2.0f;
    const float b = ((float)t1 + 1.0f - off_min + rmax) * inv_zoom + 2.0f;
    lo = a > 0.0f ? (a < (float)limit ? (int)a : limit) : 0;
    hi = b < (float)limit ? (b > 0.0f ? (int)b + 1 : 0) : limit;
    if (hi > limit) hi = limit;
    if (hi < lo) hi = lo;
}

__global__ void __launch_bounds__(FG_GT_THREADS, 1)
k_gw_tile(const GrainRec* __restrict__ grains, const uint64_t* __restrict__ excl, size_t npix_in, const uint64_t* __restrict__ n_grains_ptr,
          int iy0, int iy1, const float2* __restrict__ offsets, float* __restrict__ out, int tiles_x, RenderConsts c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GwTileSmem& sm = *reinterpret_cast<GwTileSmem*>(smem_raw);
    const int tid = threadIdx.x;
    if (cta_aborted(c)) return;
    const int tx0 = (int)(blockIdx.x % (unsigned)tiles_x) * FG_GT_W, ty0 = c.row_begin + (int)(blockIdx.x / (unsigned)tiles_x) * FG_GT_H;
    const int tx1 = min(tx0 + FG_GT_W, c.out_w), ty1 = min(ty0 + FG_GT_H, c.row_end); // exclusive
    const uint64_t total = *n_grains_ptr;

    // ---- footprint: per input row one slice of the grain array ----
    const float rmax = __fmul_rn(c.rad.rm, c.zoom); // radius is clamped to rm before the zoom (src/grainwise.rs:55-61)
    int fx0, fx1, fy0, fy1;
    gw_footprint(tx0, tx1, c.off_min_x, c.off_max_x, rmax, c.inv_zoom, c.in_w, fx0, fx1);
    gw_footprint(ty0, ty1, c.off_min_y, c.off_max_y, rmax, c.inv_zoom, c.in_h, fy0, fy1);
    fy0 = max(fy0, iy0); fy1 = min(fy1, iy1); // rows whose grains exist (the host generated every row that reaches the band)
    int rows = max(fy1 - fy0, 0);
    if (rows > FG_GT_MAXROWS) __trap(); // the host checks the footprint before choosing this kernel; never render a wrong tile silently
    for (int r = tid; r < rows; r += FG_GT_THREADS) {
        const size_t p0 = (size_t)(fy0 + r - iy0) * c.in_w + fx0, p1 = (size_t)(fy0 + r - iy0) * c.in_w + fx1;
        const uint64_t a = __ldg(excl + p0), b = p1 < npix_in ? __ldg(excl + p1) : total;
        sm.rowstart[r] = a;
        sm.rowpref[r + 1] = fx1 > fx0 ? (uint32_t)(b - a) : 0u; // lengths first, prefix below
    }
    for (int p = tid; p < FG_GT_PIX; p += FG_GT_THREADS) sm.cnt[p] = 0;
    __syncthreads();
    if (tid < 32) { // inclusive scan of the row lengths by one warp
        uint32_t run = 0;
        if (tid == 0) sm.rowpref[0] = 0;
        for (int r0 = 0; r0 < rows; r0 += 32) {
            const uint32_t v = (r0 + tid < rows) ? sm.rowpref[r0 + tid + 1] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (tid >= d) incl += u;
            }
            if (r0 + tid < rows) sm.rowpref[r0 + tid + 1] = run + incl;
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    __syncthreads();
    const uint32_t n_tile = sm.rowpref[rows];

    const int last_x = tx1 - 1, last_y = ty1 - 1;
    const uint32_t lanes32 = (c.n + 31u) / 32u;
    const uint32_t wpp = min(lanes32, (uint32_t)FG_GT_WORDS); // mask words per pass
    const f32x2 half2 = f2_make(0.5f, 0.5f), one2 = f2_make(1.0f, 1.0f);
    const uint32_t mask_s = (uint32_t)__cvta_generic_to_shared(sm.mask);

    for (uint32_t k0 = 0; k0 < c.n; k0 += 32u * wpp) {
        const uint32_t kn = min(32u * wpp, c.n - k0);
        for (uint32_t p = tid; p < wpp * FG_GT_PIX; p += FG_GT_THREADS) sm.mask[p] = 0;
        for (uint32_t t = tid; t <= kn; t += FG_GT_THREADS) sm.off[t] = t < kn ? __ldg(offsets + k0 + t) : make_float2(0.0f, 0.0f);
        __syncthreads();
        for (uint32_t flat = tid; flat < n_tile; flat += FG_GT_THREADS) {
            // flat index -> (row, grain): largest r with rowpref[r] <= flat
            int lo = 0, hi = rows; // invariant rowpref[lo] <= flat < rowpref[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sm.rowpref[mid] <= flat) lo = mid; else hi = mid;
            }
            const float4 rv = __ldg((const float4*)(grains + (sm.rowstart[lo] + (flat - sm.rowpref[lo]))));
            const float R = rv.z, R2 = rv.w;
            if (!(R > 0.0f)) continue;
            const f32x2 ctr = f2_make(rv.x, rv.y), rr = f2_make(R, R);
            float2 o = sm.off[0];
            for (uint32_t kw = 0; kw < kn; kw += 32) {
                const uint32_t mw = mask_s + (kw >> 5) * (FG_GT_PIX * 4u);
                const uint32_t ke = min(32u, kn - kw);
#pragma unroll 4
                for (uint32_t kb = 0; kb < ke; ++kb) {
                    const f32x2 t = f2_add(ctr, f2_make(o.x, o.y)); // (tx, ty) = (cx*zoom + ox, cy*zoom + oy)
                    o = sm.off[kw + kb + 1];                        // next sample's offset (entry kn is padding)
                    float lx, ly, hx, hy;
                    f2_split(f2_sub(f2_sub(t, rr), half2), lx, ly);
                    f2_split(f2_sub(f2_add(t, rr), half2), hx, hy);
                    // bounds() clipped to the tile (a subset of the image and of the band)
                    const int x_min = max(__float2int_ru(lx), tx0), x_max = min(__float2int_rd(hx), last_x);
                    const int y_min = max(__float2int_ru(ly), ty0), y_max = min(__float2int_rd(hy), last_y);
                    const int wx = x_max - x_min, wy = y_max - y_min;
                    if ((wx | wy) < 0) continue; // empty in x or y
                    const uint32_t bit = 1u << kb;
                    const uint32_t a00 = mw + (uint32_t)((y_min - ty0) * FG_GT_W + (x_min - tx0)) * 4u;
                    const f32x2 f0 = f2_add(f2_make((float)x_min, (float)y_min), half2); // pixel centre
                    float sx0, sy0;
                    f2_split(f2_mul(f2_sub(f0, t), f2_sub(f0, t)), sx0, sy0);
                    if ((wx | wy) == 0) { // one pixel
                        if (__fadd_rn(sx0, sy0) <= R2) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a00), "r"(bit) : "memory");
                        continue;
                    }
                    if (wx <= 1 && wy <= 1) {
                        const f32x2 f1 = f2_add(f0, one2); // (ox + 1) + 0.5 == (ox + 0.5) + 1 exactly for image coordinates
                        float sx1, sy1;
                        f2_split(f2_mul(f2_sub(f1, t), f2_sub(f1, t)), sx1, sy1);
                        const bool x1 = wx > 0, y1 = wy > 0;
                        if (__fadd_rn(sx0, sy0) <= R2) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a00), "r"(bit) : "memory");
                        if (x1 && __fadd_rn(sx1, sy0) <= R2) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a00 + 4u), "r"(bit) : "memory");
                        if (y1 && __fadd_rn(sx0, sy1) <= R2) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a00 + FG_GT_W * 4u), "r"(bit) : "memory");
                        if (x1 && y1 && __fadd_rn(sx1, sy1) <= R2) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a00 + FG_GT_W * 4u + 4u), "r"(bit) : "memory");
                        continue;
                    }
                    float tx, ty;
                    f2_split(t, tx, ty);
                    uint32_t arow = a00;
                    for (int oy = y_min; oy <= y_max; ++oy, arow += FG_GT_W * 4u) {
                        const float dy = __fsub_rn(__fadd_rn((float)oy, 0.5f), ty);
                        const float dy_sq = __fmul_rn(dy, dy);
                        if (dy_sq > R2) continue;
                        uint32_t a = arow;
                        for (int ox = x_min; ox <= x_max; ++ox, a += 4u) {
                            const float dx = __fsub_rn(__fadd_rn((float)ox, 0.5f), tx);
                            if (__fadd_rn(__fmul_rn(dx, dx), dy_sq) <= R2) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(bit) : "memory");
                        }
                    }
                }
            }
        }
        __syncthreads();
        // popcount of the pass (src/grainwise.rs:114-120)
        for (int p = tid; p < FG_GT_PIX; p += FG_GT_THREADS) {
            uint32_t s = 0;
            for (uint32_t w = 0; w < wpp; ++w) s += __popc(sm.mask[w * FG_GT_PIX + p]);
            sm.cnt[p] += s;
        }
        __syncthreads();
    }
    // count * (1/N)  (src/grainwise.rs:121)
    for (int p = tid; p < FG_GT_PIX; p += FG_GT_THREADS) {
        const int x = tx0 + (p & (FG_GT_W - 1)), y = ty0 + (p / FG_GT_W);
        if (x < tx1 && y < ty1) out[(size_t)y * c.out_w + x] = __fmul_rn((float)sm.cnt[p], c.inv_samples);
    }
}

} // namespace fg
