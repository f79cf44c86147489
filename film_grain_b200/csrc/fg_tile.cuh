// fg_tile.cuh -- the pixel-wise fast path: one CTA per output STRIP SEGMENT.
//
// Reference semantics (src/pixelwise.rs:11-106): pixel = (1/N) * #{k : some grain of the cells
// within rm of the k-th shifted sample point covers it}.  The reference regenerates every cell
// for every sample of every pixel; here each CTA owns a strip of 32 output columns x SEG rows and
//
//   * generates every Boolean-model cell its window touches exactly ONCE into shared memory
//     (bit-identical RNG chain: fg_rng.cuh), sliding the window down the strip TH pixel rows at
//     a time through a ring of cell rows, so the only regenerated margin is the horizontal one;
//   * generation is two-phase and compacted: phase A runs the cheap first-draw test on all
//     cells (4 of the 8 PCG seed words + 1 xoshiro draw + an integer threshold compare that is
//     exactly `U1 <= exp(-lambda')`); only the non-empty cells (~27% at lambda'=1/pi) go through
//     the dense phase B (full seeding, Knuth continuation in f64, positions) -- no lane idles on
//     an empty cell;
//   * grains are stored per cell row in cell order (CSR): P[row][i] (u16 ring positions) and
//     G[pos] = (cx, cy); a sample's candidates in one cell row are one contiguous range;
//   * evaluation: lane = output column, warp = sample subset; the per-(column,sample) cell range
//     and the per-(row,sample) cell-row range are computed ONCE (IEEE division, exactly the
//     reference's expression) and reused across the rows / columns they do not depend on.  Per
//     cell row a sample runs FG_TILE_USLOTS straight-line distance tests folded into a running
//     minimum (one compare per sample), then an early-exit remainder loop; two samples advance
//     together so their shared-memory loads overlap.
//
// Anything the fast path cannot hold (lambda' >= 12 -> rejection branch, window or grain ring
// overflow, log-normal radii, exotic geometry) is appended to a fallback list and rendered by the
// general direct kernel (fg_kernels.cuh) in the same call: results are identical either way.
#pragma once
#include <cstdlib>
#include "fg_ctx.cuh"
#include "fg_kernels.cuh"
#include "fg_stage.cuh"

namespace fg {

#define FG_THR_EMPTY (1ULL << 53)          // first-draw threshold that no draw exceeds (cell skipped)
#define FG_THR_GENERAL 0xFFFFFFFFFFFFFFFFULL // cell needs the general path (lambda' >= 12 / non-finite)

// Per input pixel: thr = floor(exp(-lambda') * 2^53), e = exp(-lambda') with lambda' = lambda*delta*delta
// (src/pixelwise.rs:71-81).  The Knuth loop's first decision `p = U1 > e` with U1 = m * 2^-53
// (m = next_u64 >> 11) is exactly `m > thr`.
__global__ void __launch_bounds__(256) k_thresholds(const float* __restrict__ lambda, size_t in_stride, size_t first, size_t n,
                                                     float delta, uint64_t* __restrict__ thr, double* __restrict__ ev) {
    // blockIdx.y = plane; elements [first, first + n) of each plane (the input rows the band's cells map to)
    const size_t base = in_stride * blockIdx.y + first;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (size_t)gridDim.x * 256) {
        float lam = lambda[base + t];
        uint64_t th = FG_THR_EMPTY;
        double e = 1.0;
        if (lam > 0.0f) {
            float expected = __fmul_rn(__fmul_rn(lam, delta), delta);
            if (expected > 0.0f) {
                if (expected < 12.0f) {
                    e = exp(-(double)expected);
                    th = (uint64_t)(e * 9007199254740992.0);
                } else {
                    th = FG_THR_GENERAL;
                    e = -1.0; // the dense phase recognises "needs the general path" by e < 0
                }
            }
        } else if (lam != lam) {
            th = FG_THR_GENERAL;
            e = -1.0;
        }
        thr[base + t] = th;
        ev[base + t] = e;
    }
}

// First-draw bitmap: one bit per Boolean-model cell and colour plane, 1 = "the cell's first Knuth draw
// exceeds its threshold", i.e. the cell holds at least one grain (or needs the general path).  The
// hash, the four PCG seed words and the first xoshiro draw depend only on (seed, i, j): they are
// computed ONCE per cell here and compared against every plane's threshold, instead of once per plane
// and per strip window inside the strip kernel.  grid (ceil(rows/FG_BM_ROWS), ceil(cols/256)), block 256,
// one thread per cell column; each warp
// stores one aligned 32-bit word per plane.
#ifndef FG_BM_ROWS
#define FG_BM_ROWS 16 // cell rows per thread in k_first_draw_bitmap (amortises the column half of the hash)
#endif
template <int SEEDING, int NP> // compile-time seeding variant and plane count (NP = 0: run-time n_planes)
__global__ void __launch_bounds__(256) k_first_draw_bitmap(const uint64_t* __restrict__ thr_planes, size_t in_stride,
                                                            int n_planes, uint32_t* __restrict__ bm, size_t bm_plane_words,
                                                            int i0, int j0, int cols, int rows, uint32_t pitchw, int row_off, RenderConsts c) {
    if (cta_aborted(c)) return;
    const int col = blockIdx.y * 256 + threadIdx.x;
    const int row0 = row_off + blockIdx.x * FG_BM_ROWS; // rows [row_off, rows) of the rectangle (chunked upload: a slice per launch)
    const bool valid = col < cols;
    const bool store = (threadIdx.x & 31) == 0 && col < (int)(pitchw * 32u);
    const int i = i0 + col;
    const uint64_t hcol = mix3_col(c.seed_cell, i);
    const int ix = min(max(floor_i32(__fmul_rn(__int2float_rn(i), c.delta)), 0), c.in_w - 1);
    // the thresholds of a column change with the input row only (every 1/delta cell rows, the same row for every thread)
    uint64_t thc[NP ? NP : 1];
    int iy_held = -1;
#pragma unroll 4
    for (int rr = 0; rr < FG_BM_ROWS; ++rr) {
        const int row = row0 + rr;
        if (row >= rows) break; // uniform
        const int j = j0 + row;
        uint64_t m1 = 0;
        const int iy = min(max(floor_i32(__fmul_rn(__int2float_rn(j), c.delta)), 0), c.in_h - 1);
        const size_t pix = (size_t)iy * c.in_w + ix;
        if (valid) {
            const uint64_t h = mix3_row(hcol, j);
            uint64_t s0, s3;
            if (SEEDING == 0) { s0 = pcg_word_pair<0>(h); s3 = pcg_word_pair<3>(h); }
            else { Xoshiro t; seed_splitmix(t, h); s0 = t.s0; s3 = t.s3; }
            m1 = (rotl64(s0 + s3, 23) + s0) >> 11;
        }
        if (NP) {
            if (iy != iy_held) { // uniform
                iy_held = iy;
#pragma unroll
                for (int pl = 0; pl < (NP ? NP : 1); ++pl) thc[pl] = valid ? __ldg(thr_planes + in_stride * pl + pix) : FG_THR_EMPTY;
            }
#pragma unroll
            for (int pl = 0; pl < (NP ? NP : 1); ++pl) {
                const bool ne = valid && ((thc[pl] == FG_THR_GENERAL) || (m1 > thc[pl]));
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, ne);
                if (store) bm[bm_plane_words * pl + (size_t)row * pitchw + (col >> 5)] = m;
            }
        } else {
            for (int pl = 0; pl < n_planes; ++pl) {
                bool ne = false;
                if (valid) {
                    const uint64_t th64 = __ldg(thr_planes + in_stride * pl + pix);
                    ne = (th64 == FG_THR_GENERAL) || (m1 > th64);
                }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, ne);
                if (store) bm[bm_plane_words * pl + (size_t)row * pitchw + (col >> 5)] = m;
            }
        }
    }
}

struct __align__(16) ColInfo { uint64_t h; float sx; int ixc; }; // per cell column: mix3 first half, i*delta, clamped input x

struct TileCfg {
    int TH, SEG;          // pixel rows per step, rows per segment
    int CWB, RH, PS;      // cell-column bound, ring rows, P row stride (u16 elements)
    int R;                // cell rows per generation group
    int GCAP;             // grain ring capacity (<= 65535 entries; P holds ring indices as u16)
    int n_strips, n_segs;
    int bm_i0, bm_j0;     // first cell column / row of the first-draw bitmap
    int bm_cols, bm_rows; // its extent in cells
    uint32_t bm_pitchw;   // 32-bit words per bitmap row
    uint32_t ppitch;      // staged mode: cell-table prefix entries per row
    float r2c;            // constant radius: squared clamped radius (src/pixelwise.rs:89-97)
    uint32_t off_col, off_P, off_G, off_R2, off_list, off_E, off_cnt, off_wtot, off_pcount, off_wpair, off_rows, total;
};

#ifndef FG_TILE_WARPS
#define FG_TILE_WARPS 32
#endif
#define FG_TILE_THREADS (FG_TILE_WARPS * 32)
#define FG_TILE_ITERS 4 // phase-A cells per thread per group: R*(CW+1) <= FG_TILE_CELLS
#define FG_TILE_CELLS (FG_TILE_THREADS * FG_TILE_ITERS)
#define FG_TILE_NE (FG_TILE_ITERS * FG_TILE_WARPS) // (iteration, warp) compaction counters
#if FG_TILE_WARPS == 32
#define FG_TILE_SPW_MAX 8  // 32 warps x 8 = 256
#elif FG_TILE_WARPS == 16
#define FG_TILE_SPW_MAX 16 // samples per warp per chunk (even): N <= 256 in one chunk
#elif FG_TILE_WARPS == 24
#define FG_TILE_SPW_MAX 12 // 24 warps x 12 = 288
#else
#define FG_TILE_SPW_MAX 14 // 20 warps x 14 = 280 >= 256
#endif
#define FG_TILE_GPAD 8  // grain ring entries mirrored past the end (unrolled reads never wrap)
#ifndef FG_TILE_U3
#define FG_TILE_U3 5    // unrolled grain tests of the merged three-cell-row path
#endif
#define FG_TILE_USLOTS 3 // unrolled, predicated grain tests per cell-row range
#ifndef FG_TILE_UROW
#define FG_TILE_UROW 2  // staged 3-row path: predicated straight-line tests per cell row (0 = the merged FG_TILE_U3 walk)
#endif

__device__ __forceinline__ void push_fallback(TileRef* list, uint32_t* count, uint32_t cap, int x0, int y0, int w,
                                              int h, int plane) {
    uint32_t idx = atomicAdd(count, 1u);
    if (idx < cap) {
        TileRef t;
        t.x0 = x0; t.y0 = y0; t.w = w; t.h = h; t.plane = plane;
        list[idx] = t;
    }
}

// floor((v -/+ rm) / delta) exactly as src/pixelwise.rs:55-58
__device__ __forceinline__ int cell_lo(float v, float rm, float delta) { return floor_i32(__fdiv_rn(__fsub_rn(v, rm), delta)); }
__device__ __forceinline__ int cell_hi(float v, float rm, float delta) { return floor_i32(__fdiv_rn(__fadd_rn(v, rm), delta)); }

// explicit shared-window loads (32-bit addresses): keeps the hot loop free of generic-pointer
// window arithmetic.  volatile: never cached across the CTA barriers that separate generation
// from evaluation.
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v; // 32-bit destination: the load zero-extends, no separate conversion
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// asynchronous global -> shared copies (LDGSTS): no registers, completion by cp.async.wait_all
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// squared distance with the packed f32x2 pipe: (dx, dy) = p - g and (dx*dx, dy*dy) are one instruction
// each (FADD2 / FMUL2, round-to-nearest per element like the scalar forms), then dx*dx + dy*dy.
__device__ __forceinline__ float dist2_packed(uint64_t p, float2 g) {
    uint64_t gg, d, s;
    asm("mov.b64 %0, {%1, %2};" : "=l"(gg) : "f"(g.x), "f"(g.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(p), "l"(gg));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(s) : "l"(d));
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s));
    return __fadd_rn(lo, hi);
}
__device__ __forceinline__ uint64_t pack_f32x2(float x, float y) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}

// One grain test: if (U < n) dmin = min(dmin, |p - G[ga + U]|^2) -- the un-fused f32 sequence of
// src/pixelwise.rs:96-98 on the packed pipe.  The LOAD and the MIN are predicated (a lane never loads a grain that is
// not its own, which keeps the shared-memory wavefronts down); the three arithmetic instructions in between run on
// whatever `g` holds (the grain just loaded or a stale one) and their result is simply not used by an inactive lane.
// `g` is a register pair threaded through every third slot: three loads of a sample are in flight at a time, and no
// register is defined under a predicate only (that would extend its live range to the kernel entry).
template <int U>
__device__ __forceinline__ void slot_test(float& dmin, uint64_t& g, uint32_t n, uint32_t ga, uint64_t p) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .b64 d, s2;\n\t"
        ".reg .f32 lo, hi, dd;\n\t"
        "setp.gt.u32 q, %2, %5;\n\t"
        "@q ld.shared.b64 %1, [%3+%6];\n\t"
        "sub.rn.f32x2 d, %4, %1;\n\t"
        "mul.rn.f32x2 s2, d, d;\n\t"
        "mov.b64 {lo, hi}, s2;\n\t"
        "add.rn.f32 dd, lo, hi;\n\t"
        "@q min.f32 %0, %0, dd;\n\t"
        "}"
        : "+f"(dmin), "+l"(g)
        : "r"(n), "r"(ga), "l"(p), "n"(U), "n"(U * 8));
}
template <int U0, int U1>
struct SlotRun {
    static __device__ __forceinline__ void run(float& dmin, uint64_t (&g)[3], uint32_t n, uint32_t ga, uint64_t p) {
        slot_test<U0>(dmin, g[U0 % 3], n, ga, p);
        SlotRun<U0 + 1, U1>::run(dmin, g, n, ga, p);
    }
};
template <int U1>
struct SlotRun<U1, U1> {
    static __device__ __forceinline__ void run(float&, uint64_t (&)[3], uint32_t, uint32_t, uint64_t) {}
};

// the same for a sample PAIR (A, B) on one cell row each: slots of A and B alternate, each sample has its own two chains
template <int U0, int U1>
struct RowSlots {
    static __device__ __forceinline__ void run(float& dA, float& dB, uint64_t (&chA)[2], uint64_t (&chB)[2], uint32_t nA, uint32_t nB,
                                               uint32_t gaA, uint32_t gaB, uint64_t pA, uint64_t pB) {
        slot_test<U0>(dA, chA[U0 & 1], nA, gaA, pA);
        slot_test<U0>(dB, chB[U0 & 1], nB, gaB, pB);
        RowSlots<U0 + 1, U1>::run(dA, dB, chA, chB, nA, nB, gaA, gaB, pA, pB);
    }
};
template <int U1>
struct RowSlots<U1, U1> {
    static __device__ __forceinline__ void run(float&, float&, uint64_t (&)[2], uint64_t (&)[2], uint32_t, uint32_t, uint32_t, uint32_t, uint64_t, uint64_t) {}
};

// packed byte offsets of P[row][i0 - i_lo] and P[row][i1 - i_lo + 1] for sample abscissa xg (0 = no cells).
// Out of line on purpose: two IEEE divisions per call, 16 call sites in the unrolled sample loop.
__device__ __noinline__ uint32_t col_range_packed(float xg, float rm, float delta, int i_lo) {
    const int i0 = cell_lo(xg, rm, delta), i1 = cell_hi(xg, rm, delta);
    if (i0 > i1) return 0u;
    return (uint32_t)(2 * (i0 - i_lo)) | ((uint32_t)(2 * (i1 - i_lo + 1)) << 16);
}

// the cell table of fg_stage.cuh (STAGED instances of the strip kernel load their windows from it)
struct CellTable {
    const uint32_t* Pg;
    const uint64_t* rowbase;
    const float2* Gg;
    const float* R2g;
    const uint16_t* Cg; // cell column (mod 2^16) of every grain
};

template <int SPWC, bool LOGN, bool STAGED>
__global__ void __launch_bounds__(FG_TILE_THREADS, 1)
k_pixelwise_strip(const uint32_t* __restrict__ bm_planes, size_t bm_plane_words, const double* __restrict__ e_planes, size_t in_stride,
                  const float2* __restrict__ offsets_input, float* __restrict__ out, size_t out_stride,
                  TileRef* __restrict__ fb_list, uint32_t* __restrict__ fb_count, uint32_t fb_cap, TileCfg cfg,
                  RenderConsts c, CellTable tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    ColInfo* colT = (ColInfo*)(smem + cfg.off_col);
    uint16_t* P = (uint16_t*)(smem + cfg.off_P);
    float2* G = (float2*)(smem + cfg.off_G);
    float* R2 = (float*)(smem + cfg.off_R2); // LOGN only: squared clamped radius per grain (-1: never covers)
    uint16_t* list = (uint16_t*)(smem + cfg.off_list);
    uint16_t* E = (uint16_t*)(smem + cfg.off_E);
    uint32_t* cntA = (uint32_t*)(smem + cfg.off_cnt);  // [FG_TILE_NE] non-empty counts per (iteration, warp)
    uint32_t* wtot = (uint32_t*)(smem + cfg.off_wtot); // [NW] warp totals, [NW] general-path flag
    uint32_t* pcount = (uint32_t*)(smem + cfg.off_pcount);
    float2* wpair = (float2*)(smem + cfg.off_wpair);
    uint32_t* rowA = (uint32_t*)(smem + cfg.off_rows); // STAGED: [RH] first table prefix of the window, [RH] count, [RH] ring position
    uint32_t* extb = rowA + (3 * cfg.RH + 3) / 4 * 4;                // STAGED: [RH][4] prefetched {Pg[first], Pg[last], rowbase lo, hi} of the next step's rows

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (cta_aborted(c)) return; // a cancelled render
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int unit = blockIdx.x;
    const int strip = unit % cfg.n_strips;
    const int seg = (unit / cfg.n_strips) % cfg.n_segs;
    const int plane = unit / (cfg.n_strips * cfg.n_segs);
    const int X0 = strip * 32;
    const int Y0 = c.row_begin + seg * cfg.SEG;
    if (X0 >= c.out_w || Y0 >= c.row_end) return;
    const int X1 = min(X0 + 32, c.out_w) - 1;      // inclusive
    const int Y1 = min(Y0 + cfg.SEG, c.row_end);   // exclusive
    const uint32_t* bm = bm_planes + bm_plane_words * plane;
    const double* ev = e_planes + in_stride * plane;
    float* outp = out + out_stride * plane;
    const float rm = c.rad.rm, delta = c.delta;
    const float r2 = cfg.r2c; // min(mean radius, rm)^2, squared on the host in f32
    const bool radius_ok = LOGN || (c.rad.mean_linear > rm ? rm : c.rad.mean_linear) > 0.0f; // const radius <= 0: grains never cover
    const uint32_t GC = (uint32_t)cfg.GCAP;

    // ---- horizontal cell window of the strip (monotone in x and in the offset) ----
    const float bx0 = __fmul_rn(__fadd_rn((float)X0, 0.5f), c.inv_zoom);
    const float bx1 = __fmul_rn(__fadd_rn((float)X1, 0.5f), c.inv_zoom);
    const int i_lo = cell_lo(__fsub_rn(bx0, c.off_max_x), rm, delta);
    const int i_hi = cell_hi(__fsub_rn(bx1, c.off_min_x), rm, delta);
    const long long CWl = (long long)i_hi - (long long)i_lo + 1;
    if (CWl < 1 || CWl > cfg.CWB || i_lo < cfg.bm_i0 || (long long)i_hi >= (long long)cfg.bm_i0 + cfg.bm_cols) { // uniform
        if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
        return;
    }
    const int CW = (int)CWl, CW1 = CW + 1;
    // STAGED: prefix rows are stored from the table column rounded down to a multiple of four (16-byte
    // vector loads); logical window column c lives at index c + sh
    const int sh = STAGED ? ((i_lo - cfg.bm_i0) & 3) : 0;
    const int PS = cfg.PS, RH = cfg.RH;

    if (!STAGED)
    for (int il = tid; il < CW; il += FG_TILE_THREADS) {
        ColInfo ci;
        int i = i_lo + il;
        ci.h = mix3_col(c.seed_cell, i);
        ci.sx = __fmul_rn(__int2float_rn(i), delta);
        ci.ixc = min(max(floor_i32(ci.sx), 0), c.in_w - 1);
        colT[il] = ci;
    }
    for (int p = tid; p < cfg.TH * 32; p += FG_TILE_THREADS) pcount[p] = 0;
    if (tid == 0) wtot[FG_TILE_WARPS] = 0;

    // phase-A cell assignment of this thread: cell c = tid + 512*it -> (r, il), il == CW is the row sentinel
    int rci[FG_TILE_ITERS];
#pragma unroll
    for (int it = 0; it < FG_TILE_ITERS; ++it) {
        int cc = tid + FG_TILE_THREADS * it;
        int r = cc / CW1;
        rci[it] = (r < cfg.R) ? ((r << 16) | (cc - r * CW1)) : -1;
    }

    // per-thread (column, sample) data: sample point x and its cell-column range [a, b)
    const int x = X0 + lane;
    const bool xvalid = x <= X1;
    const float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);
    float xg_r[SPWC];
    uint32_t ip_r[SPWC];
    const int n_chunks = (int)((c.n + FG_TILE_WARPS * SPWC - 1) / (FG_TILE_WARPS * SPWC));
    auto load_xk = [&](int chunk) {
#pragma unroll
        for (int s = 0; s < SPWC; ++s) {
            uint32_t k = (uint32_t)chunk * (FG_TILE_WARPS * SPWC) + s * FG_TILE_WARPS + warp;
            float xg = 0.0f;
            uint32_t ip = 0;
            if (k < c.n && xvalid) {
                float2 o = __ldg(offsets_input + k);
                xg = __fsub_rn(bx, o.x);
                ip = col_range_packed(xg, rm, delta, i_lo - sh);
            }
            xg_r[s] = xg;
            // no cells / inactive lane: an empty range at the first VALID prefix entry (entries below
            // the column shift hold junk)
            ip_r[s] = ip ? ip : (uint32_t)(2 * sh) * 0x10001u;
        }
    };
    bool xk_loaded = false;
    __syncthreads();

    // first-draw bitmap words of this thread's phase-A cells for the group starting at row jg
    uint32_t bmw[FG_TILE_ITERS], bsh[FG_TILE_ITERS];
#pragma unroll
    for (int it = 0; it < FG_TILE_ITERS; ++it) { bmw[it] = 0u; bsh[it] = (uint32_t)(i_lo - cfg.bm_i0 + (rci[it] & 0xFFFF)) & 31u; }
    int j_bm = INT_MIN; // group start row the words in bmw[] belong to
    auto load_bm = [&](int jg) {
        j_bm = jg;
#pragma unroll
        for (int it = 0; it < FG_TILE_ITERS; ++it) {
            const int rc = rci[it];
            const long long brow = (long long)jg - cfg.bm_j0 + (rc >> 16);
            if (rc >= 0 && (rc & 0xFFFF) < CW && brow >= 0 && brow < cfg.bm_rows) {
                const uint32_t bcol = (uint32_t)(i_lo - cfg.bm_i0 + (rc & 0xFFFF));
                bmw[it] = __ldg(bm + (size_t)brow * cfg.bm_pitchw + (bcol >> 5));
            }
        }
    };

    // ---- ring state (uniform across the CTA) ----
    uint32_t head = 0;            // ring index of the next grain, in [0, GC); P holds ring indices
    int j_gen = 0, rr_gen = 0;    // next cell row to generate and its ring row
    int j_lo_prev = 0, rr_lo = 0; // oldest live cell row and its ring row
    bool first = true;
    int r_cur = cfg.R;            // rows per generation group (adapts to the non-empty fraction)
    int pf_j0 = INT_MIN, pf_n = 0; // STAGED: cell rows whose extents were prefetched into extb

    for (int ya = Y0; ya < Y1; ya += cfg.TH) {
        const int yb = min(ya + cfg.TH, Y1) - 1; // inclusive
        const int th = yb - ya + 1;
        const float bya = __fmul_rn(__fadd_rn((float)ya, 0.5f), c.inv_zoom);
        const float byb = __fmul_rn(__fadd_rn((float)yb, 0.5f), c.inv_zoom);
        const int j_lo = cell_lo(__fsub_rn(bya, c.off_max_y), rm, delta);
        const int j_hi = cell_hi(__fsub_rn(byb, c.off_min_y), rm, delta);
        const long long WH = (long long)j_hi - (long long)j_lo + 1;
        bool fail = (WH < 1 || WH > RH || j_lo < cfg.bm_j0 || (long long)j_hi >= (long long)cfg.bm_j0 + cfg.bm_rows);
        if (!fail) {
            if (first) { j_gen = j_lo; rr_gen = 0; rr_lo = 0; first = false; }
            else {
                long long adv = (long long)j_lo - j_lo_prev; // >= 0 (monotone in y)
                if (adv >= RH || j_lo >= j_gen) { j_gen = max(j_gen, j_lo); rr_lo = rr_gen; if (j_gen > j_lo) fail = true; }
                else { rr_lo += (int)adv; if (rr_lo >= RH) rr_lo -= RH; }
            }
            j_lo_prev = j_lo;
        }
        if (fail) {
            if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, ya, X1 - X0 + 1, Y1 - ya, plane);
            return;
        }
        uint32_t used = 0u; // grains of the live rows: ring distance from the oldest live row's first grain
        if (j_lo < j_gen) {
            const uint32_t tail = P[rr_lo * PS];
            used = head - tail + (head < tail ? GC : 0u);
        }

        // =================== STAGED: load cell rows j_gen .. j_hi from the cell table ===================
        // Rows are placed LINEARLY in the grain ring: a row that would cross the end of the ring starts
        // at 0 instead (the tail is skipped), so within a row end >= start and the evaluation needs no
        // wrap arithmetic.  The two first prefix rows are mirrored behind the last one for the same reason.
        if (STAGED && j_gen <= j_hi) {
            const int nn = j_hi - j_gen + 1; // <= RH
            uint32_t* rowN = rowA + RH;
            uint32_t* rowH = rowN + RH;
            const size_t trow0 = (size_t)plane * cfg.bm_rows + (size_t)(j_gen - cfg.bm_j0);
            const uint32_t tcol0 = (uint32_t)(i_lo - cfg.bm_i0);
            const bool have_pf = (pf_j0 == j_gen && pf_n >= nn); // uniform
            // Every warp computes the placement of ALL new rows itself (same loads, same scan, same
            // verdict) and keeps the entries of the rows it copies: no CTA barrier, no exchange.
            uint32_t nhead;
            {
                const bool live = j_lo < j_gen;             // rows of earlier steps still in the window
                uint32_t cur = live ? head : 0u;            // empty ring: restart at 0
                const uint32_t tail = live ? (uint32_t)P[rr_lo * PS + sh] : 0u;
                const bool wrapped0 = cur < tail;           // live region already wraps around the end
                bool wrapped = false, bad = false;
                for (int t0 = 0; t0 < nn; t0 += 32) {
                    uint32_t a = 0, v = 0;
                    if (t0 + lane < nn) {
                        if (have_pf) { // fetched into shared memory while the previous step evaluated
                            a = extb[4 * (t0 + lane)];
                            v = extb[4 * (t0 + lane) + 1] - a;
                        } else {
                            const uint32_t* pr = tab.Pg + (trow0 + t0 + lane) * cfg.ppitch + tcol0;
                            a = __ldg(pr);
                            v = __ldg(pr + CW) - a;
                        }
                    }
                    uint32_t incl = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= d) incl += u;
                    }
                    const uint32_t excl = incl - v, tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    uint32_t pos = cur + excl;
                    const uint32_t over = __ballot_sync(0xFFFFFFFFu, pos + v > GC);
                    if (over) { // the first row that does not fit before the end restarts at 0
                        const int f = __ffs(over) - 1;
                        const uint32_t eb = __shfl_sync(0xFFFFFFFFu, excl, f);
                        if (lane >= f) pos = excl - eb;
                        if (wrapped || wrapped0) bad = true; // second wrap: cannot fit
                        wrapped = true;
                        cur = tot - eb;
                        if (cur > GC) bad = true;
                    } else {
                        cur += tot;
                    }
                    // a row's entry has one writer: the warp that copies that row below
                    if (t0 + lane < nn && (t0 + lane) % FG_TILE_WARPS == warp) { rowA[t0 + lane] = a; rowN[t0 + lane] = v; rowH[t0 + lane] = pos; }
                }
                // the new rows must end strictly before the oldest live grain
                if ((wrapped || wrapped0) && cur >= tail) bad = true;
                nhead = bad ? 0xFFFFFFFFu : cur;
            }
            if (nhead == 0xFFFFFFFFu) { // uniform: the window does not fit the grain ring
                if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, ya, X1 - X0 + 1, Y1 - ya, plane);
                return;
            }
            __syncwarp();
            for (int t = warp; t < nn; t += FG_TILE_WARPS) {
                const uint32_t a = rowA[t], n = rowN[t], h = rowH[t];
                int rr = rr_gen + t;
                if (rr >= RH) rr -= RH;
                const uint4* pr4 = (const uint4*)(tab.Pg + (trow0 + t) * cfg.ppitch + (tcol0 - (uint32_t)sh));
                const size_t gsrc = (have_pf ? (size_t)(((uint64_t)extb[4 * t + 3] << 32) | extb[4 * t + 2]) : (size_t)__ldg(tab.rowbase + trow0 + t)) + a;
                // the row's grain slice goes straight to the ring, asynchronously, while the prefix row is
                // loaded and converted below
                for (uint32_t kk = lane; kk < n; kk += 32) {
                    cp_async8(G + h + kk, tab.Gg + gsrc + kk);
                    if (LOGN) cp_async4(R2 + h + kk, tab.R2g + gsrc + kk);
                }
                uint2* prow = (uint2*)(P + rr * PS);
                uint2* pmir = (uint2*)(P + (rr + RH) * PS); // rows 0 and 1 again behind row RH - 1
                const int n4 = (CW1 + sh + 3) >> 2;
                const uint32_t adj = h - a;
                // loads first (two 16-byte loads in flight per lane), then the stores
                for (int c0 = 0; c0 < n4; c0 += 64) {
                    uint4 pv[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int q4 = c0 + 32 * k + lane;
                        if (q4 < n4) pv[k] = __ldg(pr4 + q4);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int q4 = c0 + 32 * k + lane;
                        if (q4 < n4) {
                            uint2 o; // entries left of the window (index < sh) are never read
                            o.x = ((pv[k].x + adj) & 0xFFFFu) | ((pv[k].y + adj) << 16);
                            o.y = ((pv[k].z + adj) & 0xFFFFu) | ((pv[k].w + adj) << 16);
                            prow[q4] = o;
                            if (rr < 2) pmir[q4] = o;
                        }
                    }
                }
            }
            cp_async_wait_all();
            head = nhead;
            j_gen += nn;
            rr_gen += nn;
            if (rr_gen >= RH) rr_gen -= RH;
        }
        // =================== generation: cell rows j_gen .. j_hi in groups of <= R rows ===================
        while (!STAGED && j_gen <= j_hi) {
            const int nr = min(r_cur, j_hi - j_gen + 1);
            if (j_bm != j_gen) load_bm(j_gen); // first group, or the window jumped (uniform)
            // ---- phase A: first-draw filter on every cell of the group (4 independent hash chains
            //      per thread are issued back to back before the ballots) ----
            uint32_t masks[FG_TILE_ITERS];
#pragma unroll
            for (int it = 0; it < FG_TILE_ITERS; ++it) {
                const int rc = rci[it];
                // one bit per cell from the first-draw bitmap (k_first_draw_bitmap); the words were
                // fetched while the previous group was in its dense phase
                const bool ne = (rc >= 0 && (rc >> 16) < nr && (rc & 0xFFFF) < CW) && ((bmw[it] >> bsh[it]) & 1u);
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, ne);
                masks[it] = m;
                if (lane == 0) cntA[it * FG_TILE_WARPS + warp] = __popc(m);
            }
            load_bm(j_gen + nr); // prefetch for the next group (its rows start at j_gen + nr)
            __syncthreads();
            // exclusive scan of the FG_TILE_NE (it, warp) counts, redundantly in every warp (no extra
            // barrier): lane l holds entries K*l .. K*l+K-1; the offsets this thread needs come back by shuffle
            uint32_t offs[FG_TILE_ITERS];
            uint32_t M;
            {
                constexpr int K = (FG_TILE_NE + 31) / 32;
                static_assert(K <= 4, "at most four counters per lane");
                uint32_t cv[K];
                uint32_t s = 0;
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    cv[q] = (K * lane + q < FG_TILE_NE) ? cntA[K * lane + q] : 0u;
                    s += cv[q];
                }
                uint32_t incl = s;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += v;
                }
                const uint32_t excl = incl - s;
                M = __shfl_sync(0xFFFFFFFFu, incl, 31);
#pragma unroll
                for (int it = 0; it < FG_TILE_ITERS; ++it) {
                    const int e = it * FG_TILE_WARPS + warp;
                    const int src = e / K, sub = e - K * src;
                    uint32_t o = __shfl_sync(0xFFFFFFFFu, excl, src);
#pragma unroll
                    for (int q = 0; q < K - 1; ++q) {
                        const uint32_t v = __shfl_sync(0xFFFFFFFFu, cv[q], src);
                        if (q < sub) o += v;
                    }
                    offs[it] = o;
                }
            }
#pragma unroll
            for (int it = 0; it < FG_TILE_ITERS; ++it) {
                if ((masks[it] >> lane) & 1u) {
                    const uint32_t pos = offs[it] + __popc(masks[it] & lt_mask);
                    list[pos] = (uint16_t)(((rci[it] >> 16) << 12) | (rci[it] & 0xFFF));
                }
            }
            __syncthreads();
            // ---- phase B (dense): full seeding, Knuth continuation, positions ----
            for (uint32_t base = 0; base < M; base += FG_TILE_THREADS) {
                const uint32_t t = base + tid;
                uint32_t q = 0;
                Xoshiro rng;
                float sx = 0.0f, sy = 0.0f;
                if (t < M) {
                    const uint32_t item = list[t];
                    const int il = item & 0xFFF, j = j_gen + (int)(item >> 12);
                    const ColInfo ci = colT[il];
                    sx = ci.sx;
                    sy = __fmul_rn(__int2float_rn(j), delta);
                    int iy = floor_i32(sy);
                    iy = min(max(iy, 0), c.in_h - 1);
                    const double e = __ldg(ev + (size_t)iy * c.in_w + ci.ixc);
                    seed_small_rng(rng, mix3_row(ci.h, j), c.seeding);
                    if (e < 0.0) {
                        wtot[FG_TILE_WARPS] = 1; // lambda' >= 12 / non-finite: general path only (flag read after the barrier)
                    } else {
                        double p = standard_f64(rng); // first draw (known > e from the bitmap)
                        while (p > e) { p = __dmul_rn(p, standard_f64(rng)); ++q; }
                    }
                }
                // block-wide exclusive scan of q
                uint32_t incl = q;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += v;
                }
                if (lane == 31) wtot[warp] = incl;
                __syncthreads();
                // warp totals -> this warp's offset and the pass total: one shuffle scan per warp
                uint32_t boff, total;
                {
                    const uint32_t v = (lane < FG_TILE_WARPS) ? wtot[lane] : 0u;
                    uint32_t wincl = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, wincl, d);
                        if (lane >= d) wincl += u;
                    }
                    total = __shfl_sync(0xFFFFFFFFu, wincl, FG_TILE_WARPS - 1);
                    boff = __shfl_sync(0xFFFFFFFFu, wincl - v, warp);
                }
                if (used + total >= GC || wtot[FG_TILE_WARPS]) { // uniform: grain ring overflow (== GC would alias an empty ring) or a general-path cell
                    if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, ya, X1 - X0 + 1, Y1 - ya, plane);
                    return;
                }
                if (t < M) {
                    uint32_t pos = head + boff + (incl - q);
                    if (pos >= GC) pos -= GC;
                    E[t] = (uint16_t)pos;
                    for (uint32_t g = 0; g < q; ++g) {
                        float cx = __fadd_rn(sx, uniform_f32(rng, c.uscale_cell));
                        float cy = __fadd_rn(sy, uniform_f32(rng, c.uscale_cell));
                        G[pos] = make_float2(cx, cy);
                        if (pos < FG_TILE_GPAD) G[GC + pos] = make_float2(cx, cy); // mirror: unrolled reads never wrap
                        if (LOGN) { // RadiusProfile::sample + clamp (src/model.rs:137-148, src/pixelwise.rs:89-95)
                            const float radius = radius_sample_clamped(c.rad, rng);
                            const float rr2 = radius > 0.0f ? __fmul_rn(radius, radius) : -1.0f;
                            R2[pos] = rr2;
                            if (pos < FG_TILE_GPAD) R2[GC + pos] = rr2;
                        }
                        if (++pos == GC) pos = 0;
                    }
                }
                head += total;
                if (head >= GC) head -= GC;
                used += total;
                __syncthreads(); // wtot reuse + E/G visible
            }
            // ---- P fill: ring position of the first grain at or after each cell ----
#pragma unroll
            for (int it = 0; it < FG_TILE_ITERS; ++it) {
                const int rc = rci[it];
                if (rc >= 0 && (rc >> 16) < nr) {
                    const uint32_t rank = offs[it] + __popc(masks[it] & lt_mask);
                    int rr = rr_gen + (rc >> 16);
                    if (rr >= RH) rr -= RH;
                    P[rr * PS + (rc & 0xFFFF)] = (rank == M) ? (uint16_t)head : E[rank];
                }
            }
            j_gen += nr;
            rr_gen += nr;
            if (rr_gen >= RH) rr_gen -= RH;
            // next group: as many rows as keep the dense pass within one sweep of the CTA
            if (M > (uint32_t)(FG_TILE_THREADS - 16)) r_cur = max(1, nr - 1);
            else if (M * (uint32_t)(nr + 1) <= (uint32_t)(FG_TILE_THREADS - 32) * (uint32_t)nr) r_cur = min(cfg.R, nr + 1);
            else r_cur = nr;
            // no barrier here: the next group only touches cntA/list/E behind its own barriers
        }
        __syncthreads(); // P and G complete before the evaluation reads them

        // STAGED: while this step evaluates, fetch what the next step's placement needs (first/last
        // prefix of each new row's window slice and the row's base) into shared memory
        if (STAGED) {
            pf_n = 0;
            const int ya2 = ya + cfg.TH;
            if (ya2 < Y1) {
                const int yb2 = min(ya2 + cfg.TH, Y1) - 1;
                const float bya2 = __fmul_rn(__fadd_rn((float)ya2, 0.5f), c.inv_zoom);
                const float byb2 = __fmul_rn(__fadd_rn((float)yb2, 0.5f), c.inv_zoom);
                const int j_lo2 = cell_lo(__fsub_rn(bya2, c.off_max_y), rm, delta);
                const int j_hi2 = cell_hi(__fsub_rn(byb2, c.off_min_y), rm, delta);
                const long long nn2 = (long long)j_hi2 - j_gen + 1;
                if (j_lo2 <= j_gen && nn2 > 0 && nn2 <= RH && (long long)j_hi2 < (long long)cfg.bm_j0 + cfg.bm_rows) {
                    pf_j0 = j_gen;
                    pf_n = (int)nn2;
                    const size_t trow2 = (size_t)plane * cfg.bm_rows + (size_t)(j_gen - cfg.bm_j0);
                    for (int t = tid; t < pf_n; t += FG_TILE_THREADS) {
                        const uint32_t* pr = tab.Pg + (trow2 + t) * cfg.ppitch + (uint32_t)(i_lo - cfg.bm_i0);
                        cp_async4(extb + 4 * t, pr);
                        cp_async4(extb + 4 * t + 1, pr + CW);
                        cp_async8(extb + 4 * t + 2, tab.rowbase + trow2 + t);
                    }
                }
            }
        }

        // =================== evaluation of pixel rows ya..yb ===================
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            if (n_chunks > 1 || !xk_loaded) { load_xk(chunk); xk_loaded = true; }
            // per-(row, sample) data of this warp: sample point y, byte offset of its first cell row in P
            // (bits 0-23) and the number of cell rows (bits 24-31)
            float2* wp = wpair + warp * (cfg.TH * SPWC);
            for (int qd = lane; qd < th * SPWC; qd += 32) {
                const int s = qd / th, yl = qd - s * th;
                const uint32_t k = (uint32_t)chunk * (FG_TILE_WARPS * SPWC) + s * FG_TILE_WARPS + warp;
                float yg = 0.0f;
                uint32_t jp = 0;
                if (k < c.n) {
                    const float by = __fmul_rn(__fadd_rn((float)(ya + yl), 0.5f), c.inv_zoom);
                    yg = __fsub_rn(by, __ldg(offsets_input + k).y);
                    const int j0 = cell_lo(yg, rm, delta), j1 = cell_hi(yg, rm, delta);
                    if (j0 <= j1) {
                        int rr = rr_lo + (j0 - j_lo);
                        if (rr >= RH) rr -= RH;
                        jp = (uint32_t)(rr * PS * 2) | ((uint32_t)(j1 - j0 + 1) << 24);
                    }
                }
                wp[yl * SPWC + s] = make_float2(yg, __uint_as_float(jp));
            }
            __syncwarp();
            if (xvalid && radius_ok) {
                // shared-window addresses, pinned in registers (left to itself the compiler re-derives
                // them from the CTA id and the parameter block for every sample pair)
                uint32_t Ps = (uint32_t)__cvta_generic_to_shared(P), Gs = (uint32_t)__cvta_generic_to_shared(G);
                uint32_t wps = (uint32_t)__cvta_generic_to_shared(wp);
                asm volatile("" : "+r"(Ps), "+r"(Gs), "+r"(wps));
                const uint32_t PS2 = (uint32_t)PS * 2u, RHPS2 = (uint32_t)RH * PS2;
                uint64_t chA[2] = {0ull, 0ull}, chB[2] = {0ull, 0ull}; // slot_test load chains (any defined value)
                if (LOGN) {
                    // per-grain radii: one sample at a time, per cell row FG_TILE_USLOTS predicated tests
                    // against the grain's own r^2, then an early-exit remainder loop
                    const uint32_t Rs = (uint32_t)__cvta_generic_to_shared(R2);
                    for (int yl = 0; yl < th; ++yl) {
                        uint32_t cnt = 0;
#pragma unroll
                        for (int s = 0; s < SPWC; ++s) {
                            const float2 pd = wp[yl * SPWC + s];
                            const uint32_t jp = __float_as_uint(pd.y);
                            const uint32_t a2 = ip_r[s] & 0xFFFFu, b2 = ip_r[s] >> 16;
                            const uint32_t nrow = (a2 != b2) ? (jp >> 24) : 0u;
                            uint32_t off = Ps + (jp & 0xFFFFFFu);
                            const float xg = xg_r[s], yg = pd.x;
                            uint32_t covered = 0;
#pragma unroll 1
                            for (uint32_t r = 0; r < nrow && !covered; ++r) {
                                const uint32_t s16 = lds_u16(off + a2), e16 = lds_u16(off + b2);
                                const uint32_t n = e16 - s16 + (e16 < s16 ? GC : 0u);
#pragma unroll
                                for (int u = 0; u < FG_TILE_USLOTS; ++u) {
                                    const float2 gr = lds_f32x2(Gs + (s16 + u) * 8u); // past n: stale, in-bounds (mirror pad)
                                    const float rr2 = lds_f32(Rs + (s16 + u) * 4u);
                                    const float dx = __fsub_rn(xg, gr.x), dy = __fsub_rn(yg, gr.y);
                                    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                                    covered |= ((uint32_t)u < n && d2 <= rr2) ? 1u : 0u;
                                }
                                for (uint32_t u = FG_TILE_USLOTS; u < n && !covered; ++u) {
                                    uint32_t gi = s16 + u;
                                    if (gi >= GC) gi -= GC;
                                    const float2 gr = lds_f32x2(Gs + gi * 8u);
                                    const float rr2 = lds_f32(Rs + gi * 4u);
                                    const float dx = __fsub_rn(xg, gr.x), dy = __fsub_rn(yg, gr.y);
                                    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= rr2) covered = 1u;
                                }
                                off += PS2;
                                if (off >= Ps + RHPS2) off = Ps;
                            }
                            cnt += covered;
                        }
                        if (cnt) atomicAdd(&pcount[yl * 32 + lane], cnt);
                    }
                } else
                for (int yl = 0; yl < th; ++yl) {
                    uint32_t cnt = 0;
                    // two samples (A, B) advance through their cell rows together: their shared-memory
                    // loads are issued stage by stage (prefixes of A and B, then grains of A and B, then
                    // the arithmetic), which doubles the independent work in flight per warp.
#pragma unroll
                    for (int s = 0; s < SPWC; s += 2) {
                        const float2 pdA = lds_f32x2(wps + (uint32_t)(yl * SPWC + s) * 8u), pdB = lds_f32x2(wps + (uint32_t)(yl * SPWC + s + 1) * 8u);
                        const uint32_t jpA = __float_as_uint(pdA.y), jpB = __float_as_uint(pdB.y);
                        const uint32_t a2A = ip_r[s] & 0xFFFFu, b2A = ip_r[s] >> 16;
                        const uint32_t a2B = ip_r[s + 1] & 0xFFFFu, b2B = ip_r[s + 1] >> 16;
                        uint32_t offA = Ps + (jpA & 0xFFFFFFu), offB = Ps + (jpB & 0xFFFFFFu);
                        const float xgA = xg_r[s], ygA = pdA.x, xgB = xg_r[s + 1], ygB = pdB.x;
                        float dminA = __int_as_float(0x7f800000), dminB = __int_as_float(0x7f800000);
                        if ((jpA >> 24) == 3u && (jpB >> 24) == 3u) { // warp-uniform: the common case rm == delta
                            // Three cell rows per sample: their three candidate ranges are walked as ONE
                            // list of n = n0 + n1 + n2 grains (slot u -> row by two compares), FG_TILE_U3
                            // predicated slots folded into a running minimum, then an early-exit remainder.
                            uint32_t o1A = offA + PS2, o1B = offB + PS2;
                            uint32_t o2A = o1A + PS2, o2B = o1B + PS2;
                            if (!STAGED) { // STAGED: the first two prefix rows are mirrored behind the last
                                if (o1A >= Ps + RHPS2) o1A -= RHPS2;
                                if (o1B >= Ps + RHPS2) o1B -= RHPS2;
                                o2A = o1A + PS2; o2B = o1B + PS2;
                                if (o2A >= Ps + RHPS2) o2A -= RHPS2;
                                if (o2B >= Ps + RHPS2) o2B -= RHPS2;
                            }
                            const uint32_t s0A = lds_u16(offA + a2A), e0A = lds_u16(offA + b2A);
                            const uint32_t s0B = lds_u16(offB + a2B), e0B = lds_u16(offB + b2B);
                            const uint32_t s1A = lds_u16(o1A + a2A), e1A = lds_u16(o1A + b2A);
                            const uint32_t s1B = lds_u16(o1B + a2B), e1B = lds_u16(o1B + b2B);
                            const uint32_t s2A = lds_u16(o2A + a2A), e2A = lds_u16(o2A + b2A);
                            const uint32_t s2B = lds_u16(o2B + a2B), e2B = lds_u16(o2B + b2B);
                            const uint64_t pA = pack_f32x2(xgA, ygA), pB = pack_f32x2(xgB, ygB);
#if FG_TILE_UROW > 0
                            if (STAGED) {
                                // Per cell row FG_TILE_UROW straight-line tests whose LOADS are predicated on the row's own
                                // count (slot_test): immediate-offset addresses, no slot -> row selection, and a lane never
                                // loads a grain that is not its candidate (fewer shared-memory wavefronts).  What is left of
                                // the three rows (a few % of the samples) is walked as one list by the early-exit loop below.
                                const uint32_t n0A = e0A - s0A, n1A = e1A - s1A, n2A = e2A - s2A;
                                const uint32_t n0B = e0B - s0B, n1B = e1B - s1B, n2B = e2B - s2B;
                                const uint32_t g0A = Gs + s0A * 8u, g1A = Gs + s1A * 8u, g2A = Gs + s2A * 8u;
                                const uint32_t g0B = Gs + s0B * 8u, g1B = Gs + s1B * 8u, g2B = Gs + s2B * 8u;
                                RowSlots<0, FG_TILE_UROW>::run(dminA, dminB, chA, chB, n0A, n0B, g0A, g0B, pA, pB);
                                RowSlots<0, FG_TILE_UROW>::run(dminA, dminB, chA, chB, n1A, n1B, g1A, g1B, pA, pB);
                                RowSlots<0, FG_TILE_UROW>::run(dminA, dminB, chA, chB, n2A, n2B, g2A, g2B, pA, pB);
                                constexpr uint32_t U = FG_TILE_UROW;
                                if (max(max(n0A, n1A), n2A) > U && !(dminA <= r2)) {
                                    const uint32_t m0 = n0A > U ? n0A - U : 0u, m1 = m0 + (n1A > U ? n1A - U : 0u), m = m1 + (n2A > U ? n2A - U : 0u);
                                    const uint32_t b0 = g0A + U * 8u, b1 = g1A + U * 8u - m0 * 8u, b2 = g2A + U * 8u - m1 * 8u;
                                    uint32_t u = 0;
                                    do {
                                        const float d2 = dist2_packed(pA, lds_f32x2((u < m0 ? b0 : (u < m1 ? b1 : b2)) + u * 8u));
                                        if (d2 <= r2) { dminA = d2; break; }
                                    } while (++u < m);
                                }
                                if (max(max(n0B, n1B), n2B) > U && !(dminB <= r2)) {
                                    const uint32_t m0 = n0B > U ? n0B - U : 0u, m1 = m0 + (n1B > U ? n1B - U : 0u), m = m1 + (n2B > U ? n2B - U : 0u);
                                    const uint32_t b0 = g0B + U * 8u, b1 = g1B + U * 8u - m0 * 8u, b2 = g2B + U * 8u - m1 * 8u;
                                    uint32_t u = 0;
                                    do {
                                        const float d2 = dist2_packed(pB, lds_f32x2((u < m0 ? b0 : (u < m1 ? b1 : b2)) + u * 8u));
                                        if (d2 <= r2) { dminB = d2; break; }
                                    } while (++u < m);
                                }
                            } else
#endif
                            {
                            // STAGED rows never wrap inside the ring: end >= start
                            const uint32_t c1A = e0A - s0A + ((!STAGED && e0A < s0A) ? GC : 0u), c1B = e0B - s0B + ((!STAGED && e0B < s0B) ? GC : 0u);
                            const uint32_t c2A = c1A + e1A - s1A + ((!STAGED && e1A < s1A) ? GC : 0u), c2B = c1B + e1B - s1B + ((!STAGED && e1B < s1B) ? GC : 0u);
                            const uint32_t nA = c2A + e2A - s2A + ((!STAGED && e2A < s2A) ? GC : 0u), nB = c2B + e2B - s2B + ((!STAGED && e2B < s2B) ? GC : 0u);
                            const uint32_t t1A = s1A - c1A, t2A = s2A - c2A, t1B = s1B - c1B, t2B = s2B - c2B; // mod 2^32
                            float2 gA[FG_TILE_U3], gB[FG_TILE_U3];
#pragma unroll
                            for (int u = 0; u < FG_TILE_U3; ++u) { // past n: stale but in-bounds (pad behind the ring)
                                const uint32_t iA = ((uint32_t)u < c1A ? s0A : ((uint32_t)u < c2A ? t1A : t2A)) + (uint32_t)u;
                                const uint32_t iB = ((uint32_t)u < c1B ? s0B : ((uint32_t)u < c2B ? t1B : t2B)) + (uint32_t)u;
                                gA[u] = lds_f32x2(Gs + iA * 8u);
                                gB[u] = lds_f32x2(Gs + iB * 8u);
                            }
#pragma unroll
                            for (int u = 0; u < FG_TILE_U3; ++u) {
                                const float dA = dist2_packed(pA, gA[u]), dB = dist2_packed(pB, gB[u]);
                                dminA = fminf(dminA, ((uint32_t)u < nA) ? dA : __int_as_float(0x7f800000));
                                dminB = fminf(dminB, ((uint32_t)u < nB) ? dB : __int_as_float(0x7f800000));
                            }
                            if (nA > FG_TILE_U3 && !(dminA <= r2)) { // remainder: exits on the first hit
                                uint32_t u = FG_TILE_U3;
                                do {
                                    uint32_t gi = (u < c1A ? s0A : (u < c2A ? t1A : t2A)) + u;
                                    if (!STAGED && gi >= GC) gi -= GC;
                                    const float d2 = dist2_packed(pA, lds_f32x2(Gs + gi * 8u));
                                    if (d2 <= r2) { dminA = d2; break; }
                                } while (++u < nA);
                            }
                            if (nB > FG_TILE_U3 && !(dminB <= r2)) {
                                uint32_t u = FG_TILE_U3;
                                do {
                                    uint32_t gi = (u < c1B ? s0B : (u < c2B ? t1B : t2B)) + u;
                                    if (!STAGED && gi >= GC) gi -= GC;
                                    const float d2 = dist2_packed(pB, lds_f32x2(Gs + gi * 8u));
                                    if (d2 <= r2) { dminB = d2; break; }
                                } while (++u < nB);
                            }
                            }
                        } else {
                        const uint32_t nrowA = (a2A != b2A) ? (jpA >> 24) : 0u, nrowB = (a2B != b2B) ? (jpB >> 24) : 0u;
                        const uint32_t nmax = max(nrowA, nrowB);
#pragma unroll 1
                        for (uint32_t r = 0; r < nmax; ++r) {
                            // a sample that has run out of cell rows reads an empty range (b := a)
                            const uint32_t bA = (r < nrowA) ? b2A : a2A, bB = (r < nrowB) ? b2B : a2B;
                            const uint32_t sA = lds_u16(offA + a2A), eA = lds_u16(offA + bA);
                            const uint32_t sB = lds_u16(offB + a2B), eB = lds_u16(offB + bB);
                            const uint32_t nA = eA - sA + (eA < sA ? GC : 0u), nB = eB - sB + (eB < sB ? GC : 0u);
                            const uint32_t gaA = Gs + sA * 8u, gaB = Gs + sB * 8u;
                            float2 gA[FG_TILE_USLOTS], gB[FG_TILE_USLOTS];
#pragma unroll
                            for (int u = 0; u < FG_TILE_USLOTS; ++u) gA[u] = lds_f32x2(gaA + 8u * u); // past n: stale, in-bounds
#pragma unroll
                            for (int u = 0; u < FG_TILE_USLOTS; ++u) gB[u] = lds_f32x2(gaB + 8u * u);
#pragma unroll
                            for (int u = 0; u < FG_TILE_USLOTS; ++u) {
                                const float gxA = ((uint32_t)u < nA) ? gA[u].x : __int_as_float(0x7f800000);
                                const float gxB = ((uint32_t)u < nB) ? gB[u].x : __int_as_float(0x7f800000);
                                const float dxA = __fsub_rn(xgA, gxA), dyA = __fsub_rn(ygA, gA[u].y);
                                const float dxB = __fsub_rn(xgB, gxB), dyB = __fsub_rn(ygB, gB[u].y);
                                dminA = fminf(dminA, __fadd_rn(__fmul_rn(dxA, dxA), __fmul_rn(dyA, dyA)));
                                dminB = fminf(dminB, __fadd_rn(__fmul_rn(dxB, dxB), __fmul_rn(dyB, dyB)));
                            }
                            if (nA > FG_TILE_USLOTS && !(dminA <= r2)) { // remainder: exits on the first hit
                                uint32_t u = FG_TILE_USLOTS;
                                do {
                                    uint32_t gi = sA + u;
                                    if (gi >= GC) gi -= GC;
                                    const float2 gr = lds_f32x2(Gs + gi * 8u);
                                    const float dx = __fsub_rn(xgA, gr.x), dy = __fsub_rn(ygA, gr.y);
                                    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                                    if (d2 <= r2) { dminA = d2; break; }
                                } while (++u < nA);
                            }
                            if (nB > FG_TILE_USLOTS && !(dminB <= r2)) {
                                uint32_t u = FG_TILE_USLOTS;
                                do {
                                    uint32_t gi = sB + u;
                                    if (gi >= GC) gi -= GC;
                                    const float2 gr = lds_f32x2(Gs + gi * 8u);
                                    const float dx = __fsub_rn(xgB, gr.x), dy = __fsub_rn(ygB, gr.y);
                                    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                                    if (d2 <= r2) { dminB = d2; break; }
                                } while (++u < nB);
                            }
                            // advance only inside the sample's own rows: a sample that has run out of rows re-reads
                            // its last one (as an empty range).  The rows behind it may never have been loaded, and
                            // an arbitrary prefix value would send the unconditional slot loads outside the ring
                            if (r + 1u < nrowA) { offA += PS2; if (offA >= Ps + RHPS2) offA = Ps; }
                            if (r + 1u < nrowB) { offB += PS2; if (offB >= Ps + RHPS2) offB = Ps; }
                        }
                        }
                        cnt += ((dminA <= r2) ? 1u : 0u) + ((dminB <= r2) ? 1u : 0u);
                    }
                    if (cnt) atomicAdd(&pcount[yl * 32 + lane], cnt);
                }
            }
            __syncwarp();
        }
        if (STAGED) cp_async_wait_all(); // the extents prefetched above (visible to all after the barrier)
        __syncthreads();
        for (int p = tid; p < th * 32; p += FG_TILE_THREADS) {
            const int yl = p >> 5, xl = p & 31;
            if (X0 + xl <= X1) outp[(size_t)(ya + yl) * c.out_w + X0 + xl] = __fmul_rn((float)pcount[p], c.inv_samples);
            pcount[p] = 0;
        }
        // STAGED: the loader of the next step touches neither pcount nor wpair, and the barrier that
        // ends it orders these stores before the next evaluation
        if (!STAGED) __syncthreads();
    }
}

// Fallback of the STAGED strip kernel: strip segments whose window does not fit the shared-memory
// rings (dense content: more grains per window than the grain ring holds) are evaluated straight from
// the HBM cell table -- the same candidate ranges, read through L2 instead of shared memory -- rather
// than by regenerating every cell per sample.  Work item = one 32 x 8 pixel chunk of one listed tile.
// A sample whose cells leave the table rectangle (cannot happen for planned geometry) takes the
// regenerating indicator.
template <bool LOGN>
__global__ void __launch_bounds__(256) k_pixelwise_table_tiles(const float* __restrict__ lambda, size_t lambda_stride,
                                                                const float2* __restrict__ offsets_input,
                                                                float* __restrict__ out, size_t out_stride,
                                                                const TileRef* __restrict__ tiles,
                                                                const uint32_t* __restrict__ n_tiles, uint32_t tile_cap,
                                                                uint32_t chunks_per_tile, uint32_t* __restrict__ n_total,
                                                                TileCfg cfg, RenderConsts c, CellTable tab) {
    const uint32_t nt = min(*n_tiles, tile_cap);
    if (n_total && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_total, nt);
    const uint64_t work = (uint64_t)nt * chunks_per_tile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float rm = c.rad.rm, delta = c.delta, r2c = cfg.r2c;
    const bool radius_ok = LOGN || (c.rad.mean_linear > rm ? rm : c.rad.mean_linear) > 0.0f; // the radius itself: r2c is positive for r < 0
    for (uint64_t wi = blockIdx.x; wi < work; wi += gridDim.x) {
        if (c.abort != nullptr && *(const volatile int*)c.abort != 0) return; // cancelled (no barriers below: threads may leave alone)
        const TileRef t = tiles[wi / chunks_per_tile];
        const int yl = (int)(wi % chunks_per_tile) * 8 + ty;
        if (yl >= t.h || tx >= t.w) continue;
        const int x = t.x0 + tx, y = t.y0 + yl;
        if (x >= c.out_w || y >= c.row_end) continue;
        const float* lam = lambda + lambda_stride * t.plane;
        const size_t prow0 = (size_t)t.plane * cfg.bm_rows;
        const float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);
        const float by = __fmul_rn(__fadd_rn((float)y, 0.5f), c.inv_zoom);
        uint32_t count = 0;
        for (uint32_t k = 0; k < c.n; ++k) {
            const float2 o = __ldg(offsets_input + k);
            const float xg = __fsub_rn(bx, o.x), yg = __fsub_rn(by, o.y);
            if (!(rm > 0.0f)) continue;
            const int i0 = cell_lo(xg, rm, delta), i1 = cell_hi(xg, rm, delta);
            const int j0 = cell_lo(yg, rm, delta), j1 = cell_hi(yg, rm, delta);
            if (i0 > i1 || j0 > j1) continue;
            if (i0 < cfg.bm_i0 || (long long)i1 >= (long long)cfg.bm_i0 + cfg.bm_cols || j0 < cfg.bm_j0 ||
                (long long)j1 >= (long long)cfg.bm_j0 + cfg.bm_rows) {
                count += indicator_direct(lam, c, xg, yg) ? 1u : 0u;
                continue;
            }
            if (!radius_ok) continue;
            bool hit = false;
            for (int j = j0; j <= j1 && !hit; ++j) {
                const size_t row = prow0 + (size_t)(j - cfg.bm_j0);
                const uint32_t* pr = tab.Pg + row * cfg.ppitch + (uint32_t)(i0 - cfg.bm_i0);
                const uint32_t s = __ldg(pr), e = __ldg(pr + (i1 - i0 + 1));
                const size_t base = (size_t)__ldg(tab.rowbase + row);
                for (uint32_t g = s; g < e; ++g) {
                    const float2 gr = __ldg(tab.Gg + base + g);
                    const float rr2 = LOGN ? __ldg(tab.R2g + base + g) : r2c; // LOGN: -1 = radius <= 0, never covers
                    const float dx = __fsub_rn(xg, gr.x), dy = __fsub_rn(yg, gr.y);
                    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= rr2) { hit = true; break; }
                }
            }
            count += hit ? 1u : 0u;
        }
        out[out_stride * t.plane + (size_t)y * c.out_w + x] = __fmul_rn((float)count, c.inv_samples);
    }
}

} // namespace fg
