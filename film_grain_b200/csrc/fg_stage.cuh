// fg_stage.cuh -- the CELL TABLE: every Boolean-model cell of a render band generated exactly once.
//
// Reference semantics (src/pixelwise.rs:60-104): a cell (i, j) is seeded from (seed, i, j) only, so its
// grains are the same for every pixel, sample and colour-plane-independent part of the chain that visits
// it.  The reference regenerates the cell on every visit; the strip kernel (fg_tile.cuh) used to
// regenerate it once per strip window.  Here one kernel walks the band's cell rows and writes, per
// colour plane, a CSR table in HBM:
//
//   Pg[plane][row][col]   u32, col = 0 .. cols: number of grains of the row's cells [0, col)
//   Gg[rowbase[plane,row] + Pg[...]]   float2 grain centres in cell order (+ R2g: squared radius, log-normal)
//
// so that a strip window is two contiguous slices per cell row and the strip kernel only loads and
// evaluates.  Row storage is sized from the row's expected grain count (the sum of its cells' Poisson
// means) plus eight standard deviations; a row that would not fit raises `overflow` and the caller
// falls back to in-kernel generation (never observed; P < 1e-15 per row).
#pragma once
#include "fg_kernels.cuh"

namespace fg {

#define FG_GEN_THREADS 256
#define FG_GEN_TILE 2048 // cells per generation tile = 64 first-draw bitmap words

// Expected grains of one cell row, per (plane, input row): sum over the band's cell columns of
// lambda' = lambda * delta^2 (src/pixelwise.rs:76-81) at the clamped input pixel of the column.
__global__ void __launch_bounds__(256) k_row_expect(const float* __restrict__ lambda, size_t in_stride, int i0, int cols,
                                                     double* __restrict__ S, RenderConsts c) {
    const int iy = blockIdx.x, pl = blockIdx.y;
    const float* lrow = lambda + in_stride * pl + (size_t)iy * c.in_w;
    double s = 0.0;
    for (int col = threadIdx.x; col < cols; col += 256) {
        const int ix = min(max(floor_i32(__fmul_rn(__int2float_rn(i0 + col), c.delta)), 0), c.in_w - 1);
        const float lam = __ldg(lrow + ix);
        float ex = lam > 0.0f ? __fmul_rn(__fmul_rn(lam, c.delta), c.delta) : 0.0f;
        if (!(ex > 0.0f) || !(ex < 1.0e15f)) ex = 0.0f; // poisson_f64 draws nothing for these
        s += (double)ex;
    }
    __shared__ double red[8];
    for (int d = 16; d; d >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        S[(size_t)pl * c.in_h + iy] = t;
    }
}

// Row capacities and their exclusive prefix (one CTA of 1024 threads; n = planes * rows is ~1e5).
// rowbase[n] = total entries; UINT64_MAX if a row is too large for 32-bit in-row prefixes.
__global__ void __launch_bounds__(1024) k_row_bases(const double* __restrict__ S, int j0, int rows, int n_planes,
                                                     uint64_t* __restrict__ rowbase, uint32_t* __restrict__ rowcap, RenderConsts c) {
    __shared__ uint64_t wsum[32];
    __shared__ int bad;
    const int n = rows * n_planes, per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(b + per, n);
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    auto cap_of = [&](int r) -> uint32_t {
        const int pl = r / rows, row = r - pl * rows;
        const int iy = min(max(floor_i32(__fmul_rn(__int2float_rn(j0 + row), c.delta)), 0), c.in_h - 1);
        const double s = S[(size_t)pl * c.in_h + iy];
        const double cap = s + 8.0 * sqrt(s) + 64.0;
        if (!(cap < 3.0e9)) { bad = 1; return 0u; }
        return (uint32_t)cap;
    };
    uint64_t sum = 0;
    for (int r = b; r < e; ++r) sum += cap_of(r);
    uint64_t incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = wsum[lane], wi = w;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += v;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    uint64_t run = wsum[warp] + incl - sum;
    for (int r = b; r < e; ++r) {
        const uint32_t cp = cap_of(r);
        rowbase[r] = run;
        rowcap[r] = cp;
        run += cp;
    }
    if (threadIdx.x == 1023) rowbase[n] = bad ? 0xFFFFFFFFFFFFFFFFULL : run;
    __syncthreads();
    if (bad && threadIdx.x == 0) rowbase[n] = 0xFFFFFFFFFFFFFFFFULL;
}

struct StageGeo {
    int i0, j0, cols, rows;   // cell rectangle of the band (same as the first-draw bitmap)
    uint32_t pitchw;          // bitmap words per row
    uint32_t ppitch;          // Pg entries per row (>= cols + 1, multiple of 8)
};

// One CTA per (cell row, plane).  The row is walked in tiles of FG_GEN_TILE cells: the tile's non-empty
// cells (first-draw bitmap) are compacted into a list, processed one thread per cell at full occupancy
// (seeding, Knuth continuation or the general Poisson sampler, positions, radii), and written in cell
// order; a final scan over the tile's counts writes the prefix entries of ALL its cells.
template <bool LOGN>
__global__ void __launch_bounds__(FG_GEN_THREADS) k_gen_rows(const uint32_t* __restrict__ bm_planes, size_t bm_plane_words,
                                                              const double* __restrict__ e_planes, const float* __restrict__ lambda,
                                                              size_t in_stride, uint32_t* __restrict__ Pg,
                                                              const uint64_t* __restrict__ rowbase, const uint32_t* __restrict__ rowcap,
                                                              float2* __restrict__ Gg, float* __restrict__ R2g,
                                                              uint32_t* __restrict__ overflow, StageGeo geo, RenderConsts c) {
    __shared__ uint16_t list[FG_GEN_TILE];
    __shared__ __align__(16) uint32_t cnt[FG_GEN_TILE];
    __shared__ uint32_t wsA[8], wsB[2][8], wsC[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x, plane = blockIdx.y;
    const int j = geo.j0 + row;
    const float sy = __fmul_rn(__int2float_rn(j), c.delta);
    const int iy = min(max(floor_i32(sy), 0), c.in_h - 1);
    const double* erow = e_planes + in_stride * plane + (size_t)iy * c.in_w;
    const float* lrow = lambda + in_stride * plane + (size_t)iy * c.in_w;
    const uint32_t* bmrow = bm_planes + bm_plane_words * plane + (size_t)row * geo.pitchw;
    const size_t ridx = (size_t)plane * geo.rows + row;
    const uint64_t base = rowbase[ridx];
    const uint32_t cap = rowcap[ridx];
    uint32_t* prow = Pg + ridx * geo.ppitch;
    uint32_t run = 0; // grains of the row written so far
    int batch_par = 0;

    auto warp_incl = [&](uint32_t v) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v += u;
        }
        return v;
    };
    // sum and exclusive offset of this warp from eight warp totals in shared memory
    auto block_off = [&](const uint32_t* ws, uint32_t& total) {
        const uint32_t v = (lane < 8) ? ws[lane] : 0u;
        const uint32_t wi = warp_incl(v);
        total = __shfl_sync(0xFFFFFFFFu, wi, 7);
        return __shfl_sync(0xFFFFFFFFu, wi - v, warp);
    };

    for (int tile0 = 0; tile0 < (int)geo.ppitch; tile0 += FG_GEN_TILE) {
        // ---- a: clear the counts, compact the tile's non-empty cells into `list` ----
        {
            uint4* cz = (uint4*)cnt;
            cz[tid] = make_uint4(0, 0, 0, 0);
            cz[tid + FG_GEN_THREADS] = make_uint4(0, 0, 0, 0);
        }
        const uint32_t widx = (uint32_t)(tile0 >> 5) + (uint32_t)(tid >> 2), sh = 8u * (tid & 3);
        const uint32_t word = widx < geo.pitchw ? __ldg(bmrow + widx) : 0u;
        uint32_t bits = (word >> sh) & 0xFFu;
        const uint32_t nb = __popc(bits);
        const uint32_t inclA = warp_incl(nb);
        if (lane == 31) wsA[warp] = inclA;
        __syncthreads();
        uint32_t M;
        uint32_t pos = block_off(wsA, M) + inclA - nb;
        while (bits) {
            const int b0 = __ffs(bits) - 1;
            bits &= bits - 1;
            list[pos++] = (uint16_t)((tid >> 2) * 32 + (int)sh + b0);
        }
        __syncthreads();
        // ---- b: dense batches, one thread per non-empty cell ----
        for (uint32_t b0 = 0; b0 < M; b0 += FG_GEN_THREADS) {
            const uint32_t t = b0 + tid;
            uint32_t q = 0, cl = 0;
            Xoshiro rng;
            float sx = 0.0f;
            if (t < M) {
                cl = list[t];
                const int i = geo.i0 + tile0 + (int)cl;
                sx = __fmul_rn(__int2float_rn(i), c.delta);
                const int ix = min(max(floor_i32(sx), 0), c.in_w - 1);
                const double e = __ldg(erow + ix);
                seed_small_rng(rng, mix3_row(mix3_col(c.seed_cell, i), j), c.seeding);
                if (e < 0.0) { // lambda' >= 12 or non-finite: the general sampler (src/pixelwise.rs:82-85)
                    const float lam = __ldg(lrow + ix);
                    q = poisson_f64(rng, (double)__fmul_rn(__fmul_rn(lam, c.delta), c.delta));
                } else {
                    double p = standard_f64(rng);
                    while (p > e) { p = __dmul_rn(p, standard_f64(rng)); ++q; }
                }
            }
            const uint32_t inclB = warp_incl(q);
            uint32_t* ws = wsB[batch_par];
            batch_par ^= 1;
            if (lane == 31) ws[warp] = inclB;
            __syncthreads();
            uint32_t total;
            const uint32_t off = block_off(ws, total) + inclB - q;
            if ((uint64_t)run + total > cap) { // uniform
                if (tid == 0) atomicExch(overflow, 1u);
                return;
            }
            if (t < M) {
                cnt[cl] = q;
                const uint64_t idx = base + run + off;
                for (uint32_t g = 0; g < q; ++g) {
                    const float cx = __fadd_rn(sx, uniform_f32(rng, c.uscale_cell));
                    const float cy = __fadd_rn(sy, uniform_f32(rng, c.uscale_cell));
                    Gg[idx + g] = make_float2(cx, cy);
                    if (LOGN) { // RadiusProfile::sample + clamp (src/model.rs:137-148, src/pixelwise.rs:89-95)
                        const float radius = radius_sample_clamped(c.rad, rng);
                        R2g[idx + g] = radius > 0.0f ? __fmul_rn(radius, radius) : -1.0f;
                    }
                }
            }
            run += total;
        }
        __syncthreads(); // cnt complete
        // ---- c: prefix entries of the tile's cells ----
        {
            const uint4 c0 = ((const uint4*)cnt)[2 * tid], c1 = ((const uint4*)cnt)[2 * tid + 1];
            const uint32_t s8 = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;
            const uint32_t inclC = warp_incl(s8);
            if (lane == 31) wsC[warp] = inclC;
            __syncthreads();
            uint32_t total;
            // grains before this tile = run - (tile total)
            uint32_t p0 = block_off(wsC, total) + inclC - s8;
            p0 += run - total;
            const int k0 = tile0 + 8 * tid;
            if (k0 < (int)geo.ppitch) {
                uint4 o0, o1;
                o0.x = p0; o0.y = o0.x + c0.x; o0.z = o0.y + c0.y; o0.w = o0.z + c0.z;
                o1.x = o0.w + c0.w; o1.y = o1.x + c1.x; o1.z = o1.y + c1.y; o1.w = o1.z + c1.z;
                uint4* dst = (uint4*)(prow + k0);
                dst[0] = o0;
                dst[1] = o1;
            }
        }
        __syncthreads(); // cnt / list / wsA are rewritten by the next tile
    }
}

} // namespace fg
