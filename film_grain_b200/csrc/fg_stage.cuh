// fg_stage.cuh -- the CELL TABLE: every Boolean-model cell of a render band generated exactly once.
//
// Reference semantics (src/pixelwise.rs:60-104): a cell (i, j) is seeded from (seed, i, j) only, so its
// grains are the same for every pixel, sample and colour-plane-independent part of the chain that visits
// it.  The reference regenerates the cell on every visit; the strip kernel (fg_tile.cuh) used to
// regenerate it once per strip window.  Here one kernel walks the band's cell rows and writes, per
// colour plane, a CSR table in HBM:
//
//   Pg[plane][row][col]   u32, col = 0 .. cols: number of grains of the row's cells [0, col)
//   Gg[rowbase[plane,row] + Pg[...]]   float2 grain centres in cell order (+ R2g: squared radius, log-normal;
//                                      + Cg: the grain's cell column mod 2^16, for k_pixelwise_skew's row merge)
//
// so that a strip window is two contiguous slices per cell row and the strip kernel only loads and
// evaluates.  Row storage is sized from the row's expected grain count (the sum of its cells' Poisson
// means) plus eight standard deviations; a row that would not fit raises `overflow` and the caller
// falls back to in-kernel generation (never observed; P < 1e-15 per row).
#pragma once
#include "fg_kernels.cuh"

namespace fg {


// Expected grains of one cell row, per (plane, input row): sum over the band's cell columns of
// lambda' = lambda * delta^2 (src/pixelwise.rs:76-81) at the clamped input pixel of the column.
__global__ void __launch_bounds__(256) k_row_expect(const float* __restrict__ lambda, size_t in_stride, int i0, int cols,
                                                     int iy_first, double* __restrict__ S, RenderConsts c) {
    const int iy = iy_first + blockIdx.x, pl = blockIdx.y; // only the input rows the band's cell rows map to
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
__global__ void __launch_bounds__(1024) k_row_bases(const double* __restrict__ S, int j0, int rows, int n_planes, double slack_sigma,
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
        const double cap = fmax(s + slack_sigma * sqrt(s), 0.0) + 64.0;
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

// 128-bit content hash of the lambda planes (two position-keyed splitmix sums, so the launch geometry does not matter):
// the identity of a cached cell table.  acc[0], acc[1] are zeroed by the caller.
__global__ void __launch_bounds__(256) k_hash_planes(const float* __restrict__ v, size_t n, unsigned long long* __restrict__ acc) {
    uint64_t a = 0, b = 0;
    for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (size_t)gridDim.x * 256) {
        const uint64_t x = (uint64_t)__float_as_uint(__ldg(v + t));
        a += splitmix64(x ^ ((uint64_t)t * 0xD6E8FEB86659FD93ULL));
        b += splitmix64((x << 32 | x) + rotl64((uint64_t)t + 0x2545F4914F6CDD1DULL, 29));
    }
    for (int d = 16; d; d >>= 1) {
        a += __shfl_down_sync(0xFFFFFFFFu, a, d);
        b += __shfl_down_sync(0xFFFFFFFFu, b, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, (unsigned long long)a);
        atomicAdd(acc + 1, (unsigned long long)b);
    }
}

struct StageGeo {
    int i0, j0, cols, rows;   // cell rectangle of the band (same as the first-draw bitmap)
    uint32_t pitchw;          // bitmap words per row
    uint32_t ppitch;          // Pg entries per row (>= cols + 1, multiple of 8)
};

// The general sampler as a real call: inlined three times (once per plane) it set the register count of the whole kernel,
// i.e. the occupancy of the hot Knuth path, although only saturated content ever runs it.
__device__ __noinline__ uint32_t poisson_f64_counted(XoshiroCounted& r, double lambda) { return poisson_f64(r, lambda); }

// One WARP per cell row, no CTA-wide barriers.  NPL = 3: the warp generates the row for all three colour planes
// together; NPL = 1: one warp per (plane, row), any plane count.  A cell is seeded from (seed, i, j) only
// (src/rng.rs:26-29), and its Knuth chain p_k = U1 ... Uk is the same sequence for every plane -- a plane only decides
// where it stops (p_k <= exp(-lambda'_plane), src/pixelwise.rs:76-85 via rand_distr's Poisson) -- so the hash, the
// eight PCG seed words and the chain are computed ONCE per cell that is non-empty in any plane, and planes that stop at
// the same count share their grain positions (the generator state after the Poisson draw is the same).
//
// The row is walked in tiles of FG_GW_TILE cells: the tile's non-empty cells (union of the planes' first-draw bitmaps)
// are compacted into a list and processed 32 at a time, one lane per cell at full occupancy: seeding, two Knuth draws
// (76% of the non-empty (cell, plane) pairs at lambda' = 1/pi stop there with one grain, whose position is drawn once
// and stored for every such plane), Knuth continuation for the rest.  (Cell, plane) pairs with two or more grains are
// the divergent tail: the generator state after the second draw, the number of outputs to skip and the destination are
// parked in a per-warp queue and their positions are drawn 32 pairs at a time when the queue fills.  A lane one of whose
// planes needs the general sampler (lambda' >= 12, e < 0) parks all its planes from the seed state instead, with the
// sampler's own draw count as the skip.  A final scan over the tile's counts writes the prefix entries of ALL its cells.
#define FG_GW_WARPS 4
// Tile size and resident CTAs per SM, measured at C2 (table pass, ms; joint instance): 1024 / unconstrained (130
// registers, 3 CTAs) 21.1; 1024 / 4 CTAs 19.2; 512 / 4 19.1; 512 / 5 (96 registers) 18.8; 768 / 5 18.4; 768 / 6 (80
// registers, spills) 19.3; 512 / 7 19.9; 256 / 7 21.0.  Per-plane instance: unconstrained 20.8, 5 CTAs 19.7, 6 19.3,
// 7 19.9, 8 21.1.  The kernel is latency-bound (55% issue-active at 12 warps per SM), so one more resident CTA pays
// for a few spills -- until the spills reach the chain loop.
#ifndef FG_GW_TILE
#define FG_GW_TILE 768 // cells per tile: a multiple of 256 (phase c: FG_GW_TILE / 32 cells per lane, 8 per store pair)
#endif
#ifndef FG_GW_MINB
#define FG_GW_MINB 5
#endif
#ifndef FG_GW_MINB1
#define FG_GW_MINB1 6 // the per-plane instance needs fewer registers
#endif
#define FG_GW_QCAP 64
template <int NPL>
struct GenWarpSmem {
    uint16_t list[FG_GW_TILE];
    __align__(16) uint16_t cnt[NPL][FG_GW_TILE];
    uint64_t q_s0[FG_GW_QCAP], q_s1[FG_GW_QCAP], q_s2[FG_GW_QCAP], q_s3[FG_GW_QCAP], q_dst[FG_GW_QCAP];
    uint32_t q_q[FG_GW_QCAP], q_skip[FG_GW_QCAP];
    float q_sx[FG_GW_QCAP];
    uint16_t q_col[FG_GW_QCAP];
};

template <bool LOGN, int NPL>
__global__ void __launch_bounds__(FG_GW_WARPS * 32, NPL == 1 ? FG_GW_MINB1 : FG_GW_MINB) k_gen_rows(const uint32_t* __restrict__ bm_planes, size_t bm_plane_words,
                                                                const double* __restrict__ e_planes, const float* __restrict__ lambda,
                                                                size_t in_stride, uint32_t* __restrict__ Pg,
                                                                const uint64_t* __restrict__ rowbase, const uint32_t* __restrict__ rowcap,
                                                                float2* __restrict__ Gg, float* __restrict__ R2g, uint16_t* __restrict__ Cg,
                                                                uint32_t* __restrict__ overflow, StageGeo geo, int n_planes, RenderConsts c) {
    __shared__ GenWarpSmem<NPL> sm_all[FG_GW_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GenWarpSmem<NPL>& sm = sm_all[warp];
    const size_t widx0 = (size_t)blockIdx.x * FG_GW_WARPS + warp;
    if (widx0 >= (size_t)geo.rows * (NPL == 1 ? n_planes : 1)) return; // whole warp
    if (warp_aborted(c)) return;
    const int plane0 = NPL == 1 ? (int)(widx0 / geo.rows) : 0, row = (int)(widx0 - (size_t)plane0 * geo.rows);
    const int j = geo.j0 + row;
    const float sy = __fmul_rn(__int2float_rn(j), c.delta);
    const int iy = min(max(floor_i32(sy), 0), c.in_h - 1);
    // per-plane addresses are the first plane's plus a plane stride (registers: 130 -> fewer, one more resident CTA);
    // the row's base and capacity are read from the (L1-resident) row arrays where they are needed
    const double* const erow0 = e_planes + in_stride * plane0 + (size_t)iy * c.in_w;
    const float* const lrow0 = lambda + in_stride * plane0 + (size_t)iy * c.in_w;
    const uint32_t* const bmrow0 = bm_planes + bm_plane_words * plane0 + (size_t)row * geo.pitchw;
    const size_t ridx0 = (size_t)plane0 * geo.rows + row;
    uint32_t* const prow0 = Pg + ridx0 * geo.ppitch;
    const size_t pstride = (size_t)geo.rows * geo.ppitch; // Pg entries between the same row of two planes
    uint32_t run[NPL]; // grains of the row placed so far
#pragma unroll
    for (int pl = 0; pl < NPL; ++pl) run[pl] = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t qn = 0;  // parked (cell, plane) pairs

    auto warp_incl = [&](uint32_t v) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v += u;
        }
        return v;
    };
    auto warp_incl64 = [&](uint64_t v) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v += u;
        }
        return v;
    };
    // Draw the parked pairs, 32 at a time from the top of the queue and two grains per round: a pair with more than two
    // grains left goes back into the queue with its advanced state, so every round runs at full lane occupancy whatever
    // the counts are.  `all` = the row is finished: drain the partial rounds too.
    auto flush = [&](bool all) {
        __syncwarp();
        while (qn >= 32u || (all && qn)) {
            const uint32_t b = qn >= 32u ? qn - 32u : 0u, t = b + lane;
            bool again = false;
            Xoshiro rng;
            float sx = 0.0f;
            uint64_t dst = 0;
            uint32_t q = 0;
            uint16_t col = 0;
            if (t < qn) {
                rng.s0 = sm.q_s0[t]; rng.s1 = sm.q_s1[t]; rng.s2 = sm.q_s2[t]; rng.s3 = sm.q_s3[t];
                for (uint32_t k = sm.q_skip[t]; k; --k) advance(rng);
                sx = sm.q_sx[t];
                dst = sm.q_dst[t];
                q = sm.q_q[t];
                col = sm.q_col[t];
                const uint32_t now = min(q, 2u);
                for (uint32_t g = 0; g < now; ++g) {
                    const float cx = __fadd_rn(sx, uniform_f32(rng, c.uscale_cell));
                    const float cy = __fadd_rn(sy, uniform_f32(rng, c.uscale_cell));
                    Gg[dst + g] = make_float2(cx, cy);
                    Cg[dst + g] = col; // the grain's cell column in the table (mod 2^16): the evaluation kernels merge rows by column
                    if (LOGN) { // RadiusProfile::sample + clamp (src/model.rs:137-148, src/pixelwise.rs:89-95)
                        const float radius = radius_sample_clamped(c.rad, rng);
                        R2g[dst + g] = radius > 0.0f ? __fmul_rn(radius, radius) : -1.0f;
                    }
                }
                again = q > 2u;
            }
            __syncwarp(); // every lane has read its entry before the survivors overwrite the slots
            const uint32_t am = __ballot_sync(0xFFFFFFFFu, again);
            if (again) {
                const uint32_t slot = b + __popc(am & lt_mask);
                sm.q_s0[slot] = rng.s0; sm.q_s1[slot] = rng.s1; sm.q_s2[slot] = rng.s2; sm.q_s3[slot] = rng.s3;
                sm.q_dst[slot] = dst + 2u;
                sm.q_q[slot] = q - 2u;
                sm.q_skip[slot] = 0u;
                sm.q_sx[slot] = sx;
                sm.q_col[slot] = col;
            }
            qn = b + __popc(am);
            __syncwarp();
        }
    };

    for (int tile0 = 0; tile0 < (int)geo.ppitch; tile0 += FG_GW_TILE) {
        // ---- a: clear the counts, compact the tile's non-empty cells into `list` ----
        {
            uint4* cz = (uint4*)&sm.cnt[0][0]; // NPL x 2 KiB
#pragma unroll
            for (int k = 0; k < NPL * FG_GW_TILE * 2 / 16 / 32; ++k) cz[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        }
        const uint32_t widx = (uint32_t)(tile0 >> 5) + (uint32_t)lane;
        uint32_t bits = 0u;
        if (lane < FG_GW_TILE / 32 && widx < geo.pitchw) {
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) bits |= __ldg(bmrow0 + bm_plane_words * pl + widx);
        }
        const uint32_t nb = __popc(bits);
        const uint32_t inclA = warp_incl(nb);
        const uint32_t M = __shfl_sync(0xFFFFFFFFu, inclA, 31);
        {
            uint32_t pos = inclA - nb;
            while (bits) {
                const int b0 = __ffs(bits) - 1;
                bits &= bits - 1;
                sm.list[pos++] = (uint16_t)(lane * 32 + b0);
            }
        }
        __syncwarp();
        // ---- b: 32 non-empty cells at a time ----
        for (uint32_t b0 = 0; b0 < M; b0 += 32) {
            const uint32_t t = b0 + lane;
            uint32_t q[NPL], skip[NPL];
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) { q[pl] = 0; skip[pl] = 0; }
            uint32_t cl = 0;
            Xoshiro sb; // state the lane's planes are drawn from: after two outputs, or the seed state (general lanes)
            sb.s0 = sb.s1 = sb.s2 = sb.s3 = 0;
            bool genlane = false;
            float sx = 0.0f;
            if (t < M) {
                cl = sm.list[t];
                const int i = geo.i0 + tile0 + (int)cl;
                sx = __fmul_rn(__int2float_rn(i), c.delta);
                const int ix = min(max(floor_i32(sx), 0), c.in_w - 1);
                double e[NPL];
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    e[pl] = __ldg(erow0 + in_stride * pl + ix);
                    genlane |= e[pl] < 0.0;
                }
                Xoshiro rng;
                seed_small_rng(rng, mix3_row(mix3_col(c.seed_cell, i), j), c.seeding);
                if (genlane) { // some plane has lambda' >= 12 or a non-finite mean: the general sampler (src/pixelwise.rs:82-85)
                    sb = rng;
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) {
                        XoshiroCounted rc;
                        rc.s0 = rng.s0; rc.s1 = rng.s1; rc.s2 = rng.s2; rc.s3 = rng.s3; rc.n = 0;
                        if (e[pl] < 0.0) {
                            const float lam = __ldg(lrow0 + in_stride * pl + ix);
                            q[pl] = poisson_f64_counted(rc, (double)__fmul_rn(__fmul_rn(lam, c.delta), c.delta));
                        } else {
                            double p = standard_f64(rc);
                            while (p > e[pl]) { p = __dmul_rn(p, standard_f64(rc)); ++q[pl]; }
                        }
                        skip[pl] = rc.n;
                    }
                } else {
                    double p = standard_f64(rng);
                    uint32_t act = 0;
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) act |= (p > e[pl]) ? (1u << pl) : 0u;
                    p = __dmul_rn(p, standard_f64(rng));
                    sb = rng;
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) {
                        q[pl] = (act >> pl) & 1u;
                        if (!(p > e[pl])) act &= ~(1u << pl);
                    }
                    while (act) {
                        p = __dmul_rn(p, standard_f64(rng));
#pragma unroll
                        for (int pl = 0; pl < NPL; ++pl) {
                            q[pl] += (act >> pl) & 1u;
                            if (!(p > e[pl])) act &= ~(1u << pl);
                        }
                    }
#pragma unroll
                    for (int pl = 0; pl < NPL; ++pl) skip[pl] = q[pl] - 1u; // q >= 2 pairs only
                }
            }
            bool big = false;
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) big |= q[pl] > 65535u;
            if (__any_sync(0xFFFFFFFFu, big)) { // uniform: a cell with more grains than the 16-bit counts hold
                if (lane == 0) atomicExch(overflow, 1u);
                return;
            }
            uint32_t excl[NPL], total[NPL];
            if (NPL == 3) { // three 21-bit fields: 32 x 65535 < 2^21
                const uint64_t v = (uint64_t)q[0] | ((uint64_t)q[NPL > 1 ? 1 : 0] << 21) | ((uint64_t)q[NPL > 2 ? 2 : 0] << 42);
                const uint64_t incl = warp_incl64(v);
                const uint64_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31), ex = incl - v;
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    excl[pl] = (uint32_t)(ex >> (21 * pl)) & 0x1FFFFFu;
                    total[pl] = (uint32_t)(tot >> (21 * pl)) & 0x1FFFFFu;
                }
            } else {
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl) {
                    const uint32_t incl = warp_incl(q[pl]);
                    total[pl] = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    excl[pl] = incl - q[pl];
                }
            }
            bool over = false;
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) over |= (uint64_t)run[pl] + total[pl] > __ldg(rowcap + ridx0 + (size_t)geo.rows * pl);
            if (over) { // uniform
                if (lane == 0) atomicExch(overflow, 1u);
                return;
            }
            uint64_t dst[NPL];
            const uint16_t col16 = (uint16_t)((uint32_t)tile0 + cl);
            uint32_t one = 0; // planes whose cell holds exactly one grain: its position is drawn once, here
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
                dst[pl] = __ldg(rowbase + ridx0 + (size_t)geo.rows * pl) + run[pl] + excl[pl];
                if (q[pl]) sm.cnt[pl][cl] = (uint16_t)q[pl];
                if (q[pl] == 1u && !genlane) one |= 1u << pl;
                run[pl] += total[pl];
            }
            if (one) {
                Xoshiro rng = sb;
                const float cx = __fadd_rn(sx, uniform_f32(rng, c.uscale_cell));
                const float cy = __fadd_rn(sy, uniform_f32(rng, c.uscale_cell));
                float r2 = 0.0f;
                if (LOGN) {
                    const float radius = radius_sample_clamped(c.rad, rng);
                    r2 = radius > 0.0f ? __fmul_rn(radius, radius) : -1.0f;
                }
#pragma unroll
                for (int pl = 0; pl < NPL; ++pl)
                    if ((one >> pl) & 1u) {
                        Gg[dst[pl]] = make_float2(cx, cy);
                        Cg[dst[pl]] = col16;
                        if (LOGN) R2g[dst[pl]] = r2;
                    }
            }
#pragma unroll
            for (int pl = 0; pl < NPL; ++pl) {
                const bool parked = q[pl] != 0u && !((one >> pl) & 1u);
                const uint32_t pm = __ballot_sync(0xFFFFFFFFu, parked);
                if (parked) {
                    const uint32_t slot = qn + __popc(pm & lt_mask);
                    sm.q_s0[slot] = sb.s0; sm.q_s1[slot] = sb.s1; sm.q_s2[slot] = sb.s2; sm.q_s3[slot] = sb.s3;
                    sm.q_dst[slot] = dst[pl];
                    sm.q_q[slot] = q[pl];
                    sm.q_skip[slot] = skip[pl];
                    sm.q_sx[slot] = sx;
                    sm.q_col[slot] = col16;
                }
                qn += __popc(pm);
                if (qn >= 32u) flush(false); // leaves fewer than 32 parked: the next plane adds at most 32
            }
        }
        __syncwarp(); // cnt complete
        // ---- c: prefix entries of the tile's cells: lane l owns FG_GW_TILE / 32 consecutive cells ----
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) {
            constexpr int CPLN = FG_GW_TILE / 32, NV = CPLN / 8; // cells per lane, 16-byte count vectors per lane
            const uint4* cs = (const uint4*)(sm.cnt[pl] + CPLN * lane);
            uint4 v[NV];
            uint32_t s32 = 0;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                v[k] = cs[k];
                s32 += (v[k].x & 0xFFFFu) + (v[k].x >> 16) + (v[k].y & 0xFFFFu) + (v[k].y >> 16) + (v[k].z & 0xFFFFu) + (v[k].z >> 16) +
                       (v[k].w & 0xFFFFu) + (v[k].w >> 16);
            }
            const uint32_t inclC = warp_incl(s32);
            const uint32_t tile_total = __shfl_sync(0xFFFFFFFFu, inclC, 31);
            uint32_t pacc = run[pl] - tile_total + inclC - s32; // grains of the row before this lane's first cell
            const int k0 = tile0 + CPLN * lane;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
                uint4 o0, o1;
                o0.x = pacc; pacc += w[0] & 0xFFFFu;
                o0.y = pacc; pacc += w[0] >> 16;
                o0.z = pacc; pacc += w[1] & 0xFFFFu;
                o0.w = pacc; pacc += w[1] >> 16;
                o1.x = pacc; pacc += w[2] & 0xFFFFu;
                o1.y = pacc; pacc += w[2] >> 16;
                o1.z = pacc; pacc += w[3] & 0xFFFFu;
                o1.w = pacc; pacc += w[3] >> 16;
                if (k0 + 8 * k < (int)geo.ppitch) { // ppitch is a multiple of 8
                    uint4* dstp = (uint4*)(prow0 + pstride * pl + k0 + 8 * k);
                    dstp[0] = o0;
                    dstp[1] = o1;
                }
            }
        }
        __syncwarp(); // cnt / list are rewritten by the next tile
    }
    flush(true);
}

} // namespace fg
