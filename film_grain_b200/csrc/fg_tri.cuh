// fg_tri.cuh -- k_pixelwise_tri: the pixel-wise evaluation kernel for the default geometry rm == delta, constant radius
// (cell_delta = 1 / ceil(1 / r): every sample point visits exactly three cell rows -- BASELINE configs 1, 2 and 4).
//
// Reference semantics (src/pixelwise.rs:47-106): a sample point (xg, yg) visits the cells [i0, i1] x [j0, j1] within
// rm of it and is covered if any grain of those cells is within its radius.  The cell table (fg_stage.cuh) holds every
// cell's grains, generated once.  One CTA evaluates a strip segment of 32 output columns x SEG rows of one plane:
//
//   * lane = output column, warp = SPW samples.  The cell-column range of (column, sample) is computed once per
//     segment (two IEEE divisions) and lives in registers; the cell-row range of (row, sample) is computed once per warp
//     (lanes compute a batch of rows x samples in parallel) and is fetched as one 16-byte broadcast load.
//   * SKEW.  Sample k of reference row yr evaluates output row yr + sk_k with sk_k = rint(oy_k * zoom), so that the
//     sample points of one step all lie within one output row of each other: a step needs D = m*cpr + 4 cell rows
//     instead of the whole spread of the sample offsets.
//   * MERGED TRIPLES.  For every cell row j of the window the rows j, j+1, j+2 are merged column by column into ONE
//     list M with a 16-bit prefix Q[j][i] (absolute ring index of the first entry of column i).  A sample reads
//     Q[j0][i0] and Q[j0][i1 + 1] and walks that range with U straight-line predicated grain tests (immediate-offset
//     loads, no row selection, a lane never loads a grain that is not its own, lanes that are already covered stop
//     loading); what is left (n > U and not yet covered, a few lanes per thousand) is finished after the warp's
//     samples of the step by an early-exit loop.
//   * LOADER.  Table slices travel by bulk asynchronous copies (cp.async.bulk, mbarrier completion): per source row
//     its prefix slice, grain slice and cell-column slice, issued at the start of a step and landing while it is evaluated;
//     the merge (three copies of every grain, Q = sum of three prefix rows) runs from shared memory after the step's
//     evaluation, for the next step.  One CTA
//     barrier per step.
//   * Coverage bits: one ballot word per (row, sample) in a ring; a finished row is transposed (32 x 32 bit
//     transposes through shuffles), popcounted and written as count * (1/N).
//
// Samples that do not visit exactly three cell rows, or start outside the step's window (f32 rounding of
// (y -/+ rm) / delta), are evaluated from the HBM table like k_pixelwise_table_tiles does; a segment whose merged window
// does not fit shared memory (dense content) goes to the fallback list.  Results are bit-identical to
// k_pixelwise_strip / k_pixelwise_direct / the oracle: the visited cell set, the f32 operations of the distance test
// and the count are the reference's; only the order of the (commutative) "any grain covers" changes.
#pragma once
#include "fg_tile.cuh"

namespace fg {

#define FG_TRI_WARPS 32
#define FG_TRI_THREADS (FG_TRI_WARPS * 32)
#ifndef FG_TRI_GSZ
#define FG_TRI_GSZ 4 // samples whose grain tests are interleaved (instruction-level parallelism vs registers)
#endif
#ifndef FG_TRI_U
#define FG_TRI_U 6 // straight-line predicated grain tests per sample before the deferred remainder
#endif

struct TriCfg {
    int m;            // output rows per sample and step
    int D;            // triple rows a step's samples can start in
    int AMAX;         // bound on the triple rows one step adds
    int NQ;           // ring slots of merged triple rows (>= D + AMAX)
    int NG;           // bound on the source rows of one load group (AMAX + 2 <= 32)
    int CWB, PS;      // bound on the window's cell columns; prefix row stride (entries, multiple of 8)
    int MCAP;         // merged grain ring (entries, < 65536; FG_TRI_U + 1 entries of padding behind it)
    int GS;           // staging bytes (grains + cell columns) per group buffer
    int RH;           // rows of the coverage-bit ring (power of two)
    int NB;           // steps per item batch (NB * m * SPW <= 32)
    int SEG, n_strips, n_segs;
    int bm_i0, bm_j0, bm_cols, bm_rows; // the cell table's rectangle
    uint32_t ppitch;
    float r2c, inv_delta;
    uint32_t off_zq, off_zp, off_praw, off_gs, off_sinfo, off_ext, off_minfo, off_state, off_items, off_hb, off_Q, off_M, total;
};

// ---- shared-memory / async-copy primitives (32-bit shared addresses) ----
__device__ __forceinline__ uint32_t tri_lds_u16(uint32_t addr) { // not volatile: free to be scheduled, inputs change every step
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t tri_lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 tri_lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void tri_sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void tri_sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tri_sts_f32x2(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tri_mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void tri_mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void tri_mbar_arrive_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tri_mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TRI_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra TRI_DONE;\n\t"
        "bra TRI_WAIT;\n\t"
        "TRI_DONE:\n\t"
        "}" ::"r"(a), "r"(parity) : "memory");
}
// global -> shared bulk copy (1-D TMA): 16-byte aligned source, destination and size; completes on the mbarrier
__device__ __forceinline__ void tri_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(mbar) : "memory");
}

// first triple row of the window of the step whose first reference row is yr (any deterministic, monotone rule will
// do: an item outside its step's window takes the slow path)
__device__ __forceinline__ int tri_J(int yr, float inv_zoom, float rm, float inv_delta) {
    return floor_i32(__fmul_rn(__fsub_rn(__fmul_rn((float)yr, inv_zoom), rm), inv_delta)) - 1;
}

// One sample of the step: its item {Q row address, coverage-word address, yg, -}, the two prefix loads, then FG_TRI_U
// straight-line grain tests of the merged list [s16, e16).  A test is: if (u < n && not yet covered) load M[s16 + u];
// not_covered &= |p - g|^2 > r^2 -- the un-fused f32 sequence of src/pixelwise.rs:96-98 on the packed pipe (sub / mul
// as f32x2, then one add), six instructions.  Only the load is predicated: a lane never loads a grain that is not its
// own and stops loading once it is covered; a lane that does not load re-tests the grain it tested last (same verdict),
// and a lane with no candidate at all tests the far-away grain at `zinf` in the first slot.  `pg` = "not covered"
// (unordered compare: a NaN distance does not cover, like `<=` in the reference).  Grains are stored (cy, cx): the
// sample point pairs the item's yg with the lane's xg in the register pair the item load already filled.
#define FG_TRI_STR2(x) #x
#define FG_TRI_STR(x) FG_TRI_STR2(x)
#define FG_TRI_SLOT(UU)                               \
    "setp.gt.and.u32 q, %2, " #UU ", pg;\n\t"         \
    "@q ld.shared.b64 g, [%3+" #UU "*8];\n\t"         \
    "sub.rn.f32x2 d, %4, g;\n\t"                      \
    "mul.rn.f32x2 s2, d, d;\n\t"                      \
    "mov.b64 {lo, hi}, s2;\n\t"                       \
    "add.rn.f32 dd, lo, hi;\n\t"                      \
    "setp.gtu.and.f32 pg, dd, %5, pg;\n\t"
#if FG_TRI_U == 3
#define FG_TRI_SLOTS FG_TRI_SLOT(1) FG_TRI_SLOT(2)
#elif FG_TRI_U == 4
#define FG_TRI_SLOTS FG_TRI_SLOT(1) FG_TRI_SLOT(2) FG_TRI_SLOT(3)
#elif FG_TRI_U == 5
#define FG_TRI_SLOTS FG_TRI_SLOT(1) FG_TRI_SLOT(2) FG_TRI_SLOT(3) FG_TRI_SLOT(4)
#elif FG_TRI_U == 6
#define FG_TRI_SLOTS FG_TRI_SLOT(1) FG_TRI_SLOT(2) FG_TRI_SLOT(3) FG_TRI_SLOT(4) FG_TRI_SLOT(5)
#elif FG_TRI_U == 7
#define FG_TRI_SLOTS FG_TRI_SLOT(1) FG_TRI_SLOT(2) FG_TRI_SLOT(3) FG_TRI_SLOT(4) FG_TRI_SLOT(5) FG_TRI_SLOT(6)
#else
#error "FG_TRI_U must be 3 .. 7"
#endif
__device__ __forceinline__ void tri_sample(uint32_t item_addr, float xg, uint32_t ab, uint32_t Ms, float r2, uint32_t zinf,
                                           uint32_t sbit, uint32_t& hits, uint32_t& rem) {
    uint32_t qa, hb, w2, w3;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(qa), "=r"(hb), "=r"(w2), "=r"(w3) : "r"(item_addr));
    const uint32_t s16 = tri_lds_u16(qa + (ab & 0xFFFFu)), e16 = tri_lds_u16(qa + (ab >> 16));
    const uint32_t n = e16 - s16;
    const uint32_t ga = Ms + s16 * 8u;
    const uint64_t pp = pack_f32x2(__uint_as_float(w2), xg);
    asm("{\n\t"
        ".reg .pred q, pg;\n\t"
        ".reg .b64 d, s2, g;\n\t"
        ".reg .f32 lo, hi, dd;\n\t"
        ".reg .b32 a0;\n\t"
        "setp.gt.u32 q, %2, 0;\n\t"
        "selp.b32 a0, %3, %7, q;\n\t"
        "ld.shared.b64 g, [a0];\n\t"
        "sub.rn.f32x2 d, %4, g;\n\t"
        "mul.rn.f32x2 s2, d, d;\n\t"
        "mov.b64 {lo, hi}, s2;\n\t"
        "add.rn.f32 dd, lo, hi;\n\t"
        "setp.gtu.f32 pg, dd, %5;\n\t"
        FG_TRI_SLOTS
        "setp.gt.and.u32 q, %2, " FG_TRI_STR(FG_TRI_U) ", pg;\n\t" // more candidates and still not covered: deferred
        "@q or.b32 %1, %1, %6;\n\t"
        "@!pg or.b32 %0, %0, %6;\n\t"
        "}"
        : "+r"(hits), "+r"(rem)
        : "r"(n), "r"(ga), "l"(pp), "f"(r2), "r"(sbit), "r"(zinf));
    (void)hb;
    (void)w3;
}

template <int SPW>
__global__ void __launch_bounds__(FG_TRI_THREADS, 1)
k_pixelwise_tri(const float* __restrict__ lambda, size_t in_stride, const float2* __restrict__ offsets_input,
                float* __restrict__ out, size_t out_stride, TileRef* __restrict__ fb_list, uint32_t* __restrict__ fb_count,
                uint32_t fb_cap, TriCfg cfg, RenderConsts c, CellTable tab) {
    constexpr int NW = FG_TRI_WARPS;
    constexpr uint32_t HBROW = 32u * SPW * 4u; // bytes of one coverage row: one word per sample
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int unit = blockIdx.x;
    const int strip = unit % cfg.n_strips;
    const int seg = (unit / cfg.n_strips) % cfg.n_segs;
    const int plane = unit / (cfg.n_strips * cfg.n_segs);
    const int X0 = strip * 32;
    const int Y0 = c.row_begin + seg * cfg.SEG;
    if (X0 >= c.out_w || Y0 >= c.row_end) return;
    const int X1 = min(X0 + 32, c.out_w) - 1;    // inclusive
    const int Y1 = min(Y0 + cfg.SEG, c.row_end); // exclusive
    float* outp = out + out_stride * plane;
    const float rm = c.rad.rm, delta = c.delta, r2 = cfg.r2c;
    const bool radius_ok = (c.rad.mean_linear > rm ? rm : c.rad.mean_linear) > 0.0f; // radius <= 0: grains never cover
    const int m = cfg.m, D = cfg.D, NQ = cfg.NQ, NG = cfg.NG, PS = cfg.PS;

    // ---- the strip's cell columns (monotone in x and in the offset) ----
    const float bx0 = __fmul_rn(__fadd_rn((float)X0, 0.5f), c.inv_zoom);
    const float bx1 = __fmul_rn(__fadd_rn((float)X1, 0.5f), c.inv_zoom);
    const int i_lo = cell_lo(__fsub_rn(bx0, c.off_max_x), rm, delta);
    const int i_hi = cell_hi(__fsub_rn(bx1, c.off_min_x), rm, delta);
    // prefix rows are fetched from the table column rounded down to a multiple of four (16-byte alignment): the window
    // simply starts there (up to three extra cells on the left)
    const long long tcol0 = (long long)i_lo - cfg.bm_i0;
    const long long tcolA = tcol0 & ~3LL;
    const long long CWl = (long long)i_hi - cfg.bm_i0 - tcolA + 1; // window columns
    const bool geo_bad = i_lo > i_hi || tcol0 < 0 || (long long)i_hi >= (long long)cfg.bm_i0 + cfg.bm_cols || CWl > cfg.CWB || !radius_ok;
    if (geo_bad) { // uniform
        if (!radius_ok) { // nothing ever covers: zeros (src/pixelwise.rs:93-95)
            for (int p = tid; p < (Y1 - Y0) * 32; p += FG_TRI_THREADS)
                if (X0 + (p & 31) <= X1) outp[(size_t)(Y0 + (p >> 5)) * c.out_w + X0 + (p & 31)] = 0.0f;
        } else if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
        return;
    }
    const int CW = (int)CWl;
    const int i_loA = cfg.bm_i0 + (int)tcolA;
    const uint32_t tcolA_u = (uint32_t)tcolA;

    // ---- shared-memory map ----
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("" : "+r"(sbase)); // opaque: one register instead of re-deriving the window address at every use
    const uint32_t PSB = (uint32_t)PS * 2u, PRB = (uint32_t)PS * 4u; // bytes of a Q row / a raw prefix row
    const uint32_t ZQ = sbase + cfg.off_zq, ZP = sbase + cfg.off_zp, PRAW = sbase + cfg.off_praw, GSs = sbase + cfg.off_gs;
    const uint32_t SINFO = sbase + cfg.off_sinfo, MINFO = sbase + cfg.off_minfo, STATE = sbase + cfg.off_state;
    const uint32_t HB = sbase + cfg.off_hb, Qs = sbase + cfg.off_Q;
    uint32_t Ms = sbase + cfg.off_M;
    uint32_t wps = sbase + cfg.off_items + (uint32_t)warp * 512u; // 32 items of 16 bytes per warp
    asm volatile("" : "+r"(wps), "+r"(Ms)); // opaque: kept in registers instead of being re-derived in every loop
    uint4* ext = (uint4*)(smem + cfg.off_ext);          // [2][NG] {Pg[first], Pg[last], rowbase lo, hi} of a group's source rows
    uint32_t* state = (uint32_t*)(smem + cfg.off_state); // [0] head of the merged ring, [1] failure flag; mbarriers at +16, +24
    const uint32_t MBAR = STATE + 16u;
    const uint32_t ZINF = STATE + 32u; // a grain at (inf, inf): what a sample without candidates tests
    const uint32_t HBDUMP = HB + (uint32_t)cfg.RH * HBROW; // coverage words of samples that do not exist go here

    for (uint32_t p = (uint32_t)tid * 4u; p < PSB; p += FG_TRI_THREADS * 4u) tri_sts_u32(ZQ + p, 0u);
    for (uint32_t p = (uint32_t)tid * 4u; p < PRB; p += FG_TRI_THREADS * 4u) tri_sts_u32(ZP + p, 0u);
    for (uint32_t p = (uint32_t)tid * 4u; p < (uint32_t)(cfg.RH + 1) * HBROW; p += FG_TRI_THREADS * 4u) tri_sts_u32(HB + p, 0u);
    if (tid == 0) {
        state[0] = 0u;
        state[1] = 0u;
        state[2] = 0u;
        state[8] = 0x7f800000u;
        state[9] = 0x7f800000u;
        tri_mbar_init(MBAR, 32u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- per-thread (column, sample) data: abscissa and packed byte offsets of Q[.][i0], Q[.][i1 + 1] ----
    const int x = X0 + lane;
    const bool xvalid = x <= X1;
    const float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);
    float xg_r[SPW];
    uint32_t ab_r[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
        const uint32_t k = (uint32_t)(s * NW + warp);
        float xg = 0.0f;
        uint32_t ab = 0u; // a == b: an empty range (inactive lane / sample)
        if (k < c.n && xvalid) {
            xg = __fsub_rn(bx, __ldg(offsets_input + k).x);
            ab = col_range_packed(xg, rm, delta, i_loA);
        }
        xg_r[s] = xg;
        ab_r[s] = ab;
    }

    // ---- steps.  Reference row yr of step t: Yref0 + t m + r; sample k evaluates output row yr + sk_k. ----
    const int skmax = __float2int_rn(__fmul_rn(c.off_max_y, c.zoom)), skmin = __float2int_rn(__fmul_rn(c.off_min_y, c.zoom));
    const int Yref0 = Y0 - skmax;
    const int T = (Y1 - 1 - skmin - Yref0) / m + 1;
    const float inv_zoom = c.inv_zoom, inv_delta = cfg.inv_delta;
    auto Jof = [&](int t) { return tri_J(Yref0 + t * m, inv_zoom, rm, inv_delta); };
    // load groups: group u brings the triple rows [J(u-1) + D, J(u) + D); the first one that matters is uS
    const int Jw0 = Jof(0);
    int uS = 0;
    while (Jof(uS - 1) + D > Jw0 && uS > -4096) --uS;
    const int Jbase = Jof(uS - 1) + D; // first triple row ever merged: ring slot of row d = (d - Jbase) mod NQ

    // ---- item lanes: lane -> (step of the batch, row of the step, sample of this warp) ----
    const int ipl = m * SPW;
    const int ib = lane / ipl, irem = lane - ib * ipl, ir = irem / SPW, is = irem - ir * SPW;
    const bool ilane = ib < cfg.NB;
    const uint32_t ik = (uint32_t)(is * NW + warp);
    const bool ikvalid = ilane && ik < c.n;
    const float ioy = ikvalid ? __ldg(offsets_input + ik).y : 0.0f;
    const int isk = __float2int_rn(__fmul_rn(ioy, c.zoom));
    const uint32_t ihb = (uint32_t)(warp * SPW + is) * 4u;
    uint32_t slowmask = 0u; // items of the current batch that need the general evaluation

    // Q-merge work split: item `it` = (triple row d of the group, 16-byte vector v of its prefix row), it = tid, tid + 1024, ...
    const uint32_t nv8 = (uint32_t)PS >> 3;
    const uint32_t qd0 = (uint32_t)tid / nv8, qv0 = (uint32_t)tid - qd0 * nv8;
    const uint32_t qdstep = FG_TRI_THREADS / nv8, qvstep = FG_TRI_THREADS - qdstep * nv8;

    int ydone = Y0; // rows below are written
    // rolling window starts J(tau - 1) .. J(tau + 2), the ring slot of the first triple row of group tau + 1, and the step's
    // position in its item batch (all uniform; kept in registers instead of being recomputed with divisions)
    int J0c = Jof(uS - 2), J1c = Jof(uS - 1), J2c = Jof(uS);
    uint32_t slotG = 0u; // (J(tau) + D - Jbase) mod NQ, valid once tau + 1 >= uS
    int bstep = 0;
    __syncthreads();

    // rows finished by the steps before `tau`: transpose the coverage words, count, write
    auto finalise = [&](int tau) {
        const int ynew = min(Yref0 + tau * m + skmin, Y1);
        for (int yy = ydone + ((warp - tau) & (NW - 1)); yy < ynew; yy += NW) {
            const uint32_t hrow = HB + ((uint32_t)(yy - Y0) & (uint32_t)(cfg.RH - 1)) * HBROW + (uint32_t)lane * 4u;
            uint32_t cnt = 0;
#pragma unroll
            for (int cw = 0; cw < SPW; ++cw) {
                uint32_t w = tri_lds_u32(hrow + (uint32_t)cw * 128u); // word of sample index cw * 32 + lane: bit = column
#pragma unroll
                for (int j = 16; j >= 1; j >>= 1) { // 32 x 32 bit transpose: afterwards lane = column, bit = sample
                    const uint32_t mk = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
                    const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, w, j);
                    w = (lane & j) ? ((w & ~mk) | ((y >> j) & mk)) : ((w & mk) | ((y << j) & ~mk));
                }
                cnt += __popc(w);
            }
            if (xvalid) outp[(size_t)yy * c.out_w + x] = __fmul_rn((float)cnt, c.inv_samples);
        }
        ydone = max(ydone, ynew);
    };

    for (int tau = uS - 2; tau < T; ++tau) {
        int myfail = 0; // set by the lanes of the placement warp; every thread learns it at the step's barrier
        // ---- (1) extents of the source rows of group tau + 2 (cp.async, awaited before this step's barrier) ----
        {
            const int u = tau + 2;
            if (u >= uS && u < T && warp == ((tau + 16) & (NW - 1))) {
                const int gA = J1c + D, NGu = J2c + D + 2 - gA;
                if (lane < NGu) {
                    const long long trow = (long long)gA + lane - cfg.bm_j0;
                    uint4* e = ext + (u & 1) * NG + lane;
                    if (trow >= 0 && trow < cfg.bm_rows) {
                        const size_t row = (size_t)plane * cfg.bm_rows + (size_t)trow;
                        const uint32_t* pg = tab.Pg + row * cfg.ppitch + tcolA_u;
                        cp_async4(&e->x, pg);
                        cp_async4(&e->y, pg + CW);
                        cp_async8(&e->z, tab.rowbase + row);
                    } else {
                        *e = make_uint4(0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu); // a row outside the table: empty
                    }
                }
            }
        }
        // ---- (2) group tau + 1: placement in the staging buffer and in the merged ring, bulk copies (they land while the
        //      step is evaluated; the staging buffer was released by the previous step's barrier) ----
        {
            const int u = tau + 1;
            if (u >= uS && u < T && warp == (tau & (NW - 1))) {
                const uint32_t b = (uint32_t)u & 1u; // extents are double buffered, the staging buffer is not
                const int gA = J0c + D, A_u = J1c - J0c, NGu = A_u + 2;
                uint32_t f = 0u, n = 0u, shift = 0u, cnt8 = 0u;
                uint64_t gal = 0ull;
                bool intab = false;
                if (lane < NGu) {
                    const uint4 e = ext[b * NG + lane];
                    intab = !(e.z == 0xFFFFFFFFu && e.w == 0xFFFFFFFFu);
                    if (intab) {
                        f = e.x;
                        n = e.y - e.x;
                        const uint64_t gsrc = (((uint64_t)e.w << 32) | e.z) + f;
                        shift = (uint32_t)gsrc & 7u; // grains and 16-bit columns are copied from a multiple of eight entries
                        gal = gsrc - shift;
                        cnt8 = n ? ((shift + n + 7u) & ~7u) : 0u;
                    }
                }
                const uint32_t sz = cnt8 * 10u;
                uint32_t si = sz, ti;
                const uint32_t n1 = __shfl_down_sync(0xFFFFFFFFu, n, 1), n2 = __shfl_down_sync(0xFFFFFFFFu, n, 2);
                const uint32_t Tr = lane < A_u ? n + n1 + n2 : 0u;
                ti = Tr;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t a1 = __shfl_up_sync(0xFFFFFFFFu, si, d), a2 = __shfl_up_sync(0xFFFFFFFFu, ti, d);
                    if (lane >= d) { si += a1; ti += a2; }
                }
                const uint32_t stotal = __shfl_sync(0xFFFFFFFFu, si, 31), ttotal = __shfl_sync(0xFFFFFFFFu, ti, 31);
                bool fail = stotal > (uint32_t)cfg.GS || __any_sync(0xFFFFFFFFu, n > 60000u);
                // merged ring: rows are placed linearly; a group that would cross the end continues at 0 from its first row
                // that does not fit.  Live rows: from the first row of the step evaluated while this group is merged.
                const uint32_t MC = (uint32_t)cfg.MCAP;
                const uint32_t head = state[0];
                const int jt = tau >= 0 ? J0c : Jw0; // first row of the step evaluated while this group is merged
                uint32_t tail = head;
                if (jt < gA && jt >= Jbase) { // its ring slot: D rows (or less) behind the group's first row
                    uint32_t st = slotG + (uint32_t)NQ - (uint32_t)(gA - jt);
                    if (st >= (uint32_t)NQ) st -= (uint32_t)NQ;
                    tail = tri_lds_u32(MINFO + st * 4u);
                }
                const uint32_t cross = __ballot_sync(0xFFFFFFFFu, lane < A_u && head + ti > MC);
                uint32_t mstart, newhead;
                if (cross) {
                    const int rs = __ffs(cross) - 1;
                    const uint32_t bstar = __shfl_sync(0xFFFFFFFFu, ti - Tr, rs);
                    mstart = lane >= rs ? ti - Tr - bstar : head + ti - Tr;
                    newhead = ttotal - bstar;
                    if (head >= tail) fail = fail || !(newhead < tail);
                    else fail = true;
                } else {
                    mstart = head + ti - Tr;
                    newhead = head + ttotal;
                    if (head < tail) fail = fail || !(newhead < tail);
                }
                // per source row: where its slices land, and the constant parts of the destinations of its grains
                const uint32_t f1 = __shfl_down_sync(0xFFFFFFFFu, f, 1), f2 = __shfl_down_sync(0xFFFFFFFFu, f, 2);
                const uint32_t fm1 = __shfl_up_sync(0xFFFFFFFFu, f, 1), fm2 = __shfl_up_sync(0xFFFFFFFFu, f, 2);
                const uint32_t msm1 = __shfl_up_sync(0xFFFFFFFFu, mstart, 1), msm2 = __shfl_up_sync(0xFFFFFFFFu, mstart, 2);
                const uint32_t so = GSs + (si - sz);
                const uint32_t praw = intab ? PRAW + (uint32_t)lane * PRB : ZP;
                if (lane < NGu) {
                    const uint32_t sa = SINFO + (uint32_t)lane * 32u;
                    tri_sts_v4(sa, praw, n, so + shift * 8u, so + cnt8 * 8u + shift * 2u);
                    tri_sts_v4(sa + 16u, mstart - f1 - f2, msm1 - fm1 - f1, msm2 - fm2 - fm1, mstart - f - f1 - f2);
                    if (lane < A_u) {
                        uint32_t sm = slotG + (uint32_t)lane;
                        if (sm >= (uint32_t)NQ) sm -= (uint32_t)NQ;
                        tri_sts_u32(MINFO + sm * 4u, mstart);
                    }
                }
#ifdef FG_TRI_DEBUG
                if (fail && lane == 0)
                    printf("tri fail unit %d X0 %d Y0 %d tau %d T %d uS %d A %d head %u tail %u ttotal %u stotal %u cross %x newhead %u jt %d gA %d Jbase %d\n", unit, X0, Y0, tau, T, uS,
                           A_u, head, tail, ttotal, stotal, cross, newhead, jt, gA, Jbase);
#endif
                if (!fail) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (lane < NGu && intab) {
                        const size_t row = (size_t)plane * cfg.bm_rows + (size_t)((long long)gA + lane - cfg.bm_j0);
                        tri_mbar_arrive_tx(MBAR, PRB + sz);
                        tri_bulk_g2s(praw, tab.Pg + row * cfg.ppitch + tcolA_u, PRB, MBAR);
                        if (cnt8) {
                            tri_bulk_g2s(so, tab.Gg + gal, cnt8 * 8u, MBAR);
                            tri_bulk_g2s(so + cnt8 * 8u, tab.Cg + gal, cnt8 * 2u, MBAR);
                        }
                    } else {
                        tri_mbar_arrive(MBAR);
                    }
                    if (lane == 0) state[0] = newhead;
                } else { // release the waiting warps; they skip the merge, and the step's barrier ends the segment
                    myfail = 1;
                    if (lane == 0) state[2] = 1u;
                    tri_mbar_arrive(MBAR);
                }
            }
        }
        // ---- (4) evaluate step tau ----
        if (tau >= 0) {
            if (bstep == 0) { // items of the steps tau .. tau + NB - 1: one (row, sample) per lane
                const int tb = tau + ib;
                const int y = Yref0 + tb * m + ir + isk;
                const bool ok = ikvalid && tb < T && y >= Y0 && y < Y1;
                float yg = 0.0f;
                uint32_t qa = ZQ, hb = HBDUMP + ihb, w3 = 0u;
                bool slow = false;
                if (ok) {
                    yg = __fsub_rn(__fmul_rn(__fadd_rn((float)y, 0.5f), inv_zoom), ioy);
                    const int j0 = cell_lo(yg, rm, delta), j1 = cell_hi(yg, rm, delta);
                    const int Jb = Jof(tb);
                    hb = HB + ((uint32_t)(y - Y0) & (uint32_t)(cfg.RH - 1)) * HBROW + ihb;
                    const bool fast = j1 - j0 == 2 && j0 >= Jb && j0 < Jb + D && j0 >= cfg.bm_j0 && (long long)j0 + 2 < (long long)cfg.bm_j0 + cfg.bm_rows;
                    if (fast) qa = Qs + ((uint32_t)(j0 - Jbase) % (uint32_t)NQ) * PSB;
                    else if (j0 <= j1) { slow = true; w3 = (uint32_t)y; }
                }
                __syncwarp();
                tri_sts_v4(wps + (uint32_t)lane * 16u, qa, hb, __float_as_uint(yg), w3);
                slowmask = __ballot_sync(0xFFFFFFFFu, slow);
                __syncwarp();
            }
            const uint32_t ibase = wps + (uint32_t)(bstep * ipl) * 16u;
            for (int r = 0; r < m; ++r) {
                constexpr int GSZ = SPW < FG_TRI_GSZ ? SPW : FG_TRI_GSZ; // samples whose tests are interleaved; their coverage words are stored after the group
#pragma unroll
                for (int s0 = 0; s0 < SPW; s0 += GSZ) {
                    const uint32_t ia = ibase + (uint32_t)(r * SPW + s0) * 16u;
                    uint32_t rem = 0u, hits = 0u; // per lane: bit s = sample s0 + s has more candidates / is covered
#pragma unroll
                    for (int s = 0; s < GSZ; ++s) tri_sample(ia + (uint32_t)s * 16u, xg_r[s0 + s], ab_r[s0 + s], Ms, r2, ZINF, 1u << s, hits, rem);
                    // the lanes that are not covered after FG_TRI_U candidates and have more: early-exit walk of the rest
                    if (__any_sync(0xFFFFFFFFu, rem != 0u)) {
#pragma unroll
                        for (int s = 0; s < GSZ; ++s) {
                            if ((rem >> s) & 1u) {
                                const uint4 it = tri_lds_v4(ia + (uint32_t)s * 16u);
                                const uint32_t s16 = tri_lds_u16(it.x + (ab_r[s0 + s] & 0xFFFFu)), e16 = tri_lds_u16(it.x + (ab_r[s0 + s] >> 16));
                                const uint32_t ga = Ms + s16 * 8u;
                                const uint64_t pp = pack_f32x2(__uint_as_float(it.z), xg_r[s0 + s]);
                                for (uint32_t uu = (uint32_t)FG_TRI_U; uu < e16 - s16; ++uu) {
                                    if (dist2_packed(pp, lds_f32x2(ga + uu * 8u)) <= r2) { hits |= 1u << s; break; }
                                }
                            }
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (int s = 0; s < GSZ; ++s) tri_sts_u32(tri_lds_u32(ia + (uint32_t)s * 16u + 4u), __ballot_sync(0xFFFFFFFFu, (hits >> s) & 1u));
                }
            }
            // ---- the rare items that do not visit exactly three cell rows or start outside the window: walk their cells in
            //      the HBM table, like k_pixelwise_table_tiles.  Inline on purpose (a call would cost the hot loop its
            //      registers).  Cells outside the table cannot occur for planned geometry; if they do, the segment is
            //      handed to the fallback kernel, which overwrites it. ----
            uint32_t sb = (slowmask >> (bstep * ipl)) & (ipl >= 32 ? 0xFFFFFFFFu : ((1u << ipl) - 1u));
            while (sb) { // uniform
                const int L = __ffs(sb) - 1;
                sb &= sb - 1;
                const uint4 it = tri_lds_v4(ibase + (uint32_t)L * 16u);
                const float ygs = __uint_as_float(it.z);
                const int sj0 = cell_lo(ygs, rm, delta), sj1 = cell_hi(ygs, rm, delta);
                if (sj0 < cfg.bm_j0 || (long long)sj1 >= (long long)cfg.bm_j0 + cfg.bm_rows) {
#ifdef FG_TRI_DEBUG
                    if (lane == 0) printf("tri slow-outside unit %d tau %d sj0 %d sj1 %d bm_j0 %d rows %d\n", unit, tau, sj0, sj1, cfg.bm_j0, cfg.bm_rows);
#endif
                    if (lane == 0 && atomicExch(&state[1], 2u) == 0u) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
                    continue;
                }
                const uint32_t k = (uint32_t)((L % SPW) * NW + warp);
                uint32_t hit = 0u;
                if (xvalid) {
                    const float xgs = __fsub_rn(bx, __ldg(offsets_input + k).x);
                    const int si0 = cell_lo(xgs, rm, delta), si1 = cell_hi(xgs, rm, delta);
                    if (si0 <= si1) {
                        for (int j = sj0; j <= sj1 && !hit; ++j) {
                            const size_t row = (size_t)plane * cfg.bm_rows + (size_t)(j - cfg.bm_j0);
                            const uint32_t* pr = tab.Pg + row * cfg.ppitch + (uint32_t)(si0 - cfg.bm_i0);
                            const uint32_t gs = __ldg(pr), ge = __ldg(pr + (si1 - si0 + 1));
                            const float2* gp = tab.Gg + (size_t)__ldg(tab.rowbase + row);
                            for (uint32_t g = gs; g < ge; ++g) {
                                const float2 gr = __ldg(gp + g);
                                const float dx = __fsub_rn(xgs, gr.x), dy = __fsub_rn(ygs, gr.y);
                                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= r2) { hit = 1u; break; }
                            }
                        }
                    }
                }
                tri_sts_u32(it.y, __ballot_sync(0xFFFFFFFFu, hit != 0u));
            }
        }
        // ---- (3) merge group tau + 1 (its slices arrived while the step was evaluated) ----
        {
            const int u = tau + 1;
            if (u >= uS && u < T) {
                const int A_u = J1c - J0c, NGu = A_u + 2;
                tri_mbar_wait(MBAR, (uint32_t)(u - uS) & 1u);
                const uint32_t sib = SINFO;
                if (((volatile uint32_t*)state)[2] == 0u) {
                // Q[d][e] = start of triple d + P[d][e] + P[d+1][e] + P[d+2][e] (each relative to its window start): eight
                // 16-bit entries per item
                {
                    const uint32_t slotA = slotG;
                    uint32_t d = qd0, v = qv0;
                    while (d < (uint32_t)A_u) {
                        const uint32_t sa = sib + d * 32u;
                        const uint32_t p0 = tri_lds_u32(sa), p1 = tri_lds_u32(sa + 32u), p2 = tri_lds_u32(sa + 64u);
                        const uint32_t qb = tri_lds_u32(sa + 28u);
                        const uint4 a0 = tri_lds_v4(p0 + v * 32u), a1 = tri_lds_v4(p0 + v * 32u + 16u);
                        const uint4 b0 = tri_lds_v4(p1 + v * 32u), b1 = tri_lds_v4(p1 + v * 32u + 16u);
                        const uint4 c0 = tri_lds_v4(p2 + v * 32u), c1 = tri_lds_v4(p2 + v * 32u + 16u);
                        const uint32_t q0 = a0.x + b0.x + c0.x + qb, q1 = a0.y + b0.y + c0.y + qb, q2 = a0.z + b0.z + c0.z + qb, q3 = a0.w + b0.w + c0.w + qb;
                        const uint32_t q4 = a1.x + b1.x + c1.x + qb, q5 = a1.y + b1.y + c1.y + qb, q6 = a1.z + b1.z + c1.z + qb, q7 = a1.w + b1.w + c1.w + qb;
                        uint32_t sl = slotA + d;
                        if (sl >= (uint32_t)NQ) sl -= (uint32_t)NQ;
                        tri_sts_v4(Qs + sl * PSB + v * 16u, (q0 & 0xFFFFu) | (q1 << 16), (q2 & 0xFFFFu) | (q3 << 16), (q4 & 0xFFFFu) | (q5 << 16),
                                   (q6 & 0xFFFFu) | (q7 << 16));
                        d += qdstep;
                        v += qvstep;
                        if (v >= nv8) { v -= nv8; ++d; }
                    }
                }
                // grains: warps over the source rows, lanes over a row's grains.  A grain with window-local index g in window
                // column e goes to
                //   triple r     at  start(r)   - f(r+1) - f(r+2) + P[r+1][e]   + P[r+2][e]   + g   (its row is the triple's first),
                //   triple r - 1 at  start(r-1) - f(r-1) - f(r+1) + P[r-1][e+1] + P[r+1][e]   + g   (second),
                //   triple r - 2 at  start(r-2) - f(r-2) - f(r-1) + P[r-2][e+1] + P[r-1][e+1] + g   (third).
                { // warp -> (source row r, part of its grains); NGu <= 32, so every row has at least one warp
                    int r = warp, part = 0, nparts = 1;
                    while (r >= NGu) { r -= NGu; ++part; } // warp = part * NGu + r without a division
                    for (int w2 = r + NGu; w2 < NW; w2 += NGu) ++nparts;
                    nparts = max(nparts, part + 1);
                    const uint32_t sa = sib + (uint32_t)r * 32u;
                    const uint4 ri = tri_lds_v4(sa), bi = tri_lds_v4(sa + 16u);
                    const bool t0 = r < A_u, t1 = r >= 1 && r - 1 < A_u, t2 = r >= 2;
                    const uint32_t pp1 = (t0 || t1) ? tri_lds_u32(sa + 32u) : ZP, pp2 = t0 ? tri_lds_u32(sa + 64u) : ZP;
                    const uint32_t pm1 = (t1 || t2) ? tri_lds_u32(sa - 32u) : ZP, pm2 = t2 ? tri_lds_u32(sa - 64u) : ZP;
                    for (uint32_t g = (uint32_t)(part * 32 + lane); g < ri.y; g += 32u * (uint32_t)nparts) {
                        const uint32_t e4 = ((tri_lds_u16(ri.w + g * 2u) - tcolA_u) & 0xFFFFu) * 4u; // byte offset of the grain's window column
                        const float2 gr = lds_f32x2(ri.z + g * 8u);
                        const uint32_t x1 = tri_lds_u32(pp1 + e4), x2 = tri_lds_u32(pp2 + e4);
                        const uint32_t y1 = tri_lds_u32(pm1 + e4 + 4u), y2 = tri_lds_u32(pm2 + e4 + 4u);
                        if (t0) tri_sts_f32x2(Ms + (bi.x + g + x1 + x2) * 8u, gr.y, gr.x);
                        if (t1) tri_sts_f32x2(Ms + (bi.y + g + y1 + x1) * 8u, gr.y, gr.x);
                        if (t2) tri_sts_f32x2(Ms + (bi.z + g + y2 + y1) * 8u, gr.y, gr.x);
                    }
                }
                }
            }
        }
        // ---- (5) rows every sample has passed ----
        if (tau >= 1) finalise(tau);
        cp_async_wait_all();
        if (__syncthreads_or(myfail)) { // the merged window or the staging buffer does not hold this content: the fallback kernel renders the segment
            if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
            return;
        }
        if (tau + 1 >= uS) { // the group placed in the next step starts where this one ended
            slotG += (uint32_t)(J1c - J0c);
            if (slotG >= (uint32_t)NQ) slotG -= (uint32_t)NQ;
        }
        if (tau >= 0 && ++bstep == cfg.NB) bstep = 0;
        J0c = J1c; J1c = J2c; J2c = Jof(tau + 3);
    }
    finalise(T);
}

} // namespace fg
