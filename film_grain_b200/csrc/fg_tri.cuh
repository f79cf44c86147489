// fg_tri.cuh -- k_pixelwise_tri: the pixel-wise evaluation kernel for the default geometry rm == delta, constant radius
// (cell_delta = 1 / ceil(1 / r): every sample point visits exactly three cell rows -- BASELINE configs 1, 2 and 4).
//
// Reference semantics (src/pixelwise.rs:47-106): a sample point (xg, yg) visits the cells [i0, i1] x [j0, j1] within
// rm of it and is covered if any grain of those cells is within its radius.  The cell table (fg_stage.cuh) holds every
// cell's grains, generated once.  One CTA evaluates a strip segment of 32 output columns x SEG rows of one plane:
//
//   * WARP ROLES.  16 evaluation warps (lane = output column, warp = SPW samples) and 4 loader warps that fetch, merge and
//     finalise; they meet only through mbarriers (full / empty per step parity), never at a CTA-wide barrier, so an
//     evaluation warp that is done with a step starts the next one as soon as its rows are merged.
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
//     its prefix slice, grain slice and cell-column slice, issued one step ahead; the loader warps merge them (three
//     copies of every grain, Q = sum of three prefix rows) while the evaluation warps work on the previous step.
//   * Coverage bits: one ballot word per (row, sample) in a ring; a finished row is transposed (32 x 32 bit
//     transposes through shuffles), popcounted and written as count * (1/N).
//
// Samples that visit FOUR cell rows (f32 rounding of (y -/+ rm) / delta: 0.4% of the items from y = 2048 on) are served
// from two merged triples; those with any other row count, or that start outside the step's window, are evaluated from
// the HBM table like k_pixelwise_table_tiles does; a segment whose merged window does not fit shared memory (dense
// content) goes to the fallback list.  Instances: 16 samples per evaluation warp (128 < N <= 256) and, where a step
// needs at most five cell rows per output row, 4 and 8 (32 < N <= 128); the planner (fg_pixel_host.cuh: tri_plan) chooses.  Results are bit-identical to
// k_pixelwise_strip / k_pixelwise_direct / the oracle: the visited cell set, the f32 operations of the distance test
// and the count are the reference's; only the order of the (commutative) "any grain covers" changes.
#pragma once
#include "fg_tile.cuh"

namespace fg {

#ifndef FG_TRI_EWARPS
#define FG_TRI_EWARPS 16 // evaluation warps (24 measured slower at C2: 72 registers per thread spill in the test loop)
#endif
#ifndef FG_TRI_DWARPS
#define FG_TRI_DWARPS 4  // loader / merge / output warps: loader warp 0 (placement + its share of the merge) sets the step period; 3: 29.6 ms, 4: 28.2 ms, 5 / 6: 32.2 / 32.6 ms (registers) at C2
#endif
#define FG_TRI_WARPS FG_TRI_EWARPS // samples are dealt over the evaluation warps: k = s * FG_TRI_WARPS + warp
#define FG_TRI_THREADS ((FG_TRI_EWARPS + FG_TRI_DWARPS) * 32)
// samples per evaluation warp of the three instances: N <= 64, N <= 128, N <= 256
#define FG_TRI_SPW_A ((64 + FG_TRI_EWARPS - 1) / FG_TRI_EWARPS)
#define FG_TRI_SPW_B ((128 + FG_TRI_EWARPS - 1) / FG_TRI_EWARPS)
#define FG_TRI_SPW_C ((256 + FG_TRI_EWARPS - 1) / FG_TRI_EWARPS)
#ifndef FG_TRI_GSZ
#define FG_TRI_GSZ 4 // samples whose grain tests are interleaved (1, 2 or 4: instruction-level parallelism vs registers)
#endif
#ifndef FG_TRI_U
#define FG_TRI_U 6 // straight-line predicated grain tests per sample before the deferred remainder
#endif

#ifdef FG_TRI_TIMING // debug: cycles per phase of one CTA, printed at the end (tools/tri_probe.py)
#define FG_TT_DECL long long tt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tt_last = clock64()
#define FG_TT(i) do { const long long tt_now = clock64(); tt_acc[i] += tt_now - tt_last; tt_last = tt_now; } while (0)
#else
#define FG_TT_DECL
#define FG_TT(i)
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
// wait for the phase of the given parity; the hardware suspends the thread for up to `hint_ns` per attempt instead of
// spinning (a spinning loader warp takes issue slots from the evaluation warps of its scheduler)
__device__ __forceinline__ void tri_mbar_wait(uint32_t a, uint32_t parity, uint32_t hint_ns = 2000u) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TRI_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra TRI_DONE;\n\t"
        "bra TRI_WAIT;\n\t"
        "TRI_DONE:\n\t"
        "}" ::"r"(a), "r"(parity), "r"(hint_ns) : "memory");
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

// The grain tests of G samples of a step.  Per sample: its item {Q row address, coverage-word address, yg, -}, the two
// prefix loads, then FG_TRI_U straight-line tests of the merged list [s16, e16).  A test is: if (u < n && not yet
// covered) load M[s16 + u]; not_covered &= |p - g|^2 > r^2 -- the un-fused f32 sequence of src/pixelwise.rs:96-98 on
// the packed pipe (sub / mul as f32x2, then one add), six instructions.  Only the load is predicated: a lane never loads
// a grain that is not its own and stops loading once it is covered; a lane that does not load re-tests the grain it
// tested last (same verdict), and a lane with no candidate at all tests the far-away grain at `zinf` in the first slot.
// Grains are stored (cy, cx): the sample point pairs the item's yg with the lane's xg.  The G chains are strictly serial
// each, so the PTX (fg_tri_asm.inc, generated by tools/gen_tri_asm.py) interleaves them stage by stage.
#include "fg_tri_asm.inc"
#define FG_TRI_CAT2(a, b, c, d) a##b##c##d
#define FG_TRI_CAT(a, b, c, d) FG_TRI_CAT2(a, b, c, d)
#define FG_TRI_ASM(G) FG_TRI_CAT(FG_TRI_ASM_G, G, _U, FG_TRI_U)

struct TriSample { uint32_t n, ga; uint64_t pp; };
__device__ __forceinline__ TriSample tri_fetch(uint32_t item_addr, float xg, uint32_t ab, uint32_t Ms) {
    uint32_t qa, hb, w2, w3;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(qa), "=r"(hb), "=r"(w2), "=r"(w3) : "r"(item_addr));
    const uint32_t s16 = tri_lds_u16(qa + (ab & 0xFFFFu)), e16 = tri_lds_u16(qa + (ab >> 16));
    (void)hb;
    (void)w3;
    TriSample t;
    t.n = e16 - s16;
    t.ga = Ms + s16 * 8u;
    t.pp = pack_f32x2(__uint_as_float(w2), xg);
    return t;
}
template <int G>
__device__ __forceinline__ void tri_group(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                          uint32_t& hits, uint32_t& rem);
template <>
__device__ __forceinline__ void tri_group<4>(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                             uint32_t& hits, uint32_t& rem) {
    const TriSample t0 = tri_fetch(ia, xg[0], ab[0], Ms), t1 = tri_fetch(ia + 16u, xg[1], ab[1], Ms);
    const TriSample t2 = tri_fetch(ia + 32u, xg[2], ab[2], Ms), t3 = tri_fetch(ia + 48u, xg[3], ab[3], Ms);
    asm(FG_TRI_ASM(4)
        : "+r"(hits), "+r"(rem)
        : "r"(t0.n), "r"(t1.n), "r"(t2.n), "r"(t3.n), "r"(t0.ga), "r"(t1.ga), "r"(t2.ga), "r"(t3.ga), "l"(t0.pp), "l"(t1.pp), "l"(t2.pp),
          "l"(t3.pp), "f"(r2), "r"(zinf));
}
template <>
__device__ __forceinline__ void tri_group<3>(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                             uint32_t& hits, uint32_t& rem) {
    const TriSample t0 = tri_fetch(ia, xg[0], ab[0], Ms), t1 = tri_fetch(ia + 16u, xg[1], ab[1], Ms), t2 = tri_fetch(ia + 32u, xg[2], ab[2], Ms);
    asm(FG_TRI_ASM(3)
        : "+r"(hits), "+r"(rem)
        : "r"(t0.n), "r"(t1.n), "r"(t2.n), "r"(t0.ga), "r"(t1.ga), "r"(t2.ga), "l"(t0.pp), "l"(t1.pp), "l"(t2.pp), "f"(r2), "r"(zinf));
}
template <>
__device__ __forceinline__ void tri_group<2>(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                             uint32_t& hits, uint32_t& rem) {
    const TriSample t0 = tri_fetch(ia, xg[0], ab[0], Ms), t1 = tri_fetch(ia + 16u, xg[1], ab[1], Ms);
    asm(FG_TRI_ASM(2) : "+r"(hits), "+r"(rem) : "r"(t0.n), "r"(t1.n), "r"(t0.ga), "r"(t1.ga), "l"(t0.pp), "l"(t1.pp), "f"(r2), "r"(zinf));
}
template <>
__device__ __forceinline__ void tri_group<1>(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                             uint32_t& hits, uint32_t& rem) {
    const TriSample t0 = tri_fetch(ia, xg[0], ab[0], Ms);
    asm(FG_TRI_ASM(1) : "+r"(hits), "+r"(rem) : "r"(t0.n), "r"(t0.ga), "l"(t0.pp), "f"(r2), "r"(zinf));
}

// What a lane needs to re-derive the (column, sample) data of a sample chosen at run time (the remainder walk below):
// registers cannot be indexed dynamically, so the abscissa and the packed column range are recomputed from the offset.
struct TriLane {
    const float2* offsets;
    float bx, rm, delta;
    int i_loA, NE, warp;
};

#ifndef FG_TRI_REM_ROW
#define FG_TRI_REM_ROW 0 // 1: one remainder pass per (step, row), lanes walk different samples in parallel; 0: per group of samples
#endif

// G samples of one (step, row): the straight-line tests; hits / rem are accumulated at bit S0 + s
template <int SPW, int S0>
struct TriRow {
    static __device__ __forceinline__ void run(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                               uint32_t& hitsAll, uint32_t& remAll) {
        constexpr int G = SPW - S0 < FG_TRI_GSZ ? SPW - S0 : FG_TRI_GSZ;
        uint32_t rem = 0u, hits = 0u; // per lane: bit s = sample s has more candidates / is covered
        tri_group<G>(ia + (uint32_t)S0 * 16u, xg + S0, ab + S0, Ms, r2, zinf, hits, rem);
#if !FG_TRI_REM_ROW
        // the lanes that are not covered after FG_TRI_U candidates and have more: early-exit walk of the rest
        if (__any_sync(0xFFFFFFFFu, rem != 0u)) {
#pragma unroll
            for (int s = 0; s < G; ++s) {
                if ((rem >> s) & 1u) {
                    const uint4 it = tri_lds_v4(ia + (uint32_t)(S0 + s) * 16u);
                    const uint32_t s16 = tri_lds_u16(it.x + (ab[S0 + s] & 0xFFFFu)), e16 = tri_lds_u16(it.x + (ab[S0 + s] >> 16));
                    const uint32_t ga = Ms + s16 * 8u;
                    const uint64_t pp = pack_f32x2(__uint_as_float(it.z), xg[S0 + s]);
                    for (uint32_t uu = (uint32_t)FG_TRI_U; uu < e16 - s16; ++uu) {
                        if (dist2_packed(pp, lds_f32x2(ga + uu * 8u)) <= r2) { hits |= 1u << s; break; }
                    }
                }
            }
            __syncwarp();
        }
        rem = 0u;
#endif
        hitsAll |= hits << S0;
        remAll |= rem << S0;
        TriRow<SPW, S0 + G>::run(ia, xg, ab, Ms, r2, zinf, hitsAll, remAll);
    }
};
template <int SPW>
struct TriRow<SPW, SPW> {
    static __device__ __forceinline__ void run(uint32_t, const float*, const uint32_t*, uint32_t, float, uint32_t, uint32_t&, uint32_t&) {}
};

// The SPW samples of one (step, row): tests in groups of FG_TRI_GSZ, then ONE walk of what is left -- the lanes that are
// not covered after FG_TRI_U candidates of a sample and have more (a few per thousand (lane, sample) pairs, but some
// lane of a warp in most rows).  Every lane walks its own samples in the same loop, so two lanes with work in different
// samples advance together instead of one group after the other.  Then the coverage words.
template <int SPW>
__device__ __forceinline__ void tri_eval_row(uint32_t ia, const float* xg, const uint32_t* ab, uint32_t Ms, float r2, uint32_t zinf,
                                             const TriLane& L) {
    uint32_t hits = 0u, rem = 0u;
    TriRow<SPW, 0>::run(ia, xg, ab, Ms, r2, zinf, hits, rem);
    if (__any_sync(0xFFFFFFFFu, rem != 0u)) {
        while (rem) {
            const int s = __ffs(rem) - 1;
            rem &= rem - 1u;
            const float xgs = __fsub_rn(L.bx, __ldg(L.offsets + (s * L.NE + L.warp)).x); // as in the set-up of xg_r / ab_r
            const uint32_t abs_ = col_range_packed(xgs, L.rm, L.delta, L.i_loA);
            const uint4 it = tri_lds_v4(ia + (uint32_t)s * 16u);
            const uint32_t s16 = tri_lds_u16(it.x + (abs_ & 0xFFFFu)), e16 = tri_lds_u16(it.x + (abs_ >> 16));
            const uint32_t ga = Ms + s16 * 8u;
            const uint64_t pp = pack_f32x2(__uint_as_float(it.z), xgs);
            for (uint32_t uu = (uint32_t)FG_TRI_U; uu < e16 - s16; ++uu) {
                if (dist2_packed(pp, lds_f32x2(ga + uu * 8u)) <= r2) { hits |= 1u << s; break; }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < SPW; ++s) tri_sts_u32(tri_lds_u32(ia + (uint32_t)s * 16u + 4u), __ballot_sync(0xFFFFFFFFu, (hits >> s) & 1u));
}

template <int SPW>
__global__ void __launch_bounds__(FG_TRI_THREADS, 1)
k_pixelwise_tri(const float* __restrict__ lambda, size_t in_stride, const float2* __restrict__ offsets_input,
                float* __restrict__ out, size_t out_stride, TileRef* __restrict__ fb_list, uint32_t* __restrict__ fb_count,
                uint32_t fb_cap, TriCfg cfg, RenderConsts c, CellTable tab) {
    constexpr int NE = FG_TRI_EWARPS, ND = FG_TRI_DWARPS;
    constexpr uint32_t HBROW = ((uint32_t)NE * SPW * 4u + 127u) / 128u * 128u; // bytes of one coverage row: one word per sample, padded to whole 32-word blocks (the padding stays zero)
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (cta_aborted(c)) return; // a cancelled render: the remaining CTAs of the launch drain in microseconds
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
    const int m = cfg.m, D = cfg.D, NQ = cfg.NQ, PS = cfg.PS;

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
    const uint32_t STATE = sbase + cfg.off_state;
    volatile uint32_t* state = (volatile uint32_t*)(smem + cfg.off_state); // [0] head of the merged ring, [1] "segment listed", [2] failure,
    // [3] the group being merged was skipped, [4] / [5] the window of the even / odd step holds skipped rows,
    // [32 ..] row ranges of the last eight groups if skipped (mbarriers live at bytes 64 .. 104)
    const uint32_t ZINF = STATE + 32u;  // a grain at (inf, inf): what a sample without candidates tests
    const uint32_t MB_COPY = STATE + 64u, MB_FULL = STATE + 72u, MB_EMPTY = STATE + 88u; // FULL / EMPTY: one barrier per step parity
    const uint32_t HB = sbase + cfg.off_hb, Qs = sbase + cfg.off_Q;
    const uint32_t ZQ = sbase + cfg.off_zq;

    for (uint32_t p = (uint32_t)tid * 4u; p < PSB; p += FG_TRI_THREADS * 4u) tri_sts_u32(ZQ + p, 0u);
    for (uint32_t p = (uint32_t)tid * 4u; p < PRB; p += FG_TRI_THREADS * 4u) tri_sts_u32(sbase + cfg.off_zp + p, 0u);
    for (uint32_t p = (uint32_t)tid * 4u; p < (uint32_t)(cfg.RH + 1) * HBROW; p += FG_TRI_THREADS * 4u) tri_sts_u32(HB + p, 0u);
    if (tid == 0) {
        state[0] = 0u;
        state[1] = 0u;
        state[2] = 0u;
        state[3] = 0u;
        state[4] = 0u;
        state[5] = 0u;
        for (int gq = 0; gq < 8; ++gq) { state[32 + 2 * gq] = 1u; state[33 + 2 * gq] = 0u; }
        state[8] = 0x7f800000u;
        state[9] = 0x7f800000u;
        tri_mbar_init(MB_COPY, 32u);
        tri_mbar_init(MB_FULL, (uint32_t)ND);
        tri_mbar_init(MB_FULL + 8u, (uint32_t)ND);
        tri_mbar_init(MB_EMPTY, (uint32_t)NE);
        tri_mbar_init(MB_EMPTY + 8u, (uint32_t)NE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    __syncthreads();

    if (warp < NE) {
        // =============================== evaluation warps ===============================
        uint32_t Ms = sbase + cfg.off_M;
        uint32_t wps = sbase + cfg.off_items + (uint32_t)warp * 512u; // 32 items of 16 bytes per warp
        asm volatile("" : "+r"(wps), "+r"(Ms));
        const uint32_t HBDUMP = HB + (uint32_t)cfg.RH * HBROW; // coverage words of samples that do not exist go here
        // per-thread (column, sample) data: abscissa and packed byte offsets of Q[.][i0], Q[.][i1 + 1]
        const int x = X0 + lane;
        const bool xvalid = x <= X1;
        const float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);
        float xg_r[SPW];
        uint32_t ab_r[SPW];
#pragma unroll
        for (int s = 0; s < SPW; ++s) {
            const uint32_t k = (uint32_t)(s * NE + warp);
            float xg = 0.0f;
            uint32_t ab = 0u; // a == b: an empty range (inactive lane / sample)
            if (k < c.n && xvalid) {
                xg = __fsub_rn(bx, __ldg(offsets_input + k).x);
                ab = col_range_packed(xg, rm, delta, i_loA);
            }
            xg_r[s] = xg;
            ab_r[s] = ab;
        }
        const TriLane tl{offsets_input, bx, rm, delta, i_loA, NE, warp};
        const int ipl = m * SPW; // items (row, sample) of one step
        uint32_t slowmask = 0u;  // items of the current batch that need the general evaluation
        int bstep = 0;           // position of the step in its item batch
        FG_TT_DECL;
        for (int tau = 0; tau < T; ++tau) {
            tri_mbar_wait(MB_FULL + (uint32_t)(tau & 1) * 8u, (uint32_t)(tau >> 1) & 1u); // the step's rows are merged
            FG_TT(0);
            if (state[2]) return; // the window does not fit: the fallback kernel renders the segment
            if (bstep == 0) { // items of the steps tau .. tau + NB - 1: one (row, sample) per lane
                const int ib = lane / ipl, irem = lane - ib * ipl, ir = irem / SPW, is = irem - ir * SPW;
                const uint32_t ik = (uint32_t)(is * NE + warp);
                const int tb = tau + ib;
                float yg = 0.0f;
                uint32_t qa = ZQ, hb = HBDUMP + (uint32_t)(warp * SPW + is) * 4u, w3 = 0u;
                bool slow = false;
                if (ib < cfg.NB && ik < c.n && tb < T) {
                    const float ioy = __ldg(offsets_input + ik).y;
                    const int y = Yref0 + tb * m + ir + __float2int_rn(__fmul_rn(ioy, c.zoom));
                    if (y >= Y0 && y < Y1) {
                        yg = __fsub_rn(__fmul_rn(__fadd_rn((float)y, 0.5f), inv_zoom), ioy);
                        const int j0 = cell_lo(yg, rm, delta), j1 = cell_hi(yg, rm, delta);
                        const int Jb = Jof(tb);
                        hb = HB + ((uint32_t)(y - Y0) & (uint32_t)(cfg.RH - 1)) * HBROW + (uint32_t)(warp * SPW + is) * 4u;
                        const bool fast = j1 - j0 == 2 && j0 >= Jb && j0 < Jb + D && j0 >= cfg.bm_j0 && (long long)j0 + 2 < (long long)cfg.bm_j0 + cfg.bm_rows;
                        if (fast) { w3 = (uint32_t)(j0 - Jbase); qa = Qs + (w3 % (uint32_t)NQ) * PSB; } // w3: the row, should its group turn out skipped
                        else if (j0 <= j1) {
                            slow = true;
                            // FOUR cell rows: from y = 2048 / zoom on, the f32 rounding of (yg -/+ rm) / delta makes 0.4% of the
                            // (row, sample) items span j0 .. j0 + 3 -- and 99% of the output rows hold such an item.  Rows
                            // j0 .. j0 + 3 are exactly the union of the merged triples j0 and j0 + 1, both in this step's
                            // window: the item is served from shared memory (bit 31 of w3), not by a walk of the HBM table.
                            if (j1 - j0 == 3 && j0 >= Jb && j0 + 1 < Jb + D && j0 >= cfg.bm_j0 && (long long)j0 + 3 < (long long)cfg.bm_j0 + cfg.bm_rows)
                                w3 = (uint32_t)(j0 - Jbase) | 0x80000000u;
                        }
                    }
                }
                __syncwarp();
                tri_sts_v4(wps + (uint32_t)lane * 16u, qa, hb, __float_as_uint(yg), w3);
                slowmask = __ballot_sync(0xFFFFFFFFu, slow);
                __syncwarp();
            }
            const uint32_t ibase = wps + (uint32_t)(bstep * ipl) * 16u;
            if (state[4 + (tau & 1)]) { // rows of this step's window were not merged (too dense for the ring): their items go the slow way
                bool conv = false;
                if (lane < ipl) {
                    const uint4 it = tri_lds_v4(ibase + (uint32_t)lane * 16u);
                    if (it.x != ZQ) {
#pragma unroll
                        for (int gq = 0; gq < 8; ++gq) conv = conv || (it.w >= state[32 + 2 * gq] && it.w <= state[33 + 2 * gq]); // rows of the last skipped groups
                        if (conv) tri_sts_u32(ibase + (uint32_t)lane * 16u, ZQ);
                    }
                }
                slowmask |= __ballot_sync(0xFFFFFFFFu, conv) << (bstep * ipl);
                __syncwarp();
            }
            for (int r = 0; r < m; ++r) tri_eval_row<SPW>(ibase + (uint32_t)(r * SPW) * 16u, xg_r, ab_r, Ms, r2, ZINF, tl);
            // ---- the rare items that do not visit exactly three cell rows or start outside the window: walk their cells in
            //      the HBM table, like k_pixelwise_table_tiles.  Cells outside the table cannot occur for planned geometry;
            //      if they do, the segment is handed to the fallback kernel, which overwrites it. ----
            uint32_t sb = (slowmask >> (bstep * ipl)) & (ipl >= 32 ? 0xFFFFFFFFu : ((1u << ipl) - 1u));
            while (sb) { // uniform
                const int L = __ffs(sb) - 1;
                sb &= sb - 1;
                const uint4 it = tri_lds_v4(ibase + (uint32_t)L * 16u);
                const float ygs = __uint_as_float(it.z);
                if ((it.w & 0x80000000u) && !state[4 + (tau & 1)]) { // four rows = two merged triples of the window (all rows merged)
                    const int ssel = L % SPW;
                    uint32_t ab = 0u;
                    float xg = 0.0f;
#pragma unroll
                    for (int s = 0; s < SPW; ++s)
                        if (s == ssel) { ab = ab_r[s]; xg = xg_r[s]; }
                    const uint64_t pp = pack_f32x2(ygs, xg);
                    const uint32_t wrow = it.w & 0x7FFFFFFFu;
                    bool hit4 = false;
#pragma unroll
                    for (uint32_t t = 0; t < 2u; ++t) {
                        const uint32_t qa4 = Qs + ((wrow + t) % (uint32_t)NQ) * PSB;
                        const uint32_t s16 = tri_lds_u16(qa4 + (ab & 0xFFFFu)), e16 = tri_lds_u16(qa4 + (ab >> 16));
                        for (uint32_t g = s16; g < e16 && !hit4; ++g) hit4 = dist2_packed(pp, lds_f32x2(Ms + g * 8u)) <= r2;
                    }
                    tri_sts_u32(it.y, __ballot_sync(0xFFFFFFFFu, hit4));
                    continue;
                }
                const int sj0 = cell_lo(ygs, rm, delta), sj1 = cell_hi(ygs, rm, delta);
                if (sj0 < cfg.bm_j0 || (long long)sj1 >= (long long)cfg.bm_j0 + cfg.bm_rows) {
                    if (lane == 0 && atomicExch((uint32_t*)&state[1], 1u) == 0u) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
                    continue;
                }
                const uint32_t k = (uint32_t)((L % SPW) * NE + warp);
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
            __syncwarp();
            FG_TT(1);
            if (lane == 0) tri_mbar_arrive(MB_EMPTY + (uint32_t)(tau & 1) * 8u); // this warp's coverage words of the step are written, its rows read
            if (++bstep == cfg.NB) bstep = 0;
        }
#ifdef FG_TRI_TIMING
        if (unit == 700 && lane == 0 && (warp == 0 || warp == 7)) printf("tri timing eval warp %d: T %d wait_full %lld eval %lld cycles/step\n", warp, T, tt_acc[0] / T, tt_acc[1] / T);
#endif
        return;
    }

    // =============================== loader warps ===============================
    {
        const int dw = warp - NE; // 0 .. ND - 1; dw == 0 places the groups and issues the copies, the others write the finished rows
        const int NG = cfg.NG;
        const uint32_t Ms = sbase + cfg.off_M;
        const uint32_t ZP = sbase + cfg.off_zp, PRAW = sbase + cfg.off_praw, GSs = sbase + cfg.off_gs;
        const uint32_t SINFO = sbase + cfg.off_sinfo, MINFO = sbase + cfg.off_minfo;
        uint4* ext = (uint4*)(smem + cfg.off_ext); // [2][NG] {Pg[first], Pg[last], rowbase lo, hi} of a group's source rows
        const int x = X0 + lane;
        const bool xvalid = x <= X1;
        const uint32_t nv8 = (uint32_t)PS >> 3;
        uint32_t slotP = 0u;   // ring slot of the first triple row of the next group to be placed (dw == 0)
        int lastbad = INT_MIN; // last triple row of the last skipped group (dw == 0)

        // extents of the source rows of group v (rows [Ja + D, Jb + D + 2)): cp.async, awaited before the group is placed
        auto issue_extents = [&](int v, int Ja, int Jb) {
            const int gA = Ja + D, NGv = Jb + D + 2 - gA;
            if (lane < NGv) {
                const long long trow = (long long)gA + lane - cfg.bm_j0;
                uint4* e = ext + (v & 1) * NG + lane;
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
        };
        // placement of group v (triple rows [Ja + D, Jb + D)) in the staging buffer and in the merged ring, bulk copies.
        // jt = first row of the step that is evaluated while the group is merged (older rows may be overwritten).
        auto place_group = [&](int v, int Ja, int Jb, int jt) {
            const uint32_t b = (uint32_t)v & 1u;
            const int gA = Ja + D, A_v = Jb - Ja, NGv = A_v + 2;
            uint32_t f = 0u, n = 0u, shift = 0u, cnt8 = 0u;
            uint64_t gal = 0ull;
            bool intab = false;
            if (lane < NGv) {
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
            const uint32_t Tr = lane < A_v ? n + n1 + n2 : 0u;
            ti = Tr;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t a1 = __shfl_up_sync(0xFFFFFFFFu, si, d), a2 = __shfl_up_sync(0xFFFFFFFFu, ti, d);
                if (lane >= d) { si += a1; ti += a2; }
            }
            const uint32_t stotal = __shfl_sync(0xFFFFFFFFu, si, 31), ttotal = __shfl_sync(0xFFFFFFFFu, ti, 31);
            bool fail = stotal > (uint32_t)cfg.GS || __any_sync(0xFFFFFFFFu, n > 60000u);
            // merged ring: rows are placed linearly; a group that would cross the end continues at 0 from its first row
            // that does not fit
            const uint32_t MC = (uint32_t)cfg.MCAP;
            const uint32_t head = state[0];
            uint32_t tail = head;
            if (jt < gA && jt >= Jbase) { // its ring slot: D rows (or less) behind the group's first row
                uint32_t st = slotP + (uint32_t)NQ - (uint32_t)(gA - jt);
                if (st >= (uint32_t)NQ) st -= (uint32_t)NQ;
                tail = tri_lds_u32(MINFO + st * 4u) & 0x7FFFFFFFu;
            }
            const uint32_t cross = __ballot_sync(0xFFFFFFFFu, lane < A_v && head + ti > MC);
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
#ifdef FG_TRI_DEBUG
            if (fail && lane == 0)
                printf("tri skip unit %d X0 %d Y0 %d group %d T %d uS %d A %d head %u tail %u ttotal %u stotal %u cross %x newhead %u jt %d gA %d Jbase %d\n", unit, X0, Y0, v, T, uS,
                       A_v, head, tail, ttotal, stotal, cross, newhead, jt, gA, Jbase);
#endif
            // a group that does not fit is SKIPPED: its triple rows take no room (marked in MINFO), nothing is copied or
            // merged, and the evaluation warps walk the cell table for the samples that start in them
            if (fail) { mstart = head; newhead = head; }
            // per source row: where its slices land, and the constant parts of the destinations of its grains
            const uint32_t f1 = __shfl_down_sync(0xFFFFFFFFu, f, 1), f2 = __shfl_down_sync(0xFFFFFFFFu, f, 2);
            const uint32_t fm1 = __shfl_up_sync(0xFFFFFFFFu, f, 1), fm2 = __shfl_up_sync(0xFFFFFFFFu, f, 2);
            const uint32_t msm1 = __shfl_up_sync(0xFFFFFFFFu, mstart, 1), msm2 = __shfl_up_sync(0xFFFFFFFFu, mstart, 2);
            const uint32_t so = GSs + (si - sz);
            const uint32_t praw = intab ? PRAW + (uint32_t)lane * PRB : ZP;
            if (lane < NGv) {
                const uint32_t sa = SINFO + (uint32_t)lane * 32u;
                tri_sts_v4(sa, praw, n, so + shift * 8u, so + cnt8 * 8u + shift * 2u);
                tri_sts_v4(sa + 16u, mstart - f1 - f2, msm1 - fm1 - f1, msm2 - fm2 - fm1, mstart - f - f1 - f2);
                if (lane < A_v) {
                    uint32_t sm = slotP + (uint32_t)lane;
                    if (sm >= (uint32_t)NQ) sm -= (uint32_t)NQ;
                    tri_sts_u32(MINFO + sm * 4u, mstart | (fail ? 0x80000000u : 0u));
                }
            }
            if (lane == 0) {
                state[3] = fail ? 1u : 0u;
                state[0] = newhead;
                // rows (relative to Jbase) of this group if it is skipped, else an empty range; eight groups back is further than
                // any window reaches
                state[32 + 2 * (v & 7)] = fail ? (uint32_t)(gA - Jbase) : 1u;
                state[33 + 2 * (v & 7)] = fail ? (uint32_t)(gA + A_v - 1 - Jbase) : 0u;
            }
            if (fail) lastbad = gA + A_v - 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (!fail && lane < NGv && intab) {
                const size_t row = (size_t)plane * cfg.bm_rows + (size_t)((long long)gA + lane - cfg.bm_j0);
                tri_mbar_arrive_tx(MB_COPY, PRB + sz);
                tri_bulk_g2s(praw, tab.Pg + row * cfg.ppitch + tcolA_u, PRB, MB_COPY);
                if (cnt8) {
                    tri_bulk_g2s(so, tab.Gg + gal, cnt8 * 8u, MB_COPY);
                    tri_bulk_g2s(so + cnt8 * 8u, tab.Cg + gal, cnt8 * 2u, MB_COPY);
                }
            } else {
                tri_mbar_arrive(MB_COPY);
            }
            slotP += (uint32_t)A_v;
            if (slotP >= (uint32_t)NQ) slotP -= (uint32_t)NQ;
        };

        int J0c = Jof(uS - 1), J1c = Jof(uS), J2c = Jof(uS + 1); // J(tau), J(tau + 1), J(tau + 2)
        uint32_t slotG = 0u; // ring slot of the first triple row of group tau + 1 (the one merged in iteration tau)
        int ydone = Y0;      // rows below are written
        if (dw == 0) { // prologue: the first group is placed before the loop, the extents of the second are on their way
            if (uS < T) {
                issue_extents(uS, J0c, J1c);
                cp_async_wait_all();
                __syncwarp();
                place_group(uS, J0c, J1c, Jw0);
                if (uS + 1 < T) issue_extents(uS + 1, J1c, J2c);
            }
        }
        FG_TT_DECL;
        for (int tau = uS - 1; tau <= T; ++tau) {
            const int u = tau + 1; // the group merged in this iteration, evaluated in step u
            FG_TT(7);
            const bool have = u >= uS && u < T;
            const int A_u = J1c - J0c, NGu = A_u + 2;
            // ---- the evaluation warps are done with step tau - 1: its coverage words are complete, and the ring slots and the
            //      part of M that the merge below overwrites are free ----
            FG_TT(2);
            if (tau >= 1) tri_mbar_wait(MB_EMPTY + (uint32_t)((tau - 1) & 1) * 8u, (uint32_t)((tau - 1) >> 1) & 1u);
            FG_TT(3);
            // ---- merge group u ----
            if (have) {
                tri_mbar_wait(MB_COPY, (uint32_t)(u - uS) & 1u);
                FG_TT(4);
                if (state[3] == 0u) {
                    // Q[d][e] = start of triple d + P[d][e] + P[d+1][e] + P[d+2][e] (each relative to its window start): eight
                    // 16-bit entries per item; warps over the triple rows, lanes over the 16-byte vectors of a row
                    for (int d = dw; d < A_u; d += ND) {
                        const uint32_t sa = SINFO + (uint32_t)d * 32u;
                        const uint32_t p0 = tri_lds_u32(sa), p1 = tri_lds_u32(sa + 32u), p2 = tri_lds_u32(sa + 64u);
                        const uint32_t qb = tri_lds_u32(sa + 28u);
                        uint32_t sl = slotG + (uint32_t)d;
                        if (sl >= (uint32_t)NQ) sl -= (uint32_t)NQ;
                        const uint32_t qrow = Qs + sl * PSB;
                        for (uint32_t v = (uint32_t)lane; v < nv8; v += 32u) {
                            const uint4 a0 = tri_lds_v4(p0 + v * 32u), a1 = tri_lds_v4(p0 + v * 32u + 16u);
                            const uint4 b0 = tri_lds_v4(p1 + v * 32u), b1 = tri_lds_v4(p1 + v * 32u + 16u);
                            const uint4 c0 = tri_lds_v4(p2 + v * 32u), c1 = tri_lds_v4(p2 + v * 32u + 16u);
                            const uint32_t q0 = a0.x + b0.x + c0.x + qb, q1 = a0.y + b0.y + c0.y + qb, q2 = a0.z + b0.z + c0.z + qb, q3 = a0.w + b0.w + c0.w + qb;
                            const uint32_t q4 = a1.x + b1.x + c1.x + qb, q5 = a1.y + b1.y + c1.y + qb, q6 = a1.z + b1.z + c1.z + qb, q7 = a1.w + b1.w + c1.w + qb;
                            tri_sts_v4(qrow + v * 16u, (q0 & 0xFFFFu) | (q1 << 16), (q2 & 0xFFFFu) | (q3 << 16), (q4 & 0xFFFFu) | (q5 << 16),
                                       (q6 & 0xFFFFu) | (q7 << 16));
                        }
                    }
                    // grains: warps over the source rows, lanes over a row's grains (two per lane and trip: their loads overlap).
                    // Inside a column of triple d the grains of its CENTRE row d + 1 come first (a sample's own cell row: the
                    // candidates most likely to cover it are tested first), then row d, then row d + 2.  A grain with window-local
                    // index g in window column e therefore goes to
                    //   triple r     at  start(r)   - f(r+1) - f(r+2) + P[r+1][e+1] + P[r+2][e]   + g   (its row is the triple's first),
                    //   triple r - 1 at  start(r-1) - f(r-1) - f(r+1) + P[r-1][e]   + P[r+1][e]   + g   (centre),
                    //   triple r - 2 at  start(r-2) - f(r-2) - f(r-1) + P[r-2][e+1] + P[r-1][e+1] + g   (third).
                    for (int r = dw; r < NGu; r += ND) {
                        const uint32_t sa = SINFO + (uint32_t)r * 32u;
                        const uint4 ri = tri_lds_v4(sa), bi = tri_lds_v4(sa + 16u);
                        const bool t0 = r < A_u, t1 = r >= 1 && r - 1 < A_u, t2 = r >= 2;
                        const uint32_t pp1 = (t0 || t1) ? tri_lds_u32(sa + 32u) : ZP, pp2 = t0 ? tri_lds_u32(sa + 64u) : ZP;
                        const uint32_t pm1 = (t1 || t2) ? tri_lds_u32(sa - 32u) : ZP, pm2 = t2 ? tri_lds_u32(sa - 64u) : ZP;
                        for (uint32_t g0 = (uint32_t)lane; g0 < ri.y; g0 += 64u) {
                            const uint32_t g1 = g0 + 32u;
                            const bool v1 = g1 < ri.y;
                            const uint32_t c0 = tri_lds_u16(ri.w + g0 * 2u), c1 = v1 ? tri_lds_u16(ri.w + g1 * 2u) : c0;
                            const float2 gr0 = lds_f32x2(ri.z + g0 * 8u), gr1 = lds_f32x2(ri.z + (v1 ? g1 : g0) * 8u);
                            const uint32_t e0 = ((c0 - tcolA_u) & 0xFFFFu) * 4u, e1 = ((c1 - tcolA_u) & 0xFFFFu) * 4u; // byte offsets of the grains' window columns
                            const uint32_t a10 = tri_lds_u32(pp1 + e0), b10 = tri_lds_u32(pp1 + e0 + 4u), a20 = tri_lds_u32(pp2 + e0);
                            const uint32_t m10 = tri_lds_u32(pm1 + e0), n10 = tri_lds_u32(pm1 + e0 + 4u), n20 = tri_lds_u32(pm2 + e0 + 4u);
                            const uint32_t a11 = tri_lds_u32(pp1 + e1), b11 = tri_lds_u32(pp1 + e1 + 4u), a21 = tri_lds_u32(pp2 + e1);
                            const uint32_t m11 = tri_lds_u32(pm1 + e1), n11 = tri_lds_u32(pm1 + e1 + 4u), n21 = tri_lds_u32(pm2 + e1 + 4u);
                            if (t0) tri_sts_f32x2(Ms + (bi.x + g0 + b10 + a20) * 8u, gr0.y, gr0.x);
                            if (t1) tri_sts_f32x2(Ms + (bi.y + g0 + m10 + a10) * 8u, gr0.y, gr0.x);
                            if (t2) tri_sts_f32x2(Ms + (bi.z + g0 + n20 + n10) * 8u, gr0.y, gr0.x);
                            if (v1) {
                                if (t0) tri_sts_f32x2(Ms + (bi.x + g1 + b11 + a21) * 8u, gr1.y, gr1.x);
                                if (t1) tri_sts_f32x2(Ms + (bi.y + g1 + m11 + a11) * 8u, gr1.y, gr1.x);
                                if (t2) tri_sts_f32x2(Ms + (bi.z + g1 + n21 + n11) * 8u, gr1.y, gr1.x);
                            }
                        }
                    }
                }
            }
            // ---- the loader warps agree that the group is merged (and the staging buffer free), then publish it ----
            FG_TT(5);
            asm volatile("bar.sync 1, %0;" ::"n"(ND * 32) : "memory");
            FG_TT(6);
            if (u >= 0 && u < T) {
                if (dw == 0 && lane == 0) state[4 + (u & 1)] = lastbad >= J1c ? 1u : 0u; // J1c = J(u): first row of step u's window
                __syncwarp();
                if (lane == 0) tri_mbar_arrive(MB_FULL + (uint32_t)(u & 1) * 8u);
            }
            if (dw == 0) {
                // ---- the next group: placement and copies (they land while the other loader warps write rows and this warp
                //      waits for the evaluation warps), then the extents of the one after it ----
                if (u + 1 >= uS && u + 1 < T) {
                    cp_async_wait_all(); // the extents of group u + 1
                    __syncwarp();
                    place_group(u + 1, J1c, J2c, tau + 1 >= 0 ? J1c : Jw0);
                    if (u + 2 < T) issue_extents(u + 2, J2c, Jof(tau + 3));
                }
            } else if (tau >= 1) {
                // ---- rows every sample has passed (after the group is published: not on the critical path) ----
                const int ynew = min(Yref0 + tau * m + skmin, Y1);
                for (int yy = ydone + dw - 1; yy < ynew; yy += ND - 1) {
                    const uint32_t hrow = HB + ((uint32_t)(yy - Y0) & (uint32_t)(cfg.RH - 1)) * HBROW + (uint32_t)lane * 4u;
                    uint32_t cnt = 0;
#pragma unroll 4
                    for (int cw = 0; cw < (int)(HBROW / 128u); ++cw) {
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
            }
            if (have) {
                slotG += (uint32_t)A_u;
                if (slotG >= (uint32_t)NQ) slotG -= (uint32_t)NQ;
            }
            J0c = J1c; J1c = J2c; J2c = Jof(tau + 3);
        }
#ifdef FG_TRI_TIMING
        if (unit == 700 && lane == 0) printf("tri timing loader warp %d: wait_empty %lld wait_copy %lld merge %lld bar %lld placement/finalise+rest %lld cycles/step\n", dw, tt_acc[3] / T, tt_acc[4] / T, tt_acc[5] / T, tt_acc[6] / T, (tt_acc[7] + tt_acc[2]) / T);
#endif
    }
}

} // namespace fg
