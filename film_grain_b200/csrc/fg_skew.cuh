// fg_skew.cuh -- k_pixelwise_skew: the pixel-wise evaluation kernel for the default geometry rm == delta
// (cell_delta = 1 / ceil(1 / r), constant radius: every BASELINE pixel-wise config except r = 0.12).
//
// Reference semantics (src/pixelwise.rs:47-106): a sample point visits the cells [i0, i1] x [j0, j1] within rm of
// it -- with rm == delta exactly three cell rows -- and is covered if any grain of those cells is within its radius.
// The cell table (fg_stage.cuh) already holds every cell's grains, generated once.  What this kernel changes with
// respect to k_pixelwise_strip (fg_tile.cuh) is how a CTA walks its strip of 32 output columns:
//
//   * SKEWED rows.  The strip kernel evaluates all N samples of a pixel row against one window, so the window
//     must span the whole spread of the sample offsets (about 52 cell rows at sigma = 0.8 px, r = 0.1) on top of
//     the rows it advances by.  Here a step is defined in CELL rows: step t owns the cell rows
//     [J + t D, J + (t + 1) D) and every sample k evaluates, at step t, exactly those pixel rows y whose first
//     cell row j0(y, k) falls into it -- each sample is at its own pixel row ("skew" = its y offset).  Every
//     (pixel, sample) pair is evaluated in exactly one step, and a step only needs D + 2 cell rows.
//   * MERGED triples.  With so few rows per step the window can afford a layout in which a sample's whole
//     3 x 3 cell block is ONE contiguous list: for every cell row j of the step the rows j, j + 1, j + 2 are
//     merged column by column (grains of cell (i, j), (i, j + 1), (i, j + 2), then column i + 1, ...), with a
//     16-bit prefix Q[j][i] = first list entry of column i.  A sample reads Q[j0][i0] and Q[j0][i1 + 1] and walks
//     that range: two prefix loads instead of six, no slot -> row selection, immediate-offset grain loads,
//     predicated on the range length so that no lane loads a grain that is not its own.
//     Every grain is stored three times (it belongs to three triples); the loader computes
//     Q = P[j] + P[j+1] + P[j+2] (packed 16-bit adds) and scatters each grain of the D + 2 source rows to its
//     three places with four prefix lookups (fg_stage.cuh writes the cell column of every grain for this).
//   * Double buffering: while step t is evaluated, step t + 1 is merged into the other buffer and the prefix
//     rows of step t + 2 are fetched; one CTA barrier per step.
//   * Pixel counts live in per-warp 8-bit rings (plain load / add / store, no atomics); a row is summed over
//     the warps and written once every sample's cursor has passed it.
//
// Samples that visit another number of cell rows than three (f32 rounding of (y -/+ rm) / delta can give two or
// four) are evaluated from the HBM table like k_pixelwise_table_tiles does; a strip segment whose merged window
// does not fit shared memory (dense content) goes to the fallback list.  Results are bit-identical to
// k_pixelwise_strip / k_pixelwise_direct / the oracle: the visited cell set, the f32 operations of the distance
// test and the count are the reference's; only the order of the (commutative) "any grain covers" changes.
#pragma once
#include "fg_tile.cuh"

namespace fg {

#define FG_SK_WARPS 32
#define FG_SK_THREADS (FG_SK_WARPS * 32)
#ifndef FG_SK_SLOTS
#define FG_SK_SLOTS 6 // straight-line predicated grain tests per sample before the remainder loop
#endif
#define FG_SK_DMAX 30 // triple rows per step (D + 2 source rows are scanned by one warp register per lane)

struct SkewCfg {
    int D, S;             // triple rows per step, source rows per step (D + 2)
    int CWB, PS;          // bound on the window's cell columns, prefix row stride (u16 entries, multiple of 8)
    int MCAP;             // merged grains per buffer
    int TCAP;             // of which per triple: triple d owns [d TCAP, (d + 1) TCAP)
    int R;                // rows of the per-warp pixel-count ring (power of two); slot R is the dump row
    int SEG, n_strips, n_segs;
    int bm_i0, bm_j0, bm_cols, bm_rows; // the cell table's rectangle
    uint32_t ppitch;
    float r2c;
    uint32_t off_Ps, off_Q, off_zero, off_info, off_tb, off_wp, off_pcw, off_pcc, off_sync, off_M, total;
};

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

template <int SPWC>
__global__ void __launch_bounds__(FG_SK_THREADS, 1)
k_pixelwise_skew(const float* __restrict__ lambda, size_t in_stride, const float2* __restrict__ offsets_input,
                 float* __restrict__ out, size_t out_stride, TileRef* __restrict__ fb_list, uint32_t* __restrict__ fb_count,
                 uint32_t fb_cap, SkewCfg cfg, RenderConsts c, CellTable tab) {
    constexpr int NW = FG_SK_WARPS;
    constexpr int CAND = 32 / SPWC;           // candidate pixel rows per sample and item round
    constexpr uint32_t CMASK = (1u << CAND) - 1u;
    extern __shared__ __align__(16) unsigned char smem[];
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
    const int D = cfg.D, S = cfg.S, PS = cfg.PS, R = cfg.R;

    // ---- the strip's cell window: columns (monotone in x and in the offset), first and last first-cell-row ----
    const float bx0 = __fmul_rn(__fadd_rn((float)X0, 0.5f), c.inv_zoom);
    const float bx1 = __fmul_rn(__fadd_rn((float)X1, 0.5f), c.inv_zoom);
    const int i_lo = cell_lo(__fsub_rn(bx0, c.off_max_x), rm, delta);
    const int i_hi = cell_hi(__fsub_rn(bx1, c.off_min_x), rm, delta);
    const float byF = __fmul_rn(__fadd_rn((float)Y0, 0.5f), c.inv_zoom);
    const float byL = __fmul_rn(__fadd_rn((float)(Y1 - 1), 0.5f), c.inv_zoom);
    const int J0 = cell_lo(__fsub_rn(byF, c.off_max_y), rm, delta);     // smallest j0 of any (row, sample) of the segment
    const int JL = cell_lo(__fsub_rn(byL, c.off_min_y), rm, delta);     // largest
    // prefix rows are fetched with 16-byte loads from the table column rounded down to a multiple of four: the
    // window simply starts there (up to three extra cells on the left)
    const long long tcol0 = (long long)i_lo - cfg.bm_i0;
    const long long tcolA = tcol0 & ~3LL;
    const long long CWl = (long long)i_hi - cfg.bm_i0 - tcolA + 1; // window columns
    bool geo_bad = i_lo > i_hi || tcol0 < 0 || (long long)i_hi >= (long long)cfg.bm_i0 + cfg.bm_cols || CWl > cfg.CWB ||
                   J0 < cfg.bm_j0 || (long long)JL - J0 > 2000000000LL || !radius_ok;
    if (geo_bad) { // uniform
        if (!radius_ok) { // nothing ever covers: zeros (src/pixelwise.rs:93-95)
            for (int p = tid; p < (Y1 - Y0) * 32; p += FG_SK_THREADS)
                if (X0 + (p & 31) <= X1) outp[(size_t)(Y0 + (p >> 5)) * c.out_w + X0 + (p & 31)] = 0.0f;
        } else if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
        return;
    }
    const int CW = (int)CWl;
    const int i_loA = cfg.bm_i0 + (int)tcolA;
    const uint32_t tcolA_u = (uint32_t)tcolA;
    const int T = (JL - J0) / D + 1; // steps: step t owns the first-cell-rows [J0 + t D, J0 + (t + 1) D)
    const int n4 = (CW + 1 + 3) >> 2; // 16-byte prefix vectors per row (CW + 1 entries)

    // shared-memory windows (32-bit shared addresses)
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t Pss = sbase + cfg.off_Ps, Qs = sbase + cfg.off_Q, zero_s = sbase + cfg.off_zero;
    const uint32_t Ms = sbase + cfg.off_M;
    uint32_t wps = sbase + cfg.off_wp + (uint32_t)warp * 512u; // 32 items of 16 bytes per warp
    uint32_t pcws = sbase + cfg.off_pcw + (uint32_t)warp * (uint32_t)(R + 1) * 32u + (uint32_t)lane;
    asm volatile("" : "+r"(wps), "+r"(pcws)); // opaque: kept in registers instead of being re-derived from the thread id in every loop
    uint4* info = (uint4*)(smem + cfg.off_info);       // [2][32] {first prefix of the row's window, grains in it, global grain index lo, hi}
    uint32_t* pcc = (uint32_t*)(smem + cfg.off_pcc);    // [SEG][32] counts over the sample chunks (only when there are several)
    int* ysync = (int*)(smem + cfg.off_sync);           // [0..2] min over the warps of the samples' cursors, per step mod 3; [3] fallback flag
    const uint32_t PSB = (uint32_t)PS * 2u;             // bytes per prefix row
    const uint32_t PsBuf = (uint32_t)S * PSB, QBuf = (uint32_t)D * PSB, MBuf = (uint32_t)cfg.MCAP * 8u;

    const int n_chunks = (int)((c.n + NW * SPWC - 1) / (NW * SPWC));
    for (int p = tid; p < PS / 2; p += FG_SK_THREADS) ((uint32_t*)(smem + cfg.off_zero))[p] = 0u;
    for (int p = tid; p < NW * (R + 1) * 8; p += FG_SK_THREADS) ((uint32_t*)(smem + cfg.off_pcw))[p] = 0u;

    const int x = X0 + lane;
    const bool xvalid = x <= X1;
    const float bx = __fmul_rn(__fadd_rn((float)x, 0.5f), c.inv_zoom);

    // ---- phase A of step u: prefix rows of its S source rows -> Ps[u & 1] (16-bit, relative to the window start) and the
    //      row info.  NSPLIT warps share a row (uniform row address arithmetic), one 16-byte vector = four entries per lane. ----
    const int NSPLIT = NW / S > 0 ? NW / S : 1; // S <= 32
    const int wrow = warp / NSPLIT, wpart = warp - wrow * NSPLIT;
    auto phase_a = [&](int u) {
        const uint32_t buf = (uint32_t)u & 1u;
        if (wrow >= S) return;
        const int r = wrow;
        const long long trow = (long long)(J0 + u * D) + r - cfg.bm_j0;
        uint2* prow = (uint2*)(smem + cfg.off_Ps + buf * PsBuf + (uint32_t)r * PSB);
        if (trow < 0 || trow >= cfg.bm_rows) { // beyond the table: an empty row (no sample's fast path reads it)
            for (int q4 = wpart * 32 + lane; q4 < n4; q4 += 32 * NSPLIT) prow[q4] = make_uint2(0u, 0u);
            if (wpart == 0 && lane == 0) info[buf * 32 + r] = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
        const size_t row = (size_t)plane * cfg.bm_rows + (size_t)trow;
        const uint32_t* pg = tab.Pg + row * cfg.ppitch + tcolA_u;
        const uint4* pr4 = (const uint4*)pg;
        const uint32_t a = __ldg(pg);
        for (int q4 = wpart * 32 + lane; q4 < n4; q4 += 32 * NSPLIT) {
            const uint4 pv = __ldg(pr4 + q4);
            uint2 o;
            o.x = ((pv.x - a) & 0xFFFFu) | ((pv.y - a) << 16);
            o.y = ((pv.z - a) & 0xFFFFu) | ((pv.w - a) << 16);
            prow[q4] = o;
        }
        if (wpart == 0 && lane == 0) {
            const uint64_t gsrc = __ldg(tab.rowbase + row) + a;
            info[buf * 32 + r] = make_uint4(a, __ldg(pg + CW) - a, (uint32_t)gsrc, (uint32_t)(gsrc >> 32));
        }
    };

    // ---- phase B + C of step u: merged prefixes Q[u & 1] and the grains of the source rows, each to its three
    //      triples.  Triple d owns the fixed range [d TCAP, (d + 1) TCAP) of the merged buffer, so nothing has to be
    //      scanned.  Returns false (uniformly over the CTA) when a triple does not fit its range. ----
    const uint32_t TCAP = (uint32_t)cfg.TCAP;
    auto phase_bc = [&](int u) -> bool {
        const uint32_t buf = (uint32_t)u & 1u;
        { // every warp reaches the verdict itself from the same S row counts: lane r holds source row r (S <= 32)
            const uint32_t nr = (lane < S) ? info[buf * 32 + lane].y : 0u;
            const uint32_t n1 = __shfl_down_sync(0xFFFFFFFFu, nr, 1), n2 = __shfl_down_sync(0xFFFFFFFFu, nr, 2);
            const bool over = lane < D && nr + n1 + n2 > TCAP; // lane < D <= 30: lanes + 1, + 2 exist
            if (__any_sync(0xFFFFFFFFu, over || nr > 65535u)) return false;
        }
        // B: Q[d][e] = d TCAP + Ps[d][e] + Ps[d+1][e] + Ps[d+2][e]; eight 16-bit entries (16 bytes) per item, packed adds
        // (a valid entry never exceeds D TCAP <= 65535, so no carry leaves a valid half-word)
        const int nv = PS >> 3, nitems = D * nv;
        for (int item = tid; item < nitems; item += FG_SK_THREADS) {
            const int d = item / nv, v = item - d * nv;
            const uint32_t b3 = (uint32_t)d * TCAP * 0x10001u;
            const uint32_t src = Pss + buf * PsBuf + (uint32_t)d * PSB + (uint32_t)v * 16u;
            uint4 p0, p1, p2;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(p0.x), "=r"(p0.y), "=r"(p0.z), "=r"(p0.w) : "r"(src));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(p1.x), "=r"(p1.y), "=r"(p1.z), "=r"(p1.w) : "r"(src + PSB));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(p2.x), "=r"(p2.y), "=r"(p2.z), "=r"(p2.w) : "r"(src + 2u * PSB));
            const uint32_t dst = Qs + buf * QBuf + (uint32_t)d * PSB + (uint32_t)v * 16u;
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(p0.x + p1.x + p2.x + b3), "r"(p0.y + p1.y + p2.y + b3),
                         "r"(p0.z + p1.z + p2.z + b3), "r"(p0.w + p1.w + p2.w + b3) : "memory");
        }
        // C: NSPLIT warps per source row r, lanes over its grains.  A grain with row-local index g in window column e goes to
        //   triple r   at  r      TCAP + Ps[r+1][e]   + Ps[r+2][e]   + g   (its row is the triple's first),
        //   triple r-1 at (r - 1) TCAP + Ps[r-1][e+1] + Ps[r+1][e]   + g   (second),
        //   triple r-2 at (r - 2) TCAP + Ps[r-2][e+1] + Ps[r-1][e+1] + g   (third).
        if (wrow < S) {
            const uint32_t r = (uint32_t)wrow;
            const uint4 ri = info[buf * 32 + r]; // one address for the warp: a broadcast
            const size_t gsrc = (size_t)(((uint64_t)ri.w << 32) | ri.z);
            const uint16_t* cp = tab.Cg + gsrc;
            const float2* gp = tab.Gg + gsrc;
            const uint32_t prow = Pss + buf * PsBuf + r * PSB;
            const uint32_t mrow = Ms + buf * MBuf + r * TCAP * 8u;
            const bool t0 = r < (uint32_t)D, t1 = r >= 1u && r <= (uint32_t)D, t2 = r >= 2u; // uniform
            const uint32_t o1 = t1 ? TCAP * 8u : 0u, o2 = 2u * TCAP * 8u;
            for (uint32_t g = (uint32_t)(wpart * 32 + lane); g < ri.y; g += 32u * (uint32_t)NSPLIT) {
                const uint32_t e2 = (((uint32_t)__ldg(cp + g) - tcolA_u) & 0xFFFFu) * 2u; // byte offset of the grain's window column
                const float2 gr = __ldg(gp + g);
                const uint32_t pe = prow + e2; // &Ps[r][e]
                const uint32_t x1 = (t0 || t1) ? lds_u16(pe + PSB) : 0u;       // Ps[r+1][e]
                const uint32_t x2 = t0 ? lds_u16(pe + 2u * PSB) : 0u;           // Ps[r+2][e]
                const uint32_t y1 = (t1 || t2) ? lds_u16(pe - PSB + 2u) : 0u;   // Ps[r-1][e+1]
                const uint32_t y2 = t2 ? lds_u16(pe - 2u * PSB + 2u) : 0u;      // Ps[r-2][e+1]
                const uint32_t m = mrow + g * 8u;
                if (t0) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(m + (x1 + x2) * 8u), "f"(gr.x), "f"(gr.y) : "memory");
                if (t1) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(m + (y1 + x1) * 8u - o1), "f"(gr.x), "f"(gr.y) : "memory");
                if (t2) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(m + (y2 + y1) * 8u - o2), "f"(gr.x), "f"(gr.y) : "memory");
            }
        }
        return true;
    };

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        // ---- per-lane sample data of this chunk: abscissa and packed byte offsets of Q[.][i0], Q[.][i1 + 1] ----
        float xg_r[SPWC];
        uint32_t ip_r[SPWC];
#pragma unroll
        for (int s = 0; s < SPWC; ++s) {
            const uint32_t k = (uint32_t)chunk * (NW * SPWC) + s * NW + warp;
            float xg = 0.0f;
            uint32_t ip = 0u; // a2 == b2: an empty range (inactive lane / sample)
            if (k < c.n && xvalid) {
                xg = __fsub_rn(bx, __ldg(offsets_input + k).x);
                ip = col_range_packed(xg, rm, delta, i_loA);
            }
            xg_r[s] = xg;
            ip_r[s] = ip;
        }
        // item builder: lane -> (sample ls of this warp, candidate row offset ldy); the cursor of a sample is kept by its lanes
        const int ls = lane / CAND, ldy = lane - ls * CAND;
        const uint32_t lk = (uint32_t)chunk * (NW * SPWC) + ls * NW + warp;
        const bool lvalid = lk < c.n;
        const float loy = lvalid ? __ldg(offsets_input + lk).y : 0.0f;
        int ycur = Y0;
        int ydone = Y0; // rows below are final (all samples of all warps have passed them) and written

        if (tid < 3) ysync[tid] = INT_MAX;
        if (tid == 3) ysync[3] = 0; // "segment handed to the fallback kernel" (see the slow items)
        __syncthreads(); // (zero row / rings initialised; previous chunk finished)

        // software pipeline, one barrier per iteration: prefix rows of step t + 2 and merge of step t + 1 (the loader), and
        // the evaluation of step t (t = -2, -1 are the prologue).  The two halves are independent, so odd warps evaluate
        // first and load second, even warps the other way round: the global-memory latency of one half of the warps
        // is covered by the arithmetic of the other half.
        for (int t = -2; t < T; ++t) {
            const uint32_t buf = (uint32_t)t & 1u;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                if ((pass == 0) != ((warp & 1) != 0)) { // ---- loader ----
                    if (t + 2 < T) phase_a(t + 2);
                    if (t + 1 >= 0 && t + 1 < T && !phase_bc(t + 1)) { // uniform over the CTA
                        if (tid == 0) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
                        return;
                    }
                    continue;
                }
                if (t < 0) continue;
                // =================== evaluation of step t ===================
                const int Jlo = J0 + t * D;
                const long long Jhi = (long long)Jlo + D;
                const uint32_t Mb = Ms + buf * MBuf, Qb = Qs + buf * QBuf;
                bool again;
                do {
                    // ---- items: the pixel rows each sample of this warp evaluates in this step ----
                    const int y = ycur + ldy;
                    const bool ok = lvalid && y < Y1;
                    float yg = 0.0f;
                    int j0 = 0, j1 = 0;
                    if (ok) {
                        yg = __fsub_rn(__fmul_rn(__fadd_rn((float)y, 0.5f), c.inv_zoom), loy);
                        j0 = cell_lo(yg, rm, delta);
                        j1 = cell_hi(yg, rm, delta);
                    }
                    const bool take = ok && (long long)j0 < Jhi; // monotone in y: the taken rows of a sample are a prefix of its candidates
                    const bool fast = take && (j1 - j0 == 2) && (long long)j0 + 2 < (long long)cfg.bm_j0 + cfg.bm_rows;
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, take);
                    const uint32_t slowbal = __ballot_sync(0xFFFFFFFFu, take && !fast);
                    // item = {y of the sample point, shared address of its Q row, byte offset of its pixel-count row, -}
                    uint32_t qa = zero_s, slot = (uint32_t)R; // not taken / slow: an empty range, counted into the dump row
                    if (take) slot = (uint32_t)(y - Y0) & (uint32_t)(R - 1);
                    if (fast) qa = Qb + (uint32_t)(j0 - Jlo) * PSB;
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wps + (uint32_t)lane * 16u), "r"(__float_as_uint(yg)), "r"(qa), "r"(slot * 32u), "r"(0u) : "memory");
                    const uint32_t mycnt = __popc((bal >> (ls * CAND)) & CMASK);
                    ycur += (int)mycnt;
                    again = __any_sync(0xFFFFFFFFu, mycnt == (uint32_t)CAND);
                    __syncwarp();
                    // ---- the fast items: one contiguous candidate list per (lane, item) ----
                    {
                        // loop invariants pinned in registers (left alone the compiler re-derives them from the thread id and the
                        // parameter block for every item)
                        uint32_t wpr = wps, Mbr = Mb, pcr = pcws;
                        float r2r = r2;
                        asm volatile("" : "+r"(wpr), "+r"(Mbr), "+r"(pcr), "+f"(r2r));
                        uint64_t gch[3] = {0ull, 0ull, 0ull}; // slot chains: any defined value (see slot_test)
    #pragma unroll
                        for (int s = 0; s < SPWC; ++s) {
                            const uint32_t cnt = __popc((bal >> (s * CAND)) & CMASK); // uniform
                            const uint32_t a2 = ip_r[s] & 0xFFFFu, b2 = ip_r[s] >> 16;
                            const float xg = xg_r[s];
                            uint32_t ia = wpr + (uint32_t)(s * CAND) * 16u;
    #pragma unroll 1
                            for (uint32_t r = 0; r < cnt; ++r, ia += 16u) {
                                uint32_t w0, q, po, w3;
                                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(q), "=r"(po), "=r"(w3) : "r"(ia));
                                const uint32_t s16 = lds_u16(q + a2), e16 = lds_u16(q + b2);
                                const uint32_t n = e16 - s16;
                                const uint32_t ga = Mbr + s16 * 8u;
                                const uint64_t pp = pack_f32x2(xg, __uint_as_float(w0));
                                float dmin = __int_as_float(0x7f800000);
                                SlotRun<0, FG_SK_SLOTS>::run(dmin, gch, n, ga, pp);
                                if (n > FG_SK_SLOTS && !(dmin <= r2r)) { // remainder: exits on the first hit
                                    uint32_t u = FG_SK_SLOTS;
                                    do {
                                        const float d2 = dist2_packed(pp, lds_f32x2(ga + u * 8u));
                                        if (d2 <= r2r) { dmin = d2; break; }
                                    } while (++u < n);
                                }
                                const uint32_t pa = pcr + po;
                                const uint32_t cv = lds_u8(pa);
                                sts_u8(pa, cv + ((dmin <= r2r) ? 1u : 0u));
                            }
                        }
                    }
                    // ---- the rare items that do not visit exactly three cell rows (f32 rounding of (y -/+ rm) / delta gives two
                    //      or four): walk their cells in the HBM table, like k_pixelwise_table_tiles.  Inline on purpose: a
                    //      call here costs the hot loop its registers.  Cell rows outside the table cannot occur for planned
                    //      geometry; if they do, the segment is handed to the fallback kernel, which overwrites it. ----
                    uint32_t sb = slowbal;
                    while (sb) { // uniform
                        const int L = __ffs(sb) - 1;
                        sb &= sb - 1;
                        uint32_t w0, w1, po, w3;
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(po), "=r"(w3) : "r"(wps + (uint32_t)L * 16u));
                        const float ygs = __uint_as_float(w0);
                        const int sj0 = cell_lo(ygs, rm, delta), sj1 = cell_hi(ygs, rm, delta);
                        if (sj0 < cfg.bm_j0 || (long long)sj1 >= (long long)cfg.bm_j0 + cfg.bm_rows) {
                            if (lane == 0 && atomicExch(&ysync[3], 1) == 0) push_fallback(fb_list, fb_count, fb_cap, X0, Y0, X1 - X0 + 1, Y1 - Y0, plane);
                            continue;
                        }
                        if (!xvalid) continue;
                        const uint32_t k = (uint32_t)chunk * (NW * SPWC) + (uint32_t)(L / CAND) * NW + warp;
                        const float xgs = __fsub_rn(bx, __ldg(offsets_input + k).x);
                        const int si0 = cell_lo(xgs, rm, delta), si1 = cell_hi(xgs, rm, delta);
                        uint32_t hit = 0u;
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
                        const uint32_t pa = pcws + po;
                        sts_u8(pa, lds_u8(pa) + hit);
                    }
                    __syncwarp(); // the items are rewritten by the next round
                } while (again);
            }
            if (t < 0) { __syncthreads(); continue; }
            // the slowest cursor of the CTA decides which rows are final after this step
            {
                const int wmin = __reduce_min_sync(0xFFFFFFFFu, lvalid ? ycur : INT_MAX);
                if (lane == 0) atomicMin(&ysync[t % 3], wmin);
            }
            __syncthreads();
            // ---- rows every sample has passed: sum the warps' counters, write (or accumulate over the chunks), clear ----
            {
                const int ynew = min(ysync[t % 3], Y1);
                if (tid == 0) ysync[(t + 2) % 3] = INT_MAX; // written again in step t + 2, last read in step t (before this barrier)
                for (int yy = ydone + warp; yy < ynew; yy += NW) {
                    const uint32_t sl = (uint32_t)(yy - Y0) & (uint32_t)(R - 1);
                    uint32_t a = sbase + cfg.off_pcw + sl * 32u + (uint32_t)lane, sum = 0;
#pragma unroll 8
                    for (int w = 0; w < NW; ++w) {
                        sum += lds_u8(a);
                        sts_u8(a, 0u);
                        a += (uint32_t)(R + 1) * 32u;
                    }
                    if (n_chunks > 1) {
                        uint32_t* pc = pcc + (size_t)(yy - Y0) * 32 + lane;
                        if (chunk > 0) sum += *pc;
                        if (chunk + 1 < n_chunks) { *pc = sum; continue; }
                    }
                    if (xvalid) outp[(size_t)yy * c.out_w + x] = __fmul_rn((float)sum, c.inv_samples);
                }
                ydone = max(ydone, ynew);
            }
        }
    }
}

} // namespace fg
