// fg_pixel_host.cuh -- host side of the pixel-wise fast path: geometry plans of the two evaluation kernels
// (k_pixelwise_skew, fg_skew.cuh; k_pixelwise_strip, fg_tile.cuh), the per-band pipeline
//   thresholds -> first-draw bitmap -> row capacities -> cell table -> evaluation -> fallback list
// and the split of a render into row sub-bands (table budget, cancellation points).  Host code only.
#pragma once
#include "fg_skew.cuh"
#include "fg_tri.cuh"

namespace {
using namespace fg;

struct TilePlan { TileCfg cfg; int spwc; bool ok; };

inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// Choose the strip geometry for a render; ok == false -> the tiled path does not apply.
TilePlan tile_plan(const fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int n_planes, bool staged) {
    TilePlan pl{};
    pl.ok = false;
    if (p->n_samples > (1u << 20)) return pl;
    const double inv_zoom = 1.0 / (double)p->zoom, delta = p->delta, rm = p->rm;
    const double ox = (double)c.off_max_x - (double)c.off_min_x, oy = (double)c.off_max_y - (double)c.off_min_y;
    const double cwb = (31.0 * inv_zoom + ox + 2.0 * rm) / delta + 4.0;
    if (!(cwb < 2040.0)) return pl;
    if (!(2.0 * rm / delta + 3.0 < 250.0)) return pl;   // cell rows per sample are packed in 8 bits
    const int CWB = (int)cwb;
    const int PS = staged ? (CWB + 8 + 3) / 4 * 4 : CWB + 2; // staged: + column shift (<= 3) + vector tail (<= 3), multiple of 4
    const int R = std::max(1, std::min(15, FG_TILE_CELLS / (CWB + 1)));
    const int spwc = p->n_samples <= 4u * FG_TILE_WARPS ? 4 : (p->n_samples <= 8u * FG_TILE_WARPS ? 8 : FG_TILE_SPW_MAX);
    const int band = c.row_end - c.row_begin;
    const size_t smem_max = ctx->smem_optin;
    const uint32_t bpg = c.rad.lognorm ? 12u : 8u;      // bytes per grain: (cx, cy) [+ r^2]
    // Largest step height whose window fits; the grain ring takes all remaining shared memory.
    // pass 0 insists that the ring holds the window at a plausible density (0.45 grains per cell;
    // iid-uniform 8-bit input averages 1/pi); pass 1 takes anything that fits -- denser content
    // overflows into the fallback list at run time.
    for (int pass = 0; pass < 2 && !pl.ok; ++pass) {
        static const int th_candidates[] = {32, 24, 16, 12, 8, 6, 4, 3, 2, 1};
        static const int th_forced = std::getenv("FG_B200_TH") ? std::atoi(std::getenv("FG_B200_TH")) : 0; // experiments
        for (int TH : th_candidates) {
            if (pl.ok) break;
            if (th_forced > 0 && TH != th_forced) continue;
            if (TH > 1 && TH > 2 * band) continue;
            const double rhb = ((TH - 1) * inv_zoom + oy + 2.0 * rm) / delta + 4.0;
            if (!(rhb < 30000.0)) continue;
            const int RH = (int)rhb;
            TileCfg g{};
            g.TH = TH; g.CWB = CWB; g.RH = RH; g.PS = PS; g.R = R;
            uint32_t off = 0;
            g.off_col = off; off = align_up(off + (staged ? 0u : (uint32_t)CWB * 16u), 16);
            g.off_P = off; off = align_up(off + (uint32_t)(RH + (staged ? 2 : 0)) * PS * 2u, 16);
            g.off_list = off; off = align_up(off + (staged ? 0u : (uint32_t)(R * (CWB + 1) + 2) * 2u), 16);
            g.off_E = off; off = align_up(off + (staged ? 0u : (uint32_t)(R * (CWB + 1) + 2) * 2u), 16);
            g.off_cnt = off; off = align_up(off + (staged ? 0u : (uint32_t)(FG_TILE_NE + 8) * 4u), 16);
            g.off_rows = off; off = align_up(off + (staged ? (uint32_t)RH * 28u + 16u : 0u), 16);
            g.off_wtot = off; off = align_up(off + (uint32_t)(FG_TILE_WARPS + 4) * 4u, 16);
            g.off_pcount = off; off = align_up(off + (uint32_t)TH * 32u * 4u, 16);
            g.off_wpair = off; off = align_up(off + (uint32_t)FG_TILE_WARPS * TH * spwc * 8u, 16);
            if ((size_t)off + 64 + (size_t)bpg * (2048 + FG_TILE_GPAD) > smem_max) continue;
            uint32_t gcap = (uint32_t)((smem_max - off - 64) / bpg) - FG_TILE_GPAD;
            gcap = std::min<uint32_t>(gcap / 64 * 64, 65472u);
            if (gcap < 2048) continue;
            if (pass == 0 && (double)gcap < 0.45 * (double)RH * (double)CWB) continue;
            g.GCAP = (int)gcap;
            g.off_G = off; off = align_up(off + (gcap + FG_TILE_GPAD) * 8u, 16);
            g.off_R2 = off;
            if (c.rad.lognorm) off = align_up(off + (gcap + FG_TILE_GPAD) * 4u, 16);
            g.total = off;
            if (off > smem_max) continue;
            pl.cfg = g;
            pl.ok = true;
        }
    }
    if (!pl.ok) return pl;
    TileCfg& g = pl.cfg;
    g.n_strips = (int)((p->out_w + 31) / 32);
    // segment height: maximise (wave efficiency over the SMs) x (1 - start-up share).  A segment
    // regenerates the cell rows of its first window (RH rows = RH*delta*zoom pixel rows of work).
    const double startup_rows = (double)g.RH * delta * (double)p->zoom * (staged ? 0.05 : 1.0); // staged: a load, not a generation
    const long long per_seg_units = (long long)g.n_strips * n_planes;
    int best_n = 1;
    double best_eff = -1.0;
    const int max_n = std::max(1, band / std::max(1, 4 * g.TH));
    auto seg_eff = [&](int n) {
        int seg = (band + n - 1) / n;
        seg = (seg + g.TH - 1) / g.TH * g.TH;
        const int n_eff = (band + seg - 1) / seg;
        const double waves = (double)per_seg_units * n_eff / ctx->sm_count;
        return waves / std::ceil(waves) * ((double)seg / ((double)seg + startup_rows));
    };
    for (int n = 1; n <= max_n; ++n) best_eff = std::max(best_eff, seg_eff(n));
    // among near-optimal splits prefer the finest one: more, shorter CTAs balance content-dependent cost
    for (int n = 1; n <= max_n; ++n)
        if (seg_eff(n) >= best_eff - 0.01) best_n = n;
    int seg = (band + best_n - 1) / best_n;
    seg = (seg + g.TH - 1) / g.TH * g.TH;
    g.SEG = seg;
    g.n_segs = (band + seg - 1) / seg;
    pl.spwc = spwc;
    return pl;
}

struct SkewPlan { SkewCfg cfg; int spwc; bool ok; };

// Geometry of k_pixelwise_skew for a render; ok == false -> use k_pixelwise_strip.  `dens` = grains per cell of the
// band's table (its capacity, i.e. expectation + slack): the merged window of a step must fit shared memory with room
// for local fluctuations (denser strips go through the fallback list at run time).
SkewPlan skew_plan(const fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int n_planes, double dens) {
    SkewPlan pl{};
    pl.ok = false;
    // Measured on a B200 (profiles/README.md): correct, but at C2 its loader (three copies of every grain, per-step set-up
    // over only 8 items per warp) costs more than the merged lists save -- 45.7 ms against 36 ms for k_pixelwise_strip.
    // Kept behind FG_B200_SKEW=1 for experiments and as a third, independent evaluation kernel in the parity tests.
    static const bool enabled = std::getenv("FG_B200_SKEW") && std::atoi(std::getenv("FG_B200_SKEW")) != 0;
    if (!enabled || c.rad.lognorm || p->rm != p->delta || p->n_samples > (1u << 20)) return pl;
    const double inv_zoom = 1.0 / (double)p->zoom, delta = p->delta, rm = p->rm;
    const double ox = (double)c.off_max_x - (double)c.off_min_x, oy = (double)c.off_max_y - (double)c.off_min_y;
    const double cwb = (31.0 * inv_zoom + ox + 2.0 * rm) / delta + 4.0 + 3.0; // + alignment shift
    if (!(cwb < 2000.0)) return pl;
    const int CWB = (int)cwb;
    const int PS = (CWB + 1 + 7) / 8 * 8;
    const double cpr = inv_zoom / delta; // cell rows per output pixel row
    if (!(cpr > 0.05 && cpr < 64.0)) return pl;
    const int spwc = p->n_samples <= 4u * FG_SK_WARPS ? 4 : 8;
    const int cand = 32 / spwc;
    const int n_chunks = (int)((p->n_samples + FG_SK_WARPS * spwc - 1) / (FG_SK_WARPS * spwc));
    const int band = c.row_end - c.row_begin;
    const double span_rows = oy * (double)p->zoom + 2.0; // pixel rows between the fastest and the slowest sample cursor
    const size_t smem_max = ctx->smem_optin;
    static const int th_forced = std::getenv("FG_B200_SKEW_TH") ? std::atoi(std::getenv("FG_B200_SKEW_TH")) : 0; // experiments
    static const int th_candidates[] = {8, 6, 4, 3, 2, 1};
    for (int TH : th_candidates) {
        if (th_forced > 0 && TH != th_forced) continue;
        if (TH + 1 > cand && TH > 1) continue; // a sample's rows of one step fit one item round
        const double dr = TH * cpr;
        int D = std::fabs(dr - std::nearbyint(dr)) < 1e-3 ? (int)std::nearbyint(dr) : (int)std::ceil(dr);
        D = std::max(1, std::min(D, FG_SK_DMAX));
        const int S = D + 2;
        const double rows_per_step = D / cpr;
        int R = 4;
        while (R < (int)std::ceil(span_rows + 2.0 * (rows_per_step + 1.0) + 3.0)) R <<= 1;
        if (R > 2048) continue;
        SkewCfg g{};
        g.D = D; g.S = S; g.CWB = CWB; g.PS = PS; g.R = R;
        uint32_t off = 0;
        g.off_Ps = off; off = align_up(off + 2u * (uint32_t)S * PS * 2u, 16);
        g.off_Q = off; off = align_up(off + 2u * (uint32_t)D * PS * 2u, 16);
        g.off_zero = off; off = align_up(off + (uint32_t)PS * 2u, 16);
        g.off_info = off; off = align_up(off + 2u * 32u * 16u, 16);
        g.off_tb = off; off = align_up(off + 2u * 32u * 4u, 16);
        g.off_wp = off; off = align_up(off + (uint32_t)FG_SK_WARPS * 512u, 16);
        g.off_pcw = off; off = align_up(off + (uint32_t)FG_SK_WARPS * (uint32_t)(R + 1) * 32u, 16);
        g.off_sync = off; off = align_up(off + 16u, 16);
        // segment height (also sizes the chunk accumulator): wave efficiency x (1 - ramp share)
        g.n_strips = (int)((p->out_w + 31) / 32);
        const long long per_seg_units = (long long)g.n_strips * n_planes;
        const double startup_rows = 0.35 * span_rows + rows_per_step;
        const int seg_cap = n_chunks > 1 ? 64 : 4096;
        int best_n = 1;
        double best_eff = -1.0;
        const int max_n = std::max(1, band / std::max(1, (int)std::ceil(2.0 * rows_per_step)));
        auto seg_of = [&](int n) { return std::min((band + n - 1) / n, seg_cap); };
        auto seg_eff = [&](int n) {
            const int seg = seg_of(n);
            const int n_eff = (band + seg - 1) / seg;
            const double waves = (double)per_seg_units * n_eff / ctx->sm_count;
            return waves / std::ceil(waves) * ((double)seg / ((double)seg + startup_rows));
        };
        for (int n = 1; n <= max_n; ++n) best_eff = std::max(best_eff, seg_eff(n));
        for (int n = 1; n <= max_n; ++n)
            if (seg_eff(n) >= best_eff - 0.01) best_n = n; // among near-optimal splits the finest: balances content-dependent cost
        g.SEG = seg_of(best_n);
        static const int seg_forced = std::getenv("FG_B200_SKEW_SEG") ? std::atoi(std::getenv("FG_B200_SKEW_SEG")) : 0; // experiments
        if (seg_forced > 0) g.SEG = std::min(seg_forced, seg_cap);
        g.n_segs = (band + g.SEG - 1) / g.SEG;
        g.off_pcc = off; off = align_up(off + (n_chunks > 1 ? (uint32_t)g.SEG * 128u : 0u), 16);
        if ((size_t)off + 2u * 8u * 1024u > smem_max) continue;
        uint32_t mcap = (uint32_t)((smem_max - off) / 16u); // two buffers of 8-byte grains
        mcap = std::min<uint32_t>(mcap, 65504u);
        // every triple owns a fixed range: expectation + room for the fluctuation of a 3-row window (content is
        // correlated over whole input pixels, so the relative spread does not shrink with the cell size)
        const uint32_t tcap = mcap / (uint32_t)D / 4u * 4u;
        const double expect = 3.0 * (double)CWB * dens;
        if ((double)tcap < 1.8 * expect + 64.0) continue;
        mcap = tcap * (uint32_t)D;
        g.MCAP = (int)mcap;
        g.TCAP = (int)tcap;
        g.off_M = off; off += 2u * mcap * 8u;
        g.total = off;
        if (off > smem_max) continue;
        pl.cfg = g;
        pl.spwc = spwc;
        pl.ok = true;
        break;
    }
    return pl;
}


struct TriPlan { TriCfg cfg; int spw; bool ok; };

// Geometry of k_pixelwise_tri for a render; ok == false -> another evaluation kernel.  `dens` = grains per cell of the
// band's table (its capacity, i.e. expectation + slack): the merged ring must hold the window of a step plus the rows
// being merged with room for local fluctuations (denser strips go through the fallback list at run time).
TriPlan tri_plan(const fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int n_planes, double dens) {
    TriPlan pl{};
    pl.ok = false;
    static const bool enabled = !(std::getenv("FG_B200_TRI") && std::atoi(std::getenv("FG_B200_TRI")) == 0);
    // Measured on a B200 (profiles/README.md).  128 < N <= 256: 28 ms against 36 ms for k_pixelwise_strip at C2.  N <= 128: the
    // kernel's time hardly depends on N (its loader warps set the step period: ~26 ms for a 4K RGB frame at 10 cell rows per
    // output row) while the strip kernel's does, so it only pays where a step needs few cell rows: cpr = 1 / (zoom * delta)
    // <= 5 (zoom 2 at r = 0.1, zoom 4 at r = 0.05 = C4: 74 ms against 88 ms; zoom 4 at r = 0.1: 5.5 against 10.8 ms) and
    // N > 32 (a quarter-filled warp of samples is the strip kernel's territory).  FG_B200_TRI_SMALLN = 0 / 1 forces it.
    if (!enabled || c.rad.lognorm || p->rm != p->delta || p->n_samples > (uint32_t)FG_TRI_SPW_C * FG_TRI_WARPS) return pl;
    const double inv_zoom = 1.0 / (double)p->zoom, delta = p->delta, rm = p->rm;
    const double cpr = inv_zoom / delta; // cell rows per output pixel row
    if (p->n_samples <= (uint32_t)FG_TRI_SPW_B * FG_TRI_WARPS) {
        const char* sn = std::getenv("FG_B200_TRI_SMALLN"); // read per call: tests and tools switch it
        const bool small_n = sn ? std::atoi(sn) != 0 : (p->n_samples > 32u && cpr <= 5.01);
        if (!small_n) return pl;
    }
    const double ox = (double)c.off_max_x - (double)c.off_min_x;
    const double cwb = (31.0 * inv_zoom + ox + 2.0 * rm) / delta + 4.0 + 3.0; // + alignment shift
    if (!(cwb < 2000.0)) return pl;
    const int CWB = (int)cwb;
    const int PS = (CWB + 1 + 7) / 8 * 8;
    if (!(cpr > 0.05 && cpr < 24.0)) return pl;
    const int spw = p->n_samples <= (uint32_t)FG_TRI_SPW_A * FG_TRI_WARPS ? FG_TRI_SPW_A : (p->n_samples <= (uint32_t)FG_TRI_SPW_B * FG_TRI_WARPS ? FG_TRI_SPW_B : FG_TRI_SPW_C);
    const int band = c.row_end - c.row_begin;
    const int skrange = (int)std::nearbyint((double)c.off_max_y * (double)p->zoom) - (int)std::nearbyint((double)c.off_min_y * (double)p->zoom);
    if (skrange < 0 || skrange > 200) return pl;
    const size_t smem_max = ctx->smem_optin;
    static const int m_forced = std::getenv("FG_B200_TRI_M") ? std::atoi(std::getenv("FG_B200_TRI_M")) : 0; // experiments
    // Step height m: the ring of merged rows and the staging buffer share what shared memory is left, in proportion to their
    // expected sizes.  `headroom` = capacity / expectation; local density fluctuates a lot (a window only sees a few
    // input pixels of a zoomed image, and -ln(1 - u) is heavy-tailed), so the largest m with a headroom of 2.5 is taken,
    // else the m with the most headroom; below 1.6 the strip kernel is the better choice.
    static const int m_candidates[] = {16, 12, 8, 6, 4, 3, 2, 1};
    double best_h = 0.0;
    TriCfg best{};
    for (int m : m_candidates) {
        if (m_forced > 0 && m != m_forced) continue;
        if (m * spw > 32) continue;
        const int A = (int)std::ceil(m * cpr - 1e-9);
        if (A < 1 && m < 16) continue; // at least one cell row per step
        TriCfg g{};
        g.m = m;
        g.AMAX = A + 1;
        g.D = A + 4;
        g.NG = g.AMAX + 2;
        if (g.NG > 32) continue;
        g.NQ = g.D + g.AMAX;
        g.CWB = CWB;
        g.PS = PS;
        int RH = 4;
        while (RH < 3 * m + skrange + 2) RH <<= 1; // the evaluation warps may be a step ahead of the rows being finalised
        g.RH = RH;
        g.NB = std::max(1, 32 / (m * spw));
        uint32_t off = 0;
        g.off_zq = off; off = align_up(off + (uint32_t)PS * 2u, 128);
        g.off_zp = off; off = align_up(off + (uint32_t)PS * 4u, 128);
        g.off_praw = off; off = align_up(off + (uint32_t)g.NG * PS * 4u, 128);
        g.off_sinfo = off; off = align_up(off + (uint32_t)g.NG * 32u, 128);
        g.off_ext = off; off = align_up(off + 2u * (uint32_t)g.NG * 16u, 128);
        g.off_minfo = off; off = align_up(off + (uint32_t)g.NQ * 4u, 128);
        g.off_state = off; off = align_up(off + 256u, 128);
        g.off_items = off; off = align_up(off + (uint32_t)FG_TRI_WARPS * 512u, 128);
        g.off_hb = off; off = align_up(off + (uint32_t)(RH + 1) * align_up((uint32_t)FG_TRI_WARPS * spw * 4u, 128), 128);
        g.off_Q = off; off = align_up(off + (uint32_t)g.NQ * PS * 2u, 128);
        if ((size_t)off + 16u * 1024u > smem_max) continue;
        const double row_grains = (double)CWB * dens + 1.0; // of one source row's window
        // live rows: the window, the rows being merged, and the row a wrap leaves unused at the end of the ring
        const double need_m = 3.0 * row_grains * (double)(g.NQ + 1) * 8.0;
        const double need_gs = (double)g.NG * (row_grains + 16.0) * 10.0;
        const double left = (double)(smem_max - off) - 1024.0;
        double h = left / (need_m + need_gs);
        uint32_t gs = align_up((uint32_t)std::min(h * need_gs, 100.0e3) + 256u, 128);
        uint32_t mcap = (uint32_t)std::min<double>(((double)(smem_max - off) - (double)gs) / 8.0, 65000.0);
        h = std::min(h, (double)mcap * 8.0 / need_m);
        if (std::getenv("FG_B200_DEBUG"))
            std::fprintf(stderr, "[fg] tri_plan m=%d spw=%d A=%d NQ=%d PS=%d GS=%u RH=%d fixed=%u mcap=%u headroom=%.2f dens=%.3f\n", m, spw, A, g.NQ, PS, gs, RH, off, mcap, h, dens);
        g.GS = (int)gs;
        g.off_gs = off; off += gs;
        g.MCAP = (int)mcap;
        g.off_M = off; off += mcap * 8u;
        g.total = off;
        if (off > smem_max) continue;
        if (h > best_h) { best_h = h; best = g; }
        if (h >= 2.5) break;
    }
    const char* mh = std::getenv("FG_B200_TRI_MIN_HEADROOM"); // tests: force the kernel onto dense content (skipped groups)
    if (best_h < (mh ? std::atof(mh) : 1.6)) return pl;
    {
        TriCfg g = best;
        const int m = g.m;
        // segment height: wave efficiency x (1 - ramp share)
        g.n_strips = (int)((p->out_w + 31) / 32);
        const long long per_seg_units = (long long)g.n_strips * n_planes;
        const double startup_rows = 0.5 * skrange + 3.0 * m;
        int best_n = 1;
        double best_eff = -1.0;
        const int max_n = std::max(1, band / std::max(8, 4 * m));
        auto seg_of = [&](int n) { return (band + n - 1) / n; };
        auto seg_eff = [&](int n) {
            const int seg = seg_of(n);
            const int n_eff = (band + seg - 1) / seg;
            const double waves = (double)per_seg_units * n_eff / ctx->sm_count;
            return waves / std::ceil(waves) * ((double)seg / ((double)seg + startup_rows));
        };
        for (int n = 1; n <= max_n; ++n) best_eff = std::max(best_eff, seg_eff(n));
        for (int n = 1; n <= max_n; ++n)
            if (seg_eff(n) >= best_eff - 0.005) { best_n = n; break; } // among near-optimal splits the coarsest: fewer ramps
        if (cancel_armed(ctx)) // a cancel flag is tested when a CTA starts: many short segments bound the latency of a cancel
            for (int n = best_n; n <= max_n; ++n)
                if (seg_eff(n) >= best_eff - 0.03) best_n = n;
        g.SEG = seg_of(best_n);
        static const int seg_forced = std::getenv("FG_B200_TRI_SEG") ? std::atoi(std::getenv("FG_B200_TRI_SEG")) : 0; // experiments
        if (seg_forced > 0) g.SEG = seg_forced;
        g.n_segs = (band + g.SEG - 1) / g.SEG;
        g.inv_delta = (float)(1.0 / delta);
        pl.cfg = g;
        pl.spw = spw;
        pl.ok = true;
    }
    return pl;
}

template <int SP, bool LG, bool ST>
cudaError_t strip_attr(int smem) {
    return cudaFuncSetAttribute(k_pixelwise_strip<SP, LG, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

int tile_setup(fg_ctx* ctx) {
    cudaError_t e;
    const int smem = (int)ctx->smem_optin;
#define FG_ATTR(SP)                                                                                               \
    if ((e = strip_attr<SP, false, false>(smem)) != cudaSuccess || (e = strip_attr<SP, true, false>(smem)) != cudaSuccess || \
        (e = strip_attr<SP, false, true>(smem)) != cudaSuccess || (e = strip_attr<SP, true, true>(smem)) != cudaSuccess)     \
        return map_cuda_error(ctx, e, "cudaFuncSetAttribute(k_pixelwise_strip)");
    FG_ATTR(4)
    FG_ATTR(8)
    FG_ATTR(FG_TILE_SPW_MAX)
#undef FG_ATTR
    if ((e = cudaFuncSetAttribute(k_pixelwise_skew<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pixelwise_skew<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
        return map_cuda_error(ctx, e, "cudaFuncSetAttribute(k_pixelwise_skew)");
    if ((e = cudaFuncSetAttribute(k_pixelwise_tri<FG_TRI_SPW_C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pixelwise_tri<FG_TRI_SPW_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pixelwise_tri<FG_TRI_SPW_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess)
        return map_cuda_error(ctx, e, "cudaFuncSetAttribute(k_pixelwise_tri)");
    return FG_OK;
}

template <bool LG, bool ST, typename... Args>
void launch_strip(int spwc, uint32_t units, uint32_t smem, cudaStream_t s, Args... args) {
    switch (spwc) {
    case 4: k_pixelwise_strip<4, LG, ST><<<units, FG_TILE_THREADS, smem, s>>>(args...); break;
    case 8: k_pixelwise_strip<8, LG, ST><<<units, FG_TILE_THREADS, smem, s>>>(args...); break;
    default: k_pixelwise_strip<FG_TILE_SPW_MAX, LG, ST><<<units, FG_TILE_THREADS, smem, s>>>(args...); break;
    }
}

// Upper bound on the cell table (prefixes + grains) of one band; larger renders are split into row bands.
#ifndef FG_TABLE_BYTES_MAX
#define FG_TABLE_BYTES_MAX ((size_t)48 << 30)
#endif

// One band [c.row_begin, c.row_end).  returns FG_OK (rendered), 1 (tiled path not applicable), 2 (the
// cell table of this band does not fit the memory budget: caller splits the band), 3 (the table
// overflowed: caller uses in-kernel generation) or an error.
int tile_render_band(fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int n_planes, const float* d_lambda,
                     const float* d_offsets, float* d_out, bool staged, uint32_t* d_fbtotal, bool skew_ok) {
    TilePlan pl = tile_plan(ctx, p, c, n_planes, staged);
    if (!pl.ok) return 1;
    const size_t in_stride = (size_t)p->in_w * p->in_h, out_stride = (size_t)p->out_w * p->out_h;
    const size_t n_in = in_stride * n_planes;
    // cell rectangle of the band: every cell any strip window can touch (+2 cells of slack against
    // f32-vs-f64 rounding; the kernel re-checks its windows against these bounds)
    TileCfg g = pl.cfg;
    {
        const double iz = 1.0 / (double)p->zoom, dl = p->delta, rmd = p->rm;
        const double i0 = std::floor((0.5 * iz - (double)c.off_max_x - rmd) / dl) - 2.0;
        const double i1 = std::floor((((double)p->out_w - 0.5) * iz - (double)c.off_min_x + rmd) / dl) + 2.0;
        const double j0 = std::floor((((double)c.row_begin + 0.5) * iz - (double)c.off_max_y - rmd) / dl) - 2.0;
        const double j1 = std::floor((((double)c.row_end - 0.5) * iz - (double)c.off_min_y + rmd) / dl) + 2.0;
        if (!(i0 > -2.0e9 && i1 < 2.0e9 && j0 > -2.0e9 && j1 < 2.0e9)) return 1;
        g.bm_i0 = (int)i0; g.bm_j0 = (int)j0;
        g.bm_cols = (int)(i1 - i0 + 1.0); g.bm_rows = (int)(j1 - j0 + 1.0);
        g.bm_pitchw = (uint32_t)((g.bm_cols + 31) / 32);
        g.ppitch = (uint32_t)((g.bm_cols + 1 + 7) / 8 * 8);
        const float rcl = c.rad.mean_linear > c.rad.rm ? c.rad.rm : c.rad.mean_linear;
        g.r2c = rcl * rcl;
    }
    // ---- table cache (fg_set_table_cache; the viewer's re-render after an N / sigma / zoom change, src/bin/viewer.rs:944-1067):
    // a whole-frame staged render whose table identity equals the cached one and whose cell rectangle lies inside the
    // cached rectangle evaluates from the table in the pools; otherwise the table is built over a rectangle with a margin
    // (so that a somewhat larger sigma or a smaller zoom still fits) and remembered.
    TableCache& tc = ctx->tcache;
    const bool use_cache = staged && tc.enabled && c.row_begin == 0 && c.row_end == (int)p->out_h;
    bool reuse = false;
    TableKey key{};
    if (use_cache) {
        int rc0;
        if ((rc0 = ensure(ctx, ctx->misc, 64))) return rc0;
        unsigned long long* d_h = (unsigned long long*)ctx->misc.p + 2;
        FG_CUDA(ctx, cudaMemsetAsync(d_h, 0, 16, ctx->stream));
        k_hash_planes<<<(unsigned)std::min<size_t>((n_in + 255) / 256, (size_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(d_lambda, n_in, d_h);
        FG_CUDA(ctx, cudaGetLastError());
        FG_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin + 2, d_h, 16, cudaMemcpyDeviceToHost, ctx->stream));
        FG_CUDA(ctx, wait_stream(ctx));
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
        const uint64_t hh[2] = {ctx->h_pin[2], ctx->h_pin[3]};
        ctx->stats.launches += 1;
        key.seed_cell = c.seed_cell; key.h0 = hh[0]; key.h1 = hh[1];
        key.seeding = c.seeding; key.in_w = p->in_w; key.in_h = p->in_h; key.n_planes = (uint32_t)n_planes; key.lognorm = c.rad.lognorm;
        key.delta = p->delta; key.rm = c.rad.rm; key.mean_linear = c.rad.mean_linear; key.mu = c.rad.mu; key.sigma = c.rad.sigma;
        key.slack = ctx->table_slack_sigma;
        reuse = tc.valid && key == tc.key && tc.i0 <= g.bm_i0 && tc.j0 <= g.bm_j0 &&
                (long long)tc.i0 + tc.cols >= (long long)g.bm_i0 + g.bm_cols && (long long)tc.j0 + tc.rows >= (long long)g.bm_j0 + g.bm_rows;
        if (reuse) {
            g.bm_i0 = tc.i0; g.bm_j0 = tc.j0; g.bm_cols = tc.cols; g.bm_rows = tc.rows;
        } else { // margin: one input pixel + half the largest offset, in cells, on every side
            const double moff = std::max(std::max(std::fabs((double)c.off_min_x), std::fabs((double)c.off_max_x)),
                                         std::max(std::fabs((double)c.off_min_y), std::fabs((double)c.off_max_y)));
            const int extra = (int)std::min(4096.0, std::ceil((1.0 + 0.5 * moff) / (double)p->delta));
            g.bm_i0 -= extra; g.bm_j0 -= extra; g.bm_cols += 2 * extra; g.bm_rows += 2 * extra;
        }
        g.bm_pitchw = (uint32_t)((g.bm_cols + 31) / 32);
        g.ppitch = (uint32_t)((g.bm_cols + 1 + 7) / 8 * 8);
    }
    if (!reuse) tc.valid = false; // the pools are about to be rewritten (or are not a whole-frame staged table)
    ctx->stats.table_reused = reuse ? 1u : 0u;
    const size_t bm_plane_words = (size_t)g.bm_rows * g.bm_pitchw;
    if (bm_plane_words * (size_t)n_planes * 4 > ((size_t)12 << 30) || g.bm_pitchw * 32u / 256u + 1u > 65535u) return 1;
    const size_t n_rows_all = (size_t)g.bm_rows * n_planes;
    const size_t pg_bytes = n_rows_all * g.ppitch * 4 + 256; // + slack: bulk copies of the last row read up to 3 entries past its window
    if (staged && (pg_bytes > ctx->table_max / 2 || g.bm_rows > 2000000)) return 2;
    int rc;
    if ((rc = ensure(ctx, ctx->thr, n_in * 16))) return rc;
    if ((rc = ensure(ctx, ctx->bitmap, bm_plane_words * (size_t)n_planes * 4))) return rc;
    const uint32_t units = (uint32_t)g.n_strips * g.n_segs * n_planes;
    if ((rc = ensure(ctx, ctx->tiles, (size_t)units * sizeof(TileRef) + 64))) return rc;
    uint64_t* d_thr = (uint64_t*)ctx->thr.p;
    double* d_e = (double*)((unsigned char*)ctx->thr.p + n_in * 8);
    uint32_t* d_fbcount = (uint32_t*)ctx->tiles.p;
    TileRef* d_fblist = (TileRef*)((unsigned char*)ctx->tiles.p + 64);
    cudaStream_t s = ctx->stream;
    FG_CUDA(ctx, cudaMemsetAsync(d_fbcount, 0, 64, s));
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[6], s)); // ev[6]..ev[4]: thresholds, bitmap, cell table
    // input rows the band's cell rows map to (clamped like Plane::get_clamped), one row of slack
    const int iy0 = std::min(std::max((int)std::floor((double)g.bm_j0 * (double)p->delta) - 1, 0), (int)p->in_h - 1);
    const int iy1 = std::min(std::max((int)std::floor((double)(g.bm_j0 + g.bm_rows) * (double)p->delta) + 1, 0), (int)p->in_h - 1);
    const size_t thr_first = (size_t)iy0 * p->in_w, thr_n = (size_t)(iy1 - iy0 + 1) * p->in_w;
    uint32_t* d_bm = (uint32_t*)ctx->bitmap.p;
    // thresholds of the input rows [r0, r1) and first-draw bits of the cell rows [c0, c1) of the rectangle
    auto launch_front = [&](int r0, int r1, int c0, int c1) -> int {
        if (r1 > r0) {
            const size_t first = (size_t)r0 * p->in_w, n = (size_t)(r1 - r0) * p->in_w;
            const unsigned tb = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
            k_thresholds<<<dim3(tb, (unsigned)n_planes), 256, 0, s>>>(d_lambda, in_stride, first, n, p->delta, d_thr, d_e);
            FG_CUDA(ctx, cudaGetLastError());
            ctx->stats.launches += 1;
        }
        if (c1 > c0) {
            dim3 bgrid((unsigned)((c1 - c0 + FG_BM_ROWS - 1) / FG_BM_ROWS), (g.bm_pitchw * 32u + 255u) / 256u);
#define FG_LAUNCH_BM(SD, NPL)                                                                                       \
    k_first_draw_bitmap<SD, NPL><<<bgrid, 256, 0, s>>>(d_thr, in_stride, n_planes, d_bm, bm_plane_words, g.bm_i0, g.bm_j0, \
                                                       g.bm_cols, c1, g.bm_pitchw, c0, c)
            if (c.seeding == 0) {
                if (n_planes == 3) FG_LAUNCH_BM(0, 3);
                else if (n_planes == 1) FG_LAUNCH_BM(0, 1);
                else FG_LAUNCH_BM(0, 0);
            } else {
                FG_LAUNCH_BM(1, 0);
            }
#undef FG_LAUNCH_BM
            FG_CUDA(ctx, cudaGetLastError());
            ctx->stats.launches += 1;
        }
        return FG_OK;
    };
    if (!reuse && ctx->up.pending && staged && iy0 == 0 && iy1 == (int)p->in_h - 1 && ctx->up_ev[4]) {
        // ---- chunked upload: lambda rows cross PCIe (or the host's staging buffer) on `copy_stream` in K row chunks; the
        // thresholds and the first-draw bits of the cell rows that read chunk k run while chunk k + 1 is on its way.
        // A cell row j reads the input row clamp(floor(f32(j) * delta)) (src/pixelwise.rs:69-73), non-decreasing in j.
        ctx->up.pending = false;
        const int K = 4, in_h = (int)p->in_h;
        auto in_row = [&](int row) { // the kernels' own arithmetic: one f32 multiply, floor, clamp
            const float v = std::floor((float)(g.bm_j0 + row) * p->delta);
            const int iy = v < -2.0e9f ? INT_MIN : (v > 2.0e9f ? INT_MAX : (int)v);
            return std::min(std::max(iy, 0), in_h - 1);
        };
        FG_CUDA(ctx, cudaEventRecord(ctx->up_ev[4], s)); // the copies come after whatever the main stream did to the buffer (tests: NaN fill)
        FG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->up_ev[4], 0));
        int c_done = 0;
        for (int k = 0; k < K; ++k) {
            const int r0 = (int)((long long)in_h * k / K), r1 = (int)((long long)in_h * (k + 1) / K);
            const size_t off = (size_t)r0 * p->in_w, n = (size_t)(r1 - r0) * p->in_w;
            for (int pl = 0; pl < n_planes && n; ++pl)
                FG_CUDA(ctx, cudaMemcpyAsync(ctx->up.dev + in_stride * pl + off, ctx->up.host[pl] + off, n * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_stream));
            FG_CUDA(ctx, cudaEventRecord(ctx->up_ev[k], ctx->copy_stream));
            FG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->up_ev[k], 0));
            int c1 = g.bm_rows;
            if (k + 1 < K) { // leading cell rows whose input row lies below r1
                int lo = c_done, hi = g.bm_rows; // in_row(row) < r1 for row < lo; first row with in_row >= r1 is in [lo, hi]
                while (lo < hi) {
                    const int mid = lo + (hi - lo) / 2;
                    if (in_row(mid) < r1) lo = mid + 1; else hi = mid;
                }
                c1 = lo;
            }
            int rcf;
            if ((rcf = launch_front(r0, r1, c_done, c1))) return rcf;
            c_done = c1;
        }
    } else {
        FG_CUDA(ctx, flush_upload(ctx));
        int rcf;
        if (!reuse && (rcf = launch_front(iy0, iy1 + 1, 0, g.bm_rows))) return rcf;
    }
    CellTable tab{};
    double table_dens = 0.0;
    if (staged) {
        // row capacities from the expected grain counts, then the table itself
        const size_t s_bytes = align_up((uint32_t)((size_t)n_planes * p->in_h * 8), 256);
        const size_t base_bytes = (n_rows_all + 1) * 8, cap_bytes = n_rows_all * 4;
        if ((rc = ensure(ctx, ctx->rowinfo, s_bytes + base_bytes + cap_bytes + 256 + 64))) return rc;
        if ((rc = ensure(ctx, ctx->ptab, pg_bytes))) {
            if (rc == FG_ERR_OOM) { ctx->err.clear(); return 2; }
            return rc;
        }
        double* d_S = (double*)ctx->rowinfo.p;
        uint64_t* d_rowbase = (uint64_t*)((unsigned char*)ctx->rowinfo.p + s_bytes);
        uint32_t* d_rowcap = (uint32_t*)((unsigned char*)d_rowbase + (base_bytes + 255) / 256 * 256);
        uint32_t* d_overflow = (uint32_t*)((unsigned char*)d_rowcap + (cap_bytes + 63) / 64 * 64);
        uint64_t total = tc.total;
        if (!reuse) {
            k_row_expect<<<dim3((unsigned)(iy1 - iy0 + 1), n_planes), 256, 0, s>>>(d_lambda, in_stride, g.bm_i0, g.bm_cols, iy0, d_S, c);
            FG_CUDA(ctx, cudaGetLastError());
            k_row_bases<<<1, 1024, 0, s>>>(d_S, g.bm_j0, g.bm_rows, n_planes, ctx->table_slack_sigma, d_rowbase, d_rowcap, c);
            FG_CUDA(ctx, cudaGetLastError());
            FG_CUDA(ctx, cudaMemsetAsync(d_overflow, 0, 4, s));
            FG_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin, d_rowbase + n_rows_all, 8, cudaMemcpyDeviceToHost, s));
            FG_CUDA(ctx, wait_stream(ctx));
            total = ctx->h_pin[0];
            if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
        }
        const size_t bpg = c.rad.lognorm ? 14 : 10;
        if (total == 0xFFFFFFFFFFFFFFFFULL || total * bpg + pg_bytes > ctx->table_max) return 2;
        const size_t g_bytes = ((size_t)total * 8 + 255) / 256 * 256;
        const size_t r2_bytes = c.rad.lognorm ? ((size_t)total * 4 + 255) / 256 * 256 : 0;
        if ((rc = ensure(ctx, ctx->gtab, g_bytes + r2_bytes + (size_t)total * 2 + 256))) {
            if (rc == FG_ERR_OOM) { ctx->err.clear(); return 2; }
            return rc;
        }
        float2* d_G = (float2*)ctx->gtab.p;
        float* d_R2 = (float*)((unsigned char*)ctx->gtab.p + g_bytes);
        uint16_t* d_C = (uint16_t*)((unsigned char*)ctx->gtab.p + g_bytes + r2_bytes);
        StageGeo geo{g.bm_i0, g.bm_j0, g.bm_cols, g.bm_rows, g.bm_pitchw, g.ppitch};
        if (!reuse) {
        // three planes: one warp generates a cell row for all of them (one seeding and one Knuth chain per cell)
        const bool joint = n_planes == 3 && !(std::getenv("FG_B200_GEN_JOINT") && std::atoi(std::getenv("FG_B200_GEN_JOINT")) == 0);
        const size_t gwarps = joint ? (size_t)g.bm_rows : n_rows_all;
        const unsigned ggrid = (unsigned)((gwarps + FG_GW_WARPS - 1) / FG_GW_WARPS);
#define FG_LAUNCH_GEN(LG, NPL)                                                                                                   \
    k_gen_rows<LG, NPL><<<ggrid, FG_GW_WARPS * 32, 0, s>>>(d_bm, bm_plane_words, d_e, d_lambda, in_stride, (uint32_t*)ctx->ptab.p, \
                                                           d_rowbase, d_rowcap, d_G, d_R2, d_C, d_overflow, geo, n_planes, c)
        if (c.rad.lognorm) { if (joint) FG_LAUNCH_GEN(true, 3); else FG_LAUNCH_GEN(true, 1); }
        else { if (joint) FG_LAUNCH_GEN(false, 3); else FG_LAUNCH_GEN(false, 1); }
#undef FG_LAUNCH_GEN
        FG_CUDA(ctx, cudaGetLastError());
        FG_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin + 1, d_overflow, 4, cudaMemcpyDeviceToHost, s));
        FG_CUDA(ctx, wait_stream(ctx));
        const uint32_t overflow = *(const uint32_t*)(ctx->h_pin + 1);
        if (cancelled(ctx)) return set_err(ctx, FG_ERR_CANCELLED, "cancelled");
        ctx->stats.launches += 3;
        if (overflow) return 3; // a row outgrew its expected size + 8 sigma (or a cell holds > 65535 grains): regenerate in-kernel instead
        }
        tab.Pg = (const uint32_t*)ctx->ptab.p;
        tab.rowbase = d_rowbase;
        tab.Gg = d_G;
        tab.R2g = d_R2;
        tab.Cg = d_C;
        {   // grains per cell EXPECTED in the band: the average row capacity is expectation + slack_sigma * sqrt(expectation) + 64
            const double cap_row = (double)total / (double)n_rows_all, k = ctx->table_slack_sigma;
            const double rt = (-k + std::sqrt(k * k + 4.0 * std::max(cap_row - 64.0, 0.0))) * 0.5;
            table_dens = rt * rt / (double)g.bm_cols;
        }
        if (use_cache) { // the pools now hold this table
            tc.valid = true; tc.key = key;
            tc.i0 = g.bm_i0; tc.j0 = g.bm_j0; tc.cols = g.bm_cols; tc.rows = g.bm_rows;
            tc.total = total; tc.dens = table_dens;
        }
    }
    const float2* off = (const float2*)d_offsets;
    FG_CUDA(ctx, cudaEventRecord(ctx->ev[4], s));
    // ---- evaluation of the rows [cb.row_begin, cb.row_end) of the band from its table: plan, launch, fallback list ----
    auto eval_rows = [&](const RenderConsts& cb, bool whole) -> int {
        TilePlan plc = pl;
        if (!whole) {
            plc = tile_plan(ctx, p, cb, n_planes, staged);
            if (!plc.ok) return 1;
        }
        TileCfg gc = plc.cfg;
        gc.bm_i0 = g.bm_i0; gc.bm_j0 = g.bm_j0; gc.bm_cols = g.bm_cols; gc.bm_rows = g.bm_rows; gc.bm_pitchw = g.bm_pitchw; gc.ppitch = g.ppitch; gc.r2c = g.r2c;
        const uint32_t units_c = (uint32_t)gc.n_strips * gc.n_segs * n_planes;
        if (!whole) {
            if ((rc = ensure(ctx, ctx->tiles, (size_t)units_c * sizeof(TileRef) + 64))) return rc;
            d_fbcount = (uint32_t*)ctx->tiles.p;
            d_fblist = (TileRef*)((unsigned char*)ctx->tiles.p + 64);
            FG_CUDA(ctx, cudaMemsetAsync(d_fbcount, 0, 64, s)); // every slice starts an empty fallback list
        }
        SkewPlan sk{};
        TriPlan tr{};
        if (staged && skew_ok) tr = tri_plan(ctx, p, cb, n_planes, table_dens);
        if (staged && skew_ok && !tr.ok) sk = skew_plan(ctx, p, cb, n_planes, table_dens);
        uint32_t units_run = units_c;
        if (tr.ok) {
            TriCfg& k = tr.cfg;
            k.bm_i0 = gc.bm_i0; k.bm_j0 = gc.bm_j0; k.bm_cols = gc.bm_cols; k.bm_rows = gc.bm_rows; k.ppitch = gc.ppitch; k.r2c = gc.r2c;
            units_run = (uint32_t)k.n_strips * k.n_segs * n_planes;
            if (units_run > units_c) { // the list was sized for the strip kernel's segments
                if ((rc = ensure(ctx, ctx->tiles, (size_t)units_run * sizeof(TileRef) + 64))) return rc;
                d_fbcount = (uint32_t*)ctx->tiles.p;
                d_fblist = (TileRef*)((unsigned char*)ctx->tiles.p + 64);
                FG_CUDA(ctx, cudaMemsetAsync(d_fbcount, 0, 64, s));
            }
            if (tr.spw == FG_TRI_SPW_A) k_pixelwise_tri<FG_TRI_SPW_A><<<units_run, FG_TRI_THREADS, k.total, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run, k, cb, tab);
            else if (tr.spw == FG_TRI_SPW_B) k_pixelwise_tri<FG_TRI_SPW_B><<<units_run, FG_TRI_THREADS, k.total, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run, k, cb, tab);
            else k_pixelwise_tri<FG_TRI_SPW_C><<<units_run, FG_TRI_THREADS, k.total, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run, k, cb, tab);
            gc.SEG = k.SEG; // the fallback kernel chunks a listed tile by this height
        } else if (sk.ok) {
            SkewCfg& k = sk.cfg;
            k.bm_i0 = gc.bm_i0; k.bm_j0 = gc.bm_j0; k.bm_cols = gc.bm_cols; k.bm_rows = gc.bm_rows; k.ppitch = gc.ppitch; k.r2c = gc.r2c;
            units_run = (uint32_t)k.n_strips * k.n_segs * n_planes;
            if (units_run > units_c) { // the list was sized for the strip kernel's segments
                if ((rc = ensure(ctx, ctx->tiles, (size_t)units_run * sizeof(TileRef) + 64))) return rc;
                d_fbcount = (uint32_t*)ctx->tiles.p;
                d_fblist = (TileRef*)((unsigned char*)ctx->tiles.p + 64);
                FG_CUDA(ctx, cudaMemsetAsync(d_fbcount, 0, 64, s));
            }
            if (sk.spwc == 4) k_pixelwise_skew<4><<<units_run, FG_SK_THREADS, k.total, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run, k, cb, tab);
            else k_pixelwise_skew<8><<<units_run, FG_SK_THREADS, k.total, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run, k, cb, tab);
            gc.SEG = k.SEG; // the fallback kernel chunks a listed tile by this height
        } else if (staged) {
            if (cb.rad.lognorm) launch_strip<true, true>(plc.spwc, units_c, gc.total, s, d_bm, bm_plane_words, d_e, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_c, gc, cb, tab);
            else launch_strip<false, true>(plc.spwc, units_c, gc.total, s, d_bm, bm_plane_words, d_e, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_c, gc, cb, tab);
        } else {
            if (cb.rad.lognorm) launch_strip<true, false>(plc.spwc, units_c, gc.total, s, d_bm, bm_plane_words, d_e, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_c, gc, cb, tab);
            else launch_strip<false, false>(plc.spwc, units_c, gc.total, s, d_bm, bm_plane_words, d_e, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_c, gc, cb, tab);
        }
        FG_CUDA(ctx, cudaGetLastError());
        FG_CUDA(ctx, cudaEventRecord(ctx->ev[5], s));
        const uint32_t chunks = (uint32_t)((gc.SEG + 7) / 8);
        const unsigned fbb = (unsigned)std::min<uint64_t>((uint64_t)units_run * chunks, (uint64_t)ctx->sm_count * 8);
        if (staged && cb.rad.lognorm)
            k_pixelwise_table_tiles<true><<<fbb, 256, 0, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run,
                                                              chunks, d_fbtotal, gc, cb, tab);
        else if (staged)
            k_pixelwise_table_tiles<false><<<fbb, 256, 0, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run,
                                                               chunks, d_fbtotal, gc, cb, tab);
        else
            k_pixelwise_direct_tiles<<<fbb, 256, 0, s>>>(d_lambda, in_stride, off, d_out, out_stride, d_fblist, d_fbcount, units_run,
                                                         chunks, d_fbtotal, cb);
        FG_CUDA(ctx, cudaGetLastError());
        ctx->stats.launches += 2;
        if (std::getenv("FG_B200_DEBUG"))
            std::fprintf(stderr, "[fg] band %d..%d staged=%d TH=%d SEG=%d RH=%d CWB=%d GCAP=%d smem=%u units_c=%u\n", cb.row_begin, cb.row_end,
                         (int)staged, gc.TH, gc.SEG, gc.RH, gc.CWB, gc.GCAP, gc.total, units_c);
        ctx->stats.tiles_total += units_run;
        ctx->strip_launches += 1;
        ctx->eval_kernel = tr.ok ? "k_pixelwise_tri" : (sk.ok ? "k_pixelwise_skew" : "k_pixelwise_strip");
        if (std::getenv("FG_B200_DEBUG") && tr.ok)
            std::fprintf(stderr, "[fg] tri m=%d D=%d AMAX=%d NQ=%d PS=%d MCAP=%d GS=%d RH=%d NB=%d SEG=%d segs=%d smem=%u dens=%.3f\n", tr.cfg.m, tr.cfg.D, tr.cfg.AMAX,
                         tr.cfg.NQ, tr.cfg.PS, tr.cfg.MCAP, tr.cfg.GS, tr.cfg.RH, tr.cfg.NB, tr.cfg.SEG, tr.cfg.n_segs, tr.cfg.total, table_dens);
        if (std::getenv("FG_B200_DEBUG") && sk.ok)
            std::fprintf(stderr, "[fg] skew D=%d PS=%d MCAP=%d TCAP=%d R=%d SEG=%d segs=%d smem=%u dens=%.3f\n", sk.cfg.D, sk.cfg.PS, sk.cfg.MCAP, sk.cfg.TCAP, sk.cfg.R,
                         sk.cfg.SEG, sk.cfg.n_segs, sk.cfg.total, table_dens);
        return FG_OK;
    };
    // A host caller with pageable output planes (fg_ctx::outp.want) gets the band evaluated in three row slices, each
    // followed by an event: the device->host copy of slice e runs under the evaluation of slice e + 1
    // (render_planes_host).  The table is the band's; only the evaluation launch is sliced.
    const int band_rows = c.row_end - c.row_begin;
    const int E = (ctx->outp.want && staged && band_rows >= 768 && ctx->outp.ev[0]) ? 3 : 1;
    ctx->outp.n = 0;
    for (int e = 0; e < E; ++e) {
        RenderConsts cb = c;
        cb.row_begin = c.row_begin + (int)((long long)band_rows * e / E);
        cb.row_end = c.row_begin + (int)((long long)band_rows * (e + 1) / E);
        const int rce = eval_rows(cb, E == 1);
        if (rce) { ctx->outp.n = 0; return rce; } // whoever renders the band instead writes it after the slice events
        if (E > 1) {
            FG_CUDA(ctx, cudaEventRecord(ctx->outp.ev[e], s));
            ctx->outp.row_end[e] = cb.row_end;
            ctx->outp.n = e + 1;
        }
    }
    return FG_OK;
}

// returns FG_OK (rendered), 1 (not applicable: caller uses the direct kernel) or an error.
// path: FG_PATH_AUTO / FG_PATH_STAGED try the cell table first and fall back to in-kernel generation
// (FG_PATH_TILED) when the table does not fit or a row overflowed.
int tile_render(fg_ctx* ctx, const fg_params* p, const RenderConsts& c, int n_planes, const float* d_lambda,
                const float* d_offsets, float* d_out, uint32_t path) {
    int rc;
    if ((rc = ensure(ctx, ctx->fbtotal, 64))) return rc;
    uint32_t* d_fbtotal = (uint32_t*)ctx->fbtotal.p;
    FG_CUDA(ctx, cudaMemsetAsync(d_fbtotal, 0, 4, ctx->stream));
    ctx->strip_launches = 0;
    bool staged = path != FG_PATH_TILED;
    const bool skew_ok = path != FG_PATH_TILED; // FG_PATH_TILED pins k_pixelwise_strip (tests: an independent second evaluation kernel)
    rc = tile_render_band(ctx, p, c, n_planes, d_lambda, d_offsets, d_out, staged, d_fbtotal, skew_ok);
    if (rc == 2 || rc == 3) {
        // 2: the whole band does not fit one table -> row sub-bands; whatever cannot be staged (and
        // 3: a table overflow) is rendered with in-kernel generation
        ctx->outp.want = false; // several bands: one copy of the whole result after the last one
        ctx->outp.n = 0;
        const int band = c.row_end - c.row_begin;
        int done = c.row_begin;
        for (int parts = 2; rc == 2 && parts <= 64 && done == c.row_begin; parts *= 2) {
            const int rows = (band + parts - 1) / parts;
            if (rows < 64) break;
            for (int y = c.row_begin; y < c.row_end; y += rows) {
                RenderConsts cb = c;
                cb.row_begin = y;
                cb.row_end = std::min(y + rows, c.row_end);
                const int r2 = tile_render_band(ctx, p, cb, n_planes, d_lambda, d_offsets, d_out, true, d_fbtotal, skew_ok);
                if (r2 == 1 || r2 == 2 || r2 == 3) break;
                if (r2) return r2;
                done = cb.row_end;
            }
        }
        if (done < c.row_end) { // finish (or redo) the rest with in-kernel generation
            RenderConsts cb = c;
            cb.row_begin = done;
            rc = tile_render_band(ctx, p, cb, n_planes, d_lambda, d_offsets, d_out, false, d_fbtotal, false);
            if (rc) return rc;
        }
        rc = FG_OK;
    }
    if (rc) return rc;
    FG_CUDA(ctx, cudaMemcpyAsync(&ctx->fb_count_host(), d_fbtotal, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->fb_pending = true;
    return FG_OK;
}

} // namespace
