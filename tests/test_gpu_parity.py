"""GPU parity tests: the CUDA path, called through the C ABI (include/fg.h), against the CPU
oracle on the same seeded inputs.  Bar: bit-exact grain realisations and bit-exact f32 pixel
planes (the contractual bound is <= 1/255; both sides do IEEE f32 with no contraction, so
equality is expected and asserted)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import fg_params_from, gradient_u8, lambda_from_u8, noise_u8

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import film_grain_b200 as fg
    c = fg.Context(0)
    yield c
    c.close()


def _cells(n, rng, span=40000):
    ij = rng.integers(-span, span, (n, 2)).astype(np.int32)
    ij[:16] = [[0, 0], [1, 2], [-1, -1], [38399, 21599], [-2147483648, 2147483647], [2147483647, -2147483648],
               [0, -1], [-1, 0], [5, 5], [7, -3], [100, 100], [-100, 100], [1 << 20, 1 << 20], [12345, -54321],
               [3, 4], [4, 3]]
    return ij


@pytest.mark.parametrize("dist", ["const", "lognorm"])
@pytest.mark.parametrize("stream", [O.STREAM_CELL, O.STREAM_PIXEL])
def test_grain_realisation_bit_exact(ctx, dist, stream):
    """K0: counts, centres and radii per cell are bit-identical to the oracle (both Poisson
    branches: Knuth < 12 <= Cauchy rejection; const and lognormal radii; negative cells)."""
    if dist == "const":
        p = O.make_params(radius=0.1, n_samples=1, seed=5489)
    else:
        p = O.make_params(radius=0.1, radius_dist=O.DIST_LOGNORM, radius_stddev=0.05, n_samples=1, seed=987654321987)
    d, _, _ = O.derive_common(p, 64, 64)
    rng = np.random.default_rng(7)
    n = 300_000
    ij = _cells(n, rng)
    scale = 1.0 if stream == O.STREAM_PIXEL else 1.0 / (d.delta * d.delta)
    means = np.concatenate([rng.uniform(0, 4.5, n // 2), rng.uniform(0, 30.0, n // 4), rng.uniform(11.9, 12.1, n // 8),
                            rng.uniform(100.0, 500.0, n - n // 2 - n // 4 - n // 8)])
    lam = (means * scale).astype(np.float32)
    lam[:8] = [0.0, -1.0, 1e-30, 1e-9, 0.3 * scale, 4.4 * scale, 11.99 * scale, 12.0 * scale]
    cap = 12
    q_ref, g_ref = O.gen_cells(p, d, stream, ij, lam, cap)
    q_gpu, g_gpu = ctx.dump_cells(fg_params_from(p, d), stream, ij, lam, cap)
    bad = np.flatnonzero(q_ref != q_gpu)
    assert bad.size == 0, f"{bad.size} cells differ in count, first {bad[:5]}: ref {q_ref[bad[:5]]} gpu {q_gpu[bad[:5]]} lam {lam[bad[:5]]}"
    valid = np.arange(cap)[None, :] < np.minimum(q_ref, cap)[:, None]
    assert np.array_equal(g_ref.view(np.uint32)[valid], g_gpu.view(np.uint32)[valid])
    assert q_ref.max() > 100 and (q_ref == 0).any()


def test_grain_realisation_seeding_variant_b(ctx):
    p = O.make_params(radius=0.1, n_samples=1, seed=42)
    d, _, _ = O.derive_common(p, 8, 8)
    rng = np.random.default_rng(3)
    ij = _cells(20000, rng)
    lam = rng.uniform(0, 400, 20000).astype(np.float32)
    O.lib().fgo_set_seeding_variant(2)
    try:
        q_ref, g_ref = O.gen_cells(p, d, O.STREAM_CELL, ij, lam, 8)
    finally:
        O.lib().fgo_set_seeding_variant(1)
    q_gpu, g_gpu = ctx.dump_cells(fg_params_from(p, d, seeding=1), O.STREAM_CELL, ij, lam, 8)
    assert np.array_equal(q_ref, q_gpu)
    valid = np.arange(8)[None, :] < np.minimum(q_ref, 8)[:, None]
    assert np.array_equal(g_ref.view(np.uint32)[valid], g_gpu.view(np.uint32)[valid])


PIXEL_CASES = [
    # name, w, h, kwargs
    ("r0.1_N16", 96, 64, dict(radius=0.1, n_samples=16)),
    ("r0.05_zoom2", 48, 40, dict(radius=0.05, n_samples=8, zoom=2.0)),
    ("r0.12_N32", 64, 48, dict(radius=0.12, n_samples=32)),
    ("zoom0.5", 80, 64, dict(radius=0.1, n_samples=8, zoom=0.5)),
    ("zoom1.37_size", 50, 30, dict(radius=0.2, n_samples=8, zoom=1.37, size=(77, 41))),
    ("lognorm", 48, 48, dict(radius=0.1, radius_dist=O.DIST_LOGNORM, radius_stddev=0.04, n_samples=8)),
    ("manual_cell_big_lambda", 24, 24, dict(radius=0.1, n_samples=4, cell_delta=1.0)),
    ("abs_max_radius", 40, 40, dict(radius=0.1, n_samples=8, max_radius=("absolute", 0.25))),
    ("N1", 33, 17, dict(radius=0.1, n_samples=1)),
    # multi-strip / multi-segment / multi-step geometry of the tiled kernel
    ("strips_noise_N24", 200, 150, dict(radius=0.1, n_samples=24)),
    ("strips_zoom3_r0.05", 64, 48, dict(radius=0.05, n_samples=12, zoom=3.0)),
    ("strips_zoom0.3", 200, 200, dict(radius=0.1, n_samples=8, zoom=0.3)),
    ("strips_N70", 70, 66, dict(radius=0.1, n_samples=70)),
    ("strips_N130", 40, 70, dict(radius=0.1, n_samples=130)),
    ("strips_N300_two_chunks", 34, 40, dict(radius=0.1, n_samples=300)),
    ("strips_r0.3_sigma2", 90, 80, dict(radius=0.3, n_samples=16, sigma_px=2.0)),
    ("strips_cell_half_rm", 60, 60, dict(radius=0.1, n_samples=8, cell_delta=0.05)),
    # log-normal radii inside the strip kernel (per-grain r^2 ring)
    ("strips_lognorm_ratio0.5_noise", 150, 110, dict(radius=0.1, radius_dist=O.DIST_LOGNORM, radius_stddev=0.05, n_samples=24)),
    ("strips_lognorm_ratio1_zoom2_noise", 60, 50, dict(radius=0.2, radius_dist=O.DIST_LOGNORM, radius_stddev=0.2, n_samples=12, zoom=2.0)),
    ("strips_lognorm_sigma0", 64, 40, dict(radius=0.15, radius_dist=O.DIST_LOGNORM, radius_stddev=0.0, n_samples=8)),
]


@pytest.mark.parametrize("path", [1, 2, 3], ids=["direct", "tiled", "staged"])
@pytest.mark.parametrize("name,w,h,kw", PIXEL_CASES, ids=[c[0] for c in PIXEL_CASES])
def test_pixelwise_matches_oracle(ctx, name, w, h, kw, path):
    p = O.make_params(algo=O.ALGO_PIXEL, **kw)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h, seed=11) if ("zoom" in name or "noise" in name or "N70" in name) else gradient_u8(w, h)
    lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
    ref = O.render_pixelwise(lam, p, d, off_in)
    got = ctx.render_pixelwise(fg_params_from(p, d, path=path), lam, off_in)
    diff = np.abs(ref - got)
    assert diff.max() == 0.0, f"max diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}; {np.count_nonzero(diff)} px differ"
    assert 0.0 < ref.mean() < 1.0


TRI_CASES = [
    # k_pixelwise_tri (fg_tri.cuh): rm == delta, constant radius, 128 < N <= 256
    ("tri_N256_noise", 100, 90, dict(radius=0.1, n_samples=256), "noise"),
    ("tri_N200_partial_warps", 70, 50, dict(radius=0.1, n_samples=200), "noise"),
    ("tri_N129", 40, 37, dict(radius=0.1, n_samples=129), "gradient"),
    ("tri_r0.05_zoom2", 40, 30, dict(radius=0.05, n_samples=160, zoom=2.0), "noise"),
    ("tri_r0.25_zoom0.7_sigma1.6", 90, 70, dict(radius=0.25, n_samples=144, zoom=0.7, sigma_px=1.6), "noise"),
    ("tri_r0.5", 64, 64, dict(radius=0.5, n_samples=192), "gradient"),
    ("tri_size_not_multiple_of_32", 45, 33, dict(radius=0.1, n_samples=256, zoom=1.3, size=(59, 43)), "noise"),
    # dense content: groups of cell rows that do not fit the merged ring are skipped and their samples walk the cell table
    ("tri_dense_gradient", 96, 64, dict(radius=0.1, n_samples=256), "gradient"),
    ("tri_saturated_block", 96, 80, dict(radius=0.1, n_samples=160), "saturated"),
    # N <= 128: the 4- and 8-samples-per-warp instances, chosen by the planner where a step needs few cell rows (cpr <= 5)
    ("tri_N64_zoom2", 60, 44, dict(radius=0.1, n_samples=64, zoom=2.0), "noise"),
    ("tri_N100_zoom4_r0.05", 40, 30, dict(radius=0.05, n_samples=100, zoom=4.0), "noise"),
    ("tri_N33_r0.2", 90, 70, dict(radius=0.2, n_samples=33), "gradient"),
    ("tri_N128_zoom2.5_saturated", 64, 50, dict(radius=0.1, n_samples=128, zoom=2.5), "saturated"),
]


@pytest.mark.parametrize("name,w,h,kw,content", TRI_CASES, ids=[c[0] for c in TRI_CASES])
def test_tri_kernel_matches_oracle(ctx, monkeypatch, name, w, h, kw, content):
    """The merged-triple evaluation kernel against the oracle: full render and a row band, bit for bit; the test also
    pins that the kernel under test really ran (fg_last_eval_kernel).  The planner leaves dense content to the strip
    kernel; here it is told not to, so that the skipped-group path (samples walking the cell table) is exercised."""
    monkeypatch.setenv("FG_B200_TRI_MIN_HEADROOM", "0.05")
    p = O.make_params(algo=O.ALGO_PIXEL, **kw)
    d, off, off_in = O.derive_common(p, w, h)
    img = gradient_u8(w, h) if content == "gradient" else noise_u8(w, h, seed=23)
    if content == "saturated":
        img[20:min(60, h - 4), 30:, :] = 255
    lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
    ref = O.render_pixelwise(lam, p, d, off_in)
    got = ctx.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
    assert ctx.eval_kernel_name() == "k_pixelwise_tri"
    diff = np.abs(ref - got)
    assert diff.max() == 0.0, f"max diff {diff.max()} at {np.unravel_index(diff.argmax(), diff.shape)}; {np.count_nonzero(diff)} px differ"
    oh = ref.shape[0]
    r0, r1 = oh // 3, oh // 3 + max(1, oh // 4)
    band = ctx.render_pixelwise(fg_params_from(p, d, path=3, rows=(r0, r1)), lam, off_in)
    assert np.array_equal(band[r0:r1], ref[r0:r1])


def test_tri_kernel_rows_beyond_2048_visit_four_cell_rows(ctx):
    """From y = 2048 input pixels on, the f32 rounding of (yg -/+ rm) / delta makes some (row, sample) items span FOUR cell
    rows (never below: the oracle's own arithmetic, counted here).  k_pixelwise_tri serves them from two merged triples
    in shared memory; the band across the 2048 boundary must equal the oracle bit for bit."""
    w, h, n = 40, 2200, 160
    p = O.make_params(radius=0.1, n_samples=n, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, w, h)
    rm, dl = np.float32(d.rm), np.float32(d.delta)
    yy = (np.arange(2030, 2200, dtype=np.float32)[:, None] + np.float32(0.5)) - off_in[:, 1][None, :].astype(np.float32)
    span = np.floor(((yy + rm).astype(np.float32) / dl).astype(np.float32)) - np.floor(((yy - rm).astype(np.float32) / dl).astype(np.float32))
    assert (span[:18] == 2).all() and (span[18:] == 3).any(), "the premise: four-row items exist from row 2048 on only"
    lam = lambda_from_u8(noise_u8(w, h, seed=8)[:, :, 0], d.inv_e_pi_r2)
    r0, r1 = 2030, 2200
    ref = O.render_pixelwise(lam, p, d, off_in, y0=r0, y1=r1)
    got = ctx.render_pixelwise(fg_params_from(p, d, path=3, rows=(r0, r1)), lam, off_in)
    assert ctx.eval_kernel_name() == "k_pixelwise_tri"
    assert np.array_equal(got[r0:r1], ref[r0:r1]), f"{np.count_nonzero(got[r0:r1] != ref[r0:r1])} px differ"


@pytest.mark.parametrize("path", [2, 3], ids=["tiled", "staged"])
def test_tiled_path_is_taken_and_fallback_is_exact(ctx, path):
    """The strip kernel serves ordinary content itself; saturated content (u8 255 -> 4.4 grains per
    cell) overflows its grain ring and must come back bit-identical through the fallback list.
    path 2 generates the cell windows inside the strip kernel, path 3 loads them from the cell table."""
    w, h = 160, 140
    p = O.make_params(radius=0.1, n_samples=16, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h, seed=3)
    lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
    got = ctx.render_pixelwise(fg_params_from(p, d, path=path), lam, off_in)
    st = ctx.stats()
    assert st.tiles_total > 0 and st.tiles_fallback == 0, (st.tiles_total, st.tiles_fallback)
    assert np.array_equal(got, O.render_pixelwise(lam, p, d, off_in))
    img2 = img.copy()
    img2[40:, 50:, :] = 255  # saturated block
    lam2 = lambda_from_u8(img2[:, :, 0], d.inv_e_pi_r2)
    got2 = ctx.render_pixelwise(fg_params_from(p, d, path=path), lam2, off_in)
    st2 = ctx.stats()
    assert 0 < st2.tiles_fallback <= st2.tiles_total, (st2.tiles_total, st2.tiles_fallback)
    assert np.array_equal(got2, O.render_pixelwise(lam2, p, d, off_in))
    # lambda' >= 12 somewhere (manual coarse cell): the rejection branch is only in the general kernel
    p3 = O.make_params(radius=0.1, n_samples=4, algo=O.ALGO_PIXEL, cell_delta=0.7)
    d3, _, off_in3 = O.derive_common(p3, w, h)
    lam3 = lambda_from_u8(img2[:, :, 0], d3.inv_e_pi_r2)
    got3 = ctx.render_pixelwise(fg_params_from(p3, d3, path=path), lam3, off_in3)
    assert np.array_equal(got3, O.render_pixelwise(lam3, p3, d3, off_in3))


@pytest.mark.parametrize("dist", ["const", "lognorm"])
def test_staged_fallback_reads_the_cell_table(ctx, dist):
    """Saturated content overflows the strip kernel's grain ring; in staged mode those segments are
    evaluated straight from the HBM cell table (k_pixelwise_table_tiles, both radius models) and must
    equal the oracle and the regenerating direct kernel bit for bit."""
    w, h = 128, 96
    kw = dict(radius=0.1, n_samples=12, algo=O.ALGO_PIXEL)
    if dist == "lognorm":
        kw.update(radius_dist=O.DIST_LOGNORM, radius_stddev=0.05)
    p = O.make_params(**kw)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h, seed=21)
    img[30:, 40:, :] = 255
    lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
    got = ctx.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
    st = ctx.stats()
    assert 0 < st.tiles_fallback <= st.tiles_total, (st.tiles_total, st.tiles_fallback)
    assert np.array_equal(got, O.render_pixelwise(lam, p, d, off_in))
    assert np.array_equal(got, ctx.render_pixelwise(fg_params_from(p, d, path=1), lam, off_in))


def test_staged_path_splits_into_row_bands_when_the_table_budget_is_small(monkeypatch):
    """A render whose cell table exceeds the budget is split into row sub-bands (each with its own
    table); the result and the fallback accounting do not change."""
    w, h = 256, 320
    p = O.make_params(radius=0.1, n_samples=8, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h, seed=9)[:, :, 0], d.inv_e_pi_r2)
    ref = O.render_pixelwise(lam, p, d, off_in)
    import film_grain_b200 as fg
    monkeypatch.setenv("FG_B200_TABLE_MAX_BYTES", str(24 << 20))  # whole image needs ~55 MiB
    with fg.Context(0) as small:
        got = small.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
        st = small.stats()
    assert st.strip_launches >= 2, st.strip_launches
    assert st.tiles_fallback == 0
    assert np.array_equal(ref, got)


def test_staged_path_table_overflow_falls_back_to_in_kernel_generation(monkeypatch):
    """Row capacities are expected grains + 8 sigma; with the slack forced negative every row
    overflows, the gen kernel raises its flag and the band is rendered by the in-kernel generator."""
    w, h = 96, 72
    p = O.make_params(radius=0.1, n_samples=8, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h, seed=4)[:, :, 0], d.inv_e_pi_r2)
    ref = O.render_pixelwise(lam, p, d, off_in)
    import film_grain_b200 as fg
    monkeypatch.setenv("FG_B200_TABLE_SLACK_SIGMA", "-40")
    with fg.Context(0) as tight:
        got = tight.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
        st = tight.stats()
    assert st.strip_launches == 1 and st.tiles_total > 0
    assert np.array_equal(ref, got)


GRAIN_CASES = [
    ("r0.5_N16", 64, 48, dict(radius=0.5, n_samples=16)),
    ("r0.5_N70", 40, 40, dict(radius=0.5, n_samples=70)),
    ("r2_zoom2", 24, 24, dict(radius=2.0, n_samples=8, zoom=2.0)),
    ("r0.1_bigq", 16, 16, dict(radius=0.1, n_samples=8)),
    ("lognorm", 32, 32, dict(radius=0.4, radius_dist=O.DIST_LOGNORM, radius_stddev=0.3, n_samples=8)),
    ("zoom0.6_size", 40, 30, dict(radius=0.8, n_samples=8, zoom=0.6, size=(31, None))),
    ("r0.4_zoom1.5_box2x2", 48, 40, dict(radius=0.4, n_samples=40, zoom=1.5)),  # R = 0.6: boxes of 1-2 pixels per axis
    ("r0.7_box2x2", 40, 36, dict(radius=0.7, n_samples=33)),
    ("r0.5_N2100_two_offset_passes", 20, 16, dict(radius=0.5, n_samples=2100)),  # > 2048 offsets: second shared-memory pass
    ("r0.1_N70_small_disks", 48, 40, dict(radius=0.1, n_samples=70)),             # 2R = 0.2: the sparse (survivor mask) splat
    ("r0.12_zoom2_small_disks", 40, 30, dict(radius=0.12, n_samples=33, zoom=2.0)),
    ("r0.1_wide_4096_small_disks", 4096, 6, dict(radius=0.1, n_samples=24)),      # coordinates up to 4096: f32 ulp 4.9e-4 against the sparse prefilter's slack
    ("r0.5_multi_tile", 300, 150, dict(radius=0.5, n_samples=8)),                 # 3 x 3 output tiles of 128 x 64
    ("r0.3_zoom2_multi_tile", 100, 70, dict(radius=0.3, n_samples=36, zoom=2.0)),
]


@pytest.mark.parametrize("path", [3, 1, 0], ids=["tiled", "global_mask", "auto"])
@pytest.mark.parametrize("name,w,h,kw", GRAIN_CASES, ids=[c[0] for c in GRAIN_CASES])
def test_grainwise_matches_oracle(ctx, name, w, h, kw, path):
    """path 3 (FG_PATH_STAGED): one CTA per output tile, coverage masks in shared memory (k_gw_tile); path 1
    (FG_PATH_DIRECT): the global-mask kernels (k_gw_splat + k_gw_reduce); path 0: the engine's choice."""
    p = O.make_params(algo=O.ALGO_GRAIN, **kw)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h, seed=5)[:, :, 1], d.inv_e_pi_r2)
    ref = O.render_grainwise(lam, p, d, off)
    got = ctx.render_grainwise(fg_params_from(p, d, path=path), lam, off)
    assert np.array_equal(ref, got), f"{np.count_nonzero(ref != got)} px differ, max {np.abs(ref - got).max()}"
    assert 0.0 < ref.mean() < 1.0


@pytest.mark.parametrize("algo,path,kw", [
    ("pixel", 3, dict(radius=0.1, n_samples=8)), ("pixel", 1, dict(radius=0.1, n_samples=8)),
    ("pixel", 2, dict(radius=0.1, n_samples=8)), ("pixel", 3, dict(radius=0.1, n_samples=6, cell_delta=0.7)),
    ("pixel", 3, dict(radius=0.05, n_samples=6, zoom=2.5)),
    ("grain", 3, dict(radius=0.6, n_samples=8, zoom=1.5)), ("grain", 1, dict(radius=0.6, n_samples=8, zoom=1.5)),
], ids=["pixel-staged", "pixel-direct", "pixel-tiled", "pixel-coarse-cell", "pixel-zoom2.5", "grain-tiled", "grain-global-mask"])
def test_row_bands_equal_full_render(ctx, algo, path, kw):
    """multi-GPU contract: rendering disjoint row bands reproduces the full render bit for bit.  A band call
    uploads only the input rows the band can read; conftest.py sets FG_B200_POISON, so every other row of
    the device lambda buffer holds NaN during the call."""
    w, h = 72, 57
    p = O.make_params(algo=O.ALGO_PIXEL if algo == "pixel" else O.ALGO_GRAIN, **kw)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h, seed=9)[:, :, 2], d.inv_e_pi_r2)
    offs = off_in if algo == "pixel" else off
    render = ctx.render_pixelwise if algo == "pixel" else ctx.render_grainwise
    full = render(fg_params_from(p, d, path=path), lam, offs)
    oh = d.output_height
    cuts = [0, oh // 3, oh // 3 + 1, (2 * oh) // 3, oh]
    banded = np.full_like(full, -1.0)
    for a, b in zip(cuts[:-1], cuts[1:]):
        render(fg_params_from(p, d, rows=(a, b), path=path), lam, offs, out=banded)
        st = ctx.stats()
        assert st.h2d_bytes <= (h * w + 2 * p.n_samples) * 4
    assert np.array_equal(full, banded)
    ref = O.render_pixelwise(lam, p, d, off_in) if algo == "pixel" else O.render_grainwise(lam, p, d, off)
    assert np.array_equal(full, ref)


def test_planes_batched_equals_sequential(ctx):
    w, h = 40, 36
    p = O.make_params(radius=0.1, n_samples=8, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h, seed=2)
    lams = [lambda_from_u8(img[:, :, c], d.inv_e_pi_r2) for c in range(3)]
    q = fg_params_from(p, d)
    outs = ctx.render_planes(q, 2, lams, off_in)
    for c in range(3):
        assert np.array_equal(outs[c], O.render_pixelwise(lams[c], p, d, off_in))


@pytest.mark.parametrize("dist", ["const", "lognorm"])
@pytest.mark.parametrize("joint", ["1", "0"], ids=["joint", "per-plane"])
def test_three_plane_table_generation_on_mixed_content(ctx, monkeypatch, dist, joint):
    """k_gen_rows<., 3> generates a cell row for the three planes from ONE seeding and ONE Knuth chain per cell.  Planes
    that are equal, black (lambda = 0), saturated (lambda' >= 12: the general sampler, parked from the seed state),
    dense (several grains per cell: skipped outputs) and independent noise, all in one image; every plane must equal
    the oracle's render of that plane alone, and the per-plane generator (FG_B200_GEN_JOINT=0) must agree."""
    monkeypatch.setenv("FG_B200_GEN_JOINT", joint)
    w, h = 72, 48
    kw = dict(radius=0.1, n_samples=12)
    if dist == "lognorm":
        kw.update(radius_dist=O.DIST_LOGNORM, radius_stddev=0.05)
    p = O.make_params(algo=O.ALGO_PIXEL, **kw)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h, seed=31)
    img[:, :, 1] = img[:, :, 0]            # G == R: every cell stops at the same count in both
    img[:12, :, 2] = 0                     # B black on top ...
    img[12:24, :, 2] = 255                 # ... saturated below it (general sampler beside Knuth planes)
    img[24:36, 20:50, :] = 255             # all planes saturated
    img[36:, :24, 0] = 250                 # dense Knuth cells (lambda' ~ 1.76) beside sparse ones
    lams = [lambda_from_u8(img[:, :, c], d.inv_e_pi_r2) for c in range(3)]
    lams[1][40:, 30:40] = np.float32(11.9) / np.float32(d.delta * d.delta)   # lambda' just under the Knuth limit
    lams[2][40:, 30:40] = np.float32(12.1) / np.float32(d.delta * d.delta)   # ... and just over it
    outs = ctx.render_planes(fg_params_from(p, d, path=3), 2, lams, off_in)
    for c in range(3):
        ref = O.render_pixelwise(lams[c], p, d, off_in)
        assert np.array_equal(outs[c], ref), f"plane {c}: {np.count_nonzero(outs[c] != ref)} px differ"


@pytest.mark.parametrize("mode,algo", [(1, O.ALGO_PIXEL), (0, O.ALGO_PIXEL), (1, O.ALGO_GRAIN), (0, O.ALGO_GRAIN)],
                         ids=["rgb-pixel", "luma-pixel", "rgb-grain", "luma-grain"])
def test_rgb8_pipeline_matches_oracle(ctx, mode, algo):
    """lib.rs:134-173 end to end on 8-bit images: fused load/lambda/store vs the oracle pipeline.
    RGB mode uses a host-built lambda table; luma mode computes logf on the device with the libm
    algorithm restated in csrc/fg_logf.h -- both are bit-exact."""
    w, h = 45, 33
    kw = dict(radius=0.1, n_samples=16) if algo == O.ALGO_PIXEL else dict(radius=0.5, n_samples=16)
    p = O.make_params(algo=algo, zoom=1.5, **kw)
    img = noise_u8(w, h, seed=21)
    ref, used = O.render_rgb8(img, p, mode)
    assert used == algo
    d, off, off_in = O.derive_common(p, w, h)
    offs = off_in if algo == O.ALGO_PIXEL else off
    got = ctx.render_rgb8(fg_params_from(p, d), algo, mode, img, offs)
    assert np.array_equal(ref, got), f"{np.count_nonzero(ref != got)} of {ref.size} bytes differ"


def test_fused_luma_lambda_is_bit_exact_on_a_large_image(ctx):
    """every 8-bit (r,g,b) combination of a 256x192 noise image goes through the device logf: the
    fused luma render must equal the oracle pipeline byte for byte"""
    w, h = 256, 192
    p = O.make_params(radius=0.1, n_samples=4, algo=O.ALGO_PIXEL)
    img = noise_u8(w, h, seed=77)
    ref, _ = O.render_rgb8(img, p, 0)
    d, off, off_in = O.derive_common(p, w, h)
    got = ctx.render_rgb8(fg_params_from(p, d), O.ALGO_PIXEL, 0, img, off_in)
    assert np.array_equal(ref, got)


def test_config1_512_gradient_luma_n64(ctx):
    """BASELINE.json configs[0]: pixel-wise luma 512x512 gradient, r=0.1, N=64, seed 5489."""
    w = h = 512
    p = O.make_params(radius=0.1, n_samples=64, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    assert O.choose_algorithm(p, d) == O.ALGO_PIXEL
    y = np.empty(w * h, np.float32); cb = np.empty_like(y); cr = np.empty_like(y)
    img = gradient_u8(w, h)
    O.lib().fgo_load_luma_u8(img.ctypes.data_as(C.POINTER(C.c_uint8)), w * h, y.ctypes.data_as(C.POINTER(C.c_float)),
                             cb.ctypes.data_as(C.POINTER(C.c_float)), cr.ctypes.data_as(C.POINTER(C.c_float)))
    lam = O.lambda_plane(O.normalize_plane(y.reshape(h, w)), d.inv_e_pi_r2)
    ref = O.render_pixelwise(lam, p, d, off_in)
    got = ctx.render_pixelwise(fg_params_from(p, d), lam, off_in)
    assert np.array_equal(ref, got)
    # Boolean-model identity E[pixel] = u (src/model.rs:192-194, 252-265): column means follow the ramp
    u = img[0, :, 0].astype(np.float64) / 255.0
    assert np.abs(got.mean(axis=0)[8:-8] - u[8:-8]).mean() < 0.02


def test_errors_do_not_poison_context(ctx):
    import film_grain_b200 as fg
    p = O.make_params(radius=0.1, n_samples=4, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, 16, 16)
    lam = np.ones((16, 16), np.float32)
    q = fg_params_from(p, d)
    q.struct_size = 12
    with pytest.raises(fg.GpuError) as e:
        ctx.render_pixelwise(q, lam, off_in)
    assert e.value.code == -1
    q = fg_params_from(p, d, rows=(9, 4))
    with pytest.raises(fg.GpuError):
        ctx.render_pixelwise(q, lam, off_in)
    bad = off_in.copy(); bad[0, 0] = np.nan
    with pytest.raises(fg.GpuError):
        ctx.render_pixelwise(fg_params_from(p, d), lam, bad)
    out = ctx.render_pixelwise(fg_params_from(p, d), lam, off_in)  # still usable (Validation keeps the context)
    assert np.array_equal(out, O.render_pixelwise(lam, p, d, off_in))


def test_cancel_flag(ctx):
    import film_grain_b200 as fg
    p = O.make_params(radius=0.1, n_samples=4, algo=O.ALGO_PIXEL)
    d, off, off_in = O.derive_common(p, 16, 16)
    lam = np.ones((16, 16), np.float32)
    flag = C.c_int(1)
    ctx.set_cancel_flag(flag)
    try:
        with pytest.raises(fg.Cancelled):
            ctx.render_pixelwise(fg_params_from(p, d), lam, off_in)
    finally:
        ctx.set_cancel_flag(None)
    ctx.render_pixelwise(fg_params_from(p, d), lam, off_in)


@pytest.mark.parametrize("algo,path", [("pixel", 1), ("pixel", 3), ("grain", 0)], ids=["pixel-direct", "pixel-staged", "grain"])
def test_cancel_flag_raised_mid_render_stops_the_launch(algo, path):
    """The viewer's latest-job-wins worker (src/bin/viewer.rs:975-1028) cancels a render that is already running.  A
    second thread raises the flag while the kernels execute: the call must return Cancelled well before an
    uncancelled render of the same frame would have finished, and the context must render correctly afterwards."""
    import threading
    import time
    import film_grain_b200 as fg
    w, h = (3072, 2048) if path == 3 else (2048, 1024)
    if algo == "pixel":
        p = O.make_params(radius=0.1, n_samples=256 if path == 3 else 48, algo=O.ALGO_PIXEL)
    else:
        p = O.make_params(radius=0.5, n_samples=512, algo=O.ALGO_GRAIN)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h, seed=3)[:, :, 0], d.inv_e_pi_r2)
    lams = [lam, lam, lam]
    offs = off_in if algo == "pixel" else off
    a = O.ALGO_PIXEL if algo == "pixel" else O.ALGO_GRAIN
    q = fg_params_from(p, d, path=path)
    with fg.Context(0) as c:
        c.render_planes(q, a, lams, offs)  # warm-up: pools, attributes
        flag = C.c_int(0)
        c.set_cancel_flag(flag)            # armed but never raised: the kernels poll the device word, the result is unchanged
        t0 = time.perf_counter()
        full = c.render_planes(q, a, lams, offs)
        t_full = time.perf_counter() - t0
        assert t_full > 0.02, "the frame must be long enough to cancel"
        for delay in (0.2, 0.4):  # pixel-wise staged: inside the table pass, inside the evaluation kernel
            best = None
            for attempt in range(3):  # a shared box can stall any single attempt; the mechanism is judged by its best
                flag.value = 0
                raised = [0.0]

                def raise_flag():
                    raised[0] = time.perf_counter()
                    flag.value = 1
                th = threading.Timer(delay * t_full, raise_flag)
                th.start()
                with pytest.raises(fg.Cancelled):
                    c.render_planes(q, a, lams, offs)
                t_after = time.perf_counter() - raised[0]
                th.join()
                best = t_after if best is None else min(best, t_after)
            # an uncancelled frame would run (1 - delay) * t_full >= 0.6 * t_full past the raise; a cancelled one stops
            # within a CTA lifetime plus the drain of the queued launches (measured: ~1 ms on a 4K frame)
            assert best < 0.3 * t_full, (delay, best, t_full)
        flag.value = 0
        again = c.render_planes(q, a, lams, offs)  # the device word was lowered again
        c.set_cancel_flag(None)
        for k in range(3):
            assert np.array_equal(again[k], full[k])
    if algo == "pixel":  # a few rows of the cancelled-then-repeated frame against the oracle
        ref = O.render_pixelwise(lam, p, d, off_in, y0=500, y1=504)
        assert np.array_equal(again[0][500:504], ref[500:504])


def test_table_cache_serves_renders_that_change_only_samples_sigma_and_zoom():
    """fg_set_table_cache: N, the offsets (sigma) and the zoom do not enter the cell table, so the second and later renders
    of a viewer session evaluate from the cached table (stats.table_reused) -- and must equal a from-scratch render and
    the oracle bit for bit.  New content, a new seed or a rectangle the cached table does not cover rebuild it."""
    import film_grain_b200 as fg
    w, h = 160, 120
    img = noise_u8(w, h, seed=41)
    with fg.Context(0) as c, fg.Context(0) as fresh:
        c.set_table_cache(True)
        def render(ctxx, **kw):
            p = O.make_params(algo=O.ALGO_PIXEL, **kw)
            d, off, off_in = O.derive_common(p, w, h)
            lams = [lambda_from_u8(img[:, :, k], d.inv_e_pi_r2) for k in range(3)]
            out = ctxx.render_planes(fg_params_from(p, d, path=3), O.ALGO_PIXEL, lams, off_in)
            return out, ctxx.stats().table_reused, (p, d, off_in, lams)
        base = dict(radius=0.1, n_samples=16)
        _, reused, _ = render(c, **base)
        assert reused == 0
        for kw, expect in [(dict(base, n_samples=64), 1),                       # more samples (same sigma: offsets are a prefix)
                           (dict(base, n_samples=200), 1),                      # k_pixelwise_tri from the cached table
                           (dict(base, n_samples=32, sigma_px=0.6), 1),         # smaller blur
                           (dict(base, n_samples=32, zoom=1.5), 1),             # zoom in: a smaller cell rectangle
                           (dict(base, n_samples=32, sigma_px=4.0), 0),         # much larger blur: outside the margin -> rebuild
                           (dict(base, n_samples=32, sigma_px=3.0), 1),         # ... and the rebuilt (larger) table serves this
                           (dict(base, n_samples=32, seed=77), 0),              # another realisation
                           (dict(base, n_samples=32, seed=77, zoom=0.8), 1)]:
            got, reused, (p, d, off_in, lams) = render(c, **kw)
            assert reused == expect, (kw, reused)
            ref, r0, _ = render(fresh, **kw)
            assert r0 == 0
            for k in range(3):
                assert np.array_equal(got[k], ref[k]), kw
            assert np.array_equal(got[1], O.render_pixelwise(lams[1], p, d, off_in)), kw
        img[10:20, 10:20, 1] ^= 0x55  # new content in one plane: the hash changes
        _, reused, _ = render(c, **base)
        assert reused == 0
        c.set_table_cache(False)
        _, reused, _ = render(c, **base)
        assert reused == 0


@pytest.mark.parametrize("algo", ["pixel", "grain"])
def test_progressive_refinement_equals_one_render(ctx, algo):
    """fg_refine_planes: 16 -> 48 -> 200 samples through the same table; after every step the image equals a single render
    of that many samples (the oracle's), bit for bit; a slice that does not continue the previous one is refused."""
    import film_grain_b200 as fg
    w, h, n = 90, 70, 200
    kw = dict(radius=0.1) if algo == "pixel" else dict(radius=0.5)
    a = O.ALGO_PIXEL if algo == "pixel" else O.ALGO_GRAIN
    p = O.make_params(algo=a, n_samples=n, **kw)
    d, off, off_in = O.derive_common(p, w, h)
    offs = off_in if algo == "pixel" else off
    img = noise_u8(w, h, seed=5)
    lams = [lambda_from_u8(img[:, :, k], d.inv_e_pi_r2) for k in range(2)]
    q = fg_params_from(p, d)
    ctx.set_table_cache(True)
    try:
        outs = None
        k0 = 0
        for k1 in (16, 48, 200):
            outs = ctx.refine_planes(q, a, lams, offs, k0, k1, outs)
            if algo == "pixel" and k0 > 0:
                assert ctx.stats().table_reused in (0, 1)
            pk = O.make_params(algo=a, n_samples=k1, **kw)
            dk, offk, offk_in = O.derive_common(pk, w, h)
            o_k = offk_in if algo == "pixel" else offk
            assert np.array_equal(o_k, offs[:k1]), "the offsets of a smaller N are a prefix of the larger N's"
            for k in range(2):
                ref = O.render_pixelwise(lams[k], pk, dk, offk_in) if algo == "pixel" else O.render_grainwise(lams[k], pk, dk, offk)
                assert np.array_equal(outs[k], ref), (k1, k)
            k0 = k1
        with pytest.raises(fg.GpuError):
            ctx.refine_planes(q, a, lams, offs, 48, 64, outs)  # the context's refinement stands at 200
    finally:
        ctx.set_table_cache(False)


def test_full_size_config2_plane_properties(ctx):
    """BASELINE.json configs[1] at full size (3840x2160, r=0.1, N=256, one colour plane): the strip
    kernel against (a) the oracle on two 6-row bands, (b) the independent direct kernel on three
    48-row bands (top edge, middle, bottom edge), (c) the Boolean-model identity E[pixel] = u on the image mean."""
    w, h, n = 3840, 2160, 256
    p = O.make_params(radius=0.1, n_samples=n, algo=O.ALGO_PIXEL, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h)  # the bench input
    lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
    full = ctx.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
    st = ctx.stats()
    assert st.tiles_total > 0 and st.tiles_fallback == 0
    tiled = ctx.render_pixelwise(fg_params_from(p, d, path=2), lam, off_in)
    assert np.array_equal(tiled, full)  # in-kernel generation == cell table, every pixel
    for a, b in ((0, 6), (1237, 1243)):
        ref = O.render_pixelwise(lam, p, d, off_in, a, b)
        assert np.array_equal(ref[a:b], full[a:b])
    direct = np.full_like(full, -1.0)
    for a, b in ((0, 48), (1050, 1098), (2112, 2160)):
        ctx.render_pixelwise(fg_params_from(p, d, path=1, rows=(a, b)), lam, off_in, out=direct)
        assert np.array_equal(direct[a:b], full[a:b])
    # iid 8-bit noise is the worst case for the identity: grains of radius r straddle pixels of very
    # different lambda and 1-exp(-x) is concave, so coverage is biased upwards (Jensen) by ~0.013;
    # the image mean still has to land next to the input mean
    u = img[:, :, 0].astype(np.float64) / 255.0
    assert 0.0 <= full.mean() - u.mean() < 0.03


def test_full_size_config4_band(ctx):
    """BASELINE.json configs[3] geometry (2048^2 input, zoom 4, r=0.05, N=64): one 96-row output band,
    strip kernel == direct kernel, and == the oracle on 4 rows."""
    w = h = 2048
    p = O.make_params(radius=0.05, n_samples=64, zoom=4.0, algo=O.ALGO_PIXEL, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    assert (d.output_width, d.output_height) == (8192, 8192)
    lam = lambda_from_u8(noise_u8(w, h)[:, :, 1], d.inv_e_pi_r2)
    a, b = 4000, 4096
    tiled = np.zeros((8192, 8192), np.float32)
    direct = np.zeros((8192, 8192), np.float32)
    staged = np.zeros((8192, 8192), np.float32)
    ctx.render_pixelwise(fg_params_from(p, d, path=2, rows=(a, b)), lam, off_in, out=tiled)
    ctx.render_pixelwise(fg_params_from(p, d, path=3, rows=(a, b)), lam, off_in, out=staged)
    ctx.render_pixelwise(fg_params_from(p, d, path=1, rows=(a, b)), lam, off_in, out=direct)
    assert np.array_equal(tiled[a:b], direct[a:b])
    assert np.array_equal(staged[a:b], direct[a:b])
    ref = O.render_pixelwise(lam, p, d, off_in, a, a + 4)
    assert np.array_equal(ref[a:a + 4], tiled[a:a + 4])


def test_full_size_config3_grainwise_crop_consistency(ctx):
    """BASELINE.json configs[2] parameters (grain-wise, r=0.5, N=128): a 1024x1024 render equals the
    oracle on the same input (the oracle needs ~10 s for this size), and a band split reproduces it."""
    w = h = 1024
    p = O.make_params(radius=0.5, n_samples=128, algo=O.ALGO_GRAIN, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h)[:, :, 0], d.inv_e_pi_r2)
    got = ctx.render_grainwise(fg_params_from(p, d), lam, off)
    ref = O.render_grainwise(lam, p, d, off)
    assert np.array_equal(ref, got)
    banded = np.zeros_like(got)
    for a, b in ((0, 300), (300, 301), (301, 1024)):
        ctx.render_grainwise(fg_params_from(p, d, rows=(a, b)), lam, off, out=banded)
    assert np.array_equal(banded, got)


def test_full_size_config3_grainwise_4096_against_the_oracle(ctx):
    """BASELINE.json configs[2] at FULL size (grain-wise luma 4096x4096 noise, r = 0.5, N = 128; 21 M grains, 2048
    output tiles): the shared-memory tile rasteriser, the global-mask rasteriser and the engine's own pick, every
    pixel against the oracle's full render (about 16 s on 8 host threads)."""
    w = h = 4096
    p = O.make_params(radius=0.5, n_samples=128, algo=O.ALGO_GRAIN, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    lam = lambda_from_u8(noise_u8(w, h)[:, :, 0], d.inv_e_pi_r2)
    ref = O.render_grainwise(lam, p, d, off)
    for path in (3, 1, 0):
        got = ctx.render_grainwise(fg_params_from(p, d, path=path), lam, off)
        bad = np.flatnonzero(ref != got)
        assert bad.size == 0, f"path {path}: {bad.size} px differ, first at {np.unravel_index(bad[0], ref.shape)}"
    assert 0.5 < ref.mean() < 0.6


def test_full_size_config2_three_plane_render_against_the_oracle(ctx):
    """BASELINE.json configs[1] exactly as bench.py times it -- ONE batched fg_render_planes call over the three colour
    planes of the 3840x2160 noise image, r = 0.1, N = 256 -- against the oracle on 48 rows of every plane: 16 at the
    top edge, 16 in the middle, 16 at the bottom edge (144 rows x 3840 px x 256 samples = 1.4e8 oracle evaluations)."""
    w, h, n = 3840, 2160, 256
    p = O.make_params(radius=0.1, n_samples=n, algo=O.ALGO_PIXEL, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    img = noise_u8(w, h)
    lams = [lambda_from_u8(img[:, :, c], d.inv_e_pi_r2) for c in range(3)]
    outs = ctx.render_planes(fg_params_from(p, d), O.ALGO_PIXEL, lams, off_in)
    st = ctx.stats()
    assert st.tiles_total > 0
    for c in range(3):
        for a, b in ((0, 16), (1072, 1088), (2144, 2160)):
            ref = O.render_pixelwise(lams[c], p, d, off_in, a, b)
            bad = np.flatnonzero(ref[a:b] != outs[c][a:b])
            assert bad.size == 0, f"plane {c} rows {a}..{b}: {bad.size} px differ"


def _multi_devices():
    import film_grain_b200 as fg
    n = fg.device_count()
    return list(range(min(n, 4))) if n >= 2 else [0]


@pytest.mark.parametrize("algo", ["pixel", "grain"])
def test_single_process_multi_device_context_equals_single_device(ctx, algo):
    """fg_context_create_multi (the reference's caller is ONE process, src/lib.rs:141-163): a context over several devices
    splits every render into row bands, one host thread per device.  Host planes (pageable and pinned), a row-band request,
    the u8 pipeline and the device-pointer entry point (bands stored into device 0's image) all reproduce the single-device
    render bit for bit.  On a one-GPU box the same code runs with the single device [0] (routing, band arithmetic, stats)."""
    import torch

    import film_grain_b200 as fg
    devs = _multi_devices()
    w, h = 230, 170
    if algo == "pixel":
        p = O.make_params(radius=0.1, n_samples=24, algo=O.ALGO_PIXEL)
    else:
        p = O.make_params(radius=0.5, n_samples=40, algo=O.ALGO_GRAIN)
    d, off, off_in = O.derive_common(p, w, h)
    offs = off_in if algo == "pixel" else off
    a = O.ALGO_PIXEL if algo == "pixel" else O.ALGO_GRAIN
    img = noise_u8(w, h, seed=21)
    lams = [lambda_from_u8(img[:, :, c], d.inv_e_pi_r2) for c in range(3)]
    ref = ctx.render_planes(fg_params_from(p, d), a, lams, offs)
    oracle0 = O.render_pixelwise(lams[0], p, d, off_in) if algo == "pixel" else O.render_grainwise(lams[0], p, d, off)
    assert np.array_equal(ref[0], oracle0)
    with fg.Context(devices=devs) as mctx:
        assert mctx.device_count() == len(devs)
        got = mctx.render_planes(fg_params_from(p, d), a, lams, offs)
        st = mctx.stats()
        for c in range(3):
            assert np.array_equal(got[c], ref[c])
        assert st.launches > 0 and st.d2h_bytes == 3 * d.output_height * d.output_width * 4
        one = mctx.render_pixelwise(fg_params_from(p, d), lams[1], offs) if algo == "pixel" else mctx.render_grainwise(fg_params_from(p, d), lams[1], offs)
        assert np.array_equal(one, ref[1])
        # a row-band request is split again; rows outside stay untouched
        pin = torch.full((3, d.output_height, d.output_width), -7.0, dtype=torch.float32).pin_memory()
        outs = [pin[c].numpy() for c in range(3)]
        rows = (31, 149)
        mctx.render_planes(fg_params_from(p, d, rows=rows), a, lams, offs, outs)
        for c in range(3):
            assert np.array_equal(outs[c][rows[0]:rows[1]], ref[c][rows[0]:rows[1]])
            assert np.all(outs[c][:rows[0]] == -7.0) and np.all(outs[c][rows[1]:] == -7.0)
        # device-resident planes on devices[0]: the other devices store their bands into its image
        dev0 = torch.device("cuda", devs[0])
        d_lam = torch.from_numpy(np.stack(lams)).to(dev0)
        d_off = torch.from_numpy(np.ascontiguousarray(offs)).to(dev0)
        d_out = torch.full((3, d.output_height, d.output_width), -1.0, dtype=torch.float32, device=dev0)
        torch.cuda.synchronize(dev0)
        mctx.render_planes_device(fg_params_from(p, d), a, 3, d_lam.data_ptr(), d_off.data_ptr(), d_out.data_ptr(), sync=True)
        res = d_out.cpu().numpy()
        for c in range(3):
            assert np.array_equal(res[c], ref[c])
        if algo == "pixel":
            ref8 = ctx.render_rgb8(fg_params_from(p, d), a, 1, img, offs)
            got8 = mctx.render_rgb8(fg_params_from(p, d), a, 1, img, offs)
            assert np.array_equal(got8, ref8)
        del d_lam, d_off, d_out


def test_multi_device_context_rejects_bad_device_lists():
    import ctypes as C2

    import film_grain_b200 as fg
    from film_grain_b200 import _lib
    lib = _lib.load()
    h = C2.c_void_p()
    assert lib.fg_context_create_multi(C2.byref(h), (C2.c_int * 2)(0, 0), 2) == _lib.FG_ERR_INVALID  # duplicate
    assert lib.fg_context_create_multi(C2.byref(h), (C2.c_int * 1)(fg.device_count()), 1) == _lib.FG_ERR_NO_DEVICE
    assert lib.fg_context_create_multi(C2.byref(h), None, 0) == _lib.FG_ERR_INVALID
    assert not h.value


def test_multi_gpu_bands_into_peer_image_equal_single_gpu_render():
    """N > 1 (needs >= 2 GPUs, skipped otherwise): every rank renders its row band straight into GPU 0's
    peer-mapped image over NVLink (film_grain_b200/dist.py PeerImage) and, separately, through the NCCL
    gather; both assembled images must be bitwise identical to the single-GPU render
    (tools/p2p_check.py under torchrun, pixel-wise zoom 1 / zoom 2.5 and grain-wise)."""
    import os
    import subprocess
    import sys

    import film_grain_b200 as fg
    if fg.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "p2p_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout and r.stdout.count("bitwise equal") >= 9, r.stdout


@pytest.mark.parametrize("algo", ["pixel", "grain"])
def test_pinned_contiguous_output_planes_are_written_in_place(ctx, algo):
    """fg_render_planes with output planes that are one block of page-locked host memory: the kernels store
    straight into the mapped block (no staged device->host copy); the result equals the pageable-buffer call
    and the oracle, for a full render and for a row band (rows outside the band stay untouched)."""
    import torch
    w, h = 200, 120
    if algo == "pixel":
        p = O.make_params(radius=0.1, n_samples=24, algo=O.ALGO_PIXEL)
    else:
        p = O.make_params(radius=0.5, n_samples=40, algo=O.ALGO_GRAIN)
    d, off, off_in = O.derive_common(p, w, h)
    offs = off_in if algo == "pixel" else off
    a = O.ALGO_PIXEL if algo == "pixel" else O.ALGO_GRAIN
    img = noise_u8(w, h, seed=11)
    lams = [lambda_from_u8(img[:, :, c], d.inv_e_pi_r2) for c in range(3)]
    ref = ctx.render_planes(fg_params_from(p, d), a, lams, offs)
    oracle0 = O.render_pixelwise(lams[0], p, d, off_in) if algo == "pixel" else O.render_grainwise(lams[0], p, d, off)
    assert np.array_equal(ref[0], oracle0)
    pin = torch.full((3, d.output_height, d.output_width), -7.0, dtype=torch.float32).pin_memory()
    outs = [pin[c].numpy() for c in range(3)]
    ctx.render_planes(fg_params_from(p, d), a, lams, offs, outs)
    st = ctx.stats()
    for c in range(3):
        assert np.array_equal(outs[c], ref[c])
    assert st.d2h_bytes == 3 * d.output_height * d.output_width * 4
    pin.fill_(-7.0)
    rows = (37, 90)
    ctx.render_planes(fg_params_from(p, d, rows=rows), a, lams, offs, outs)
    for c in range(3):
        assert np.array_equal(outs[c][rows[0]:rows[1]], ref[c][rows[0]:rows[1]])
        assert np.all(outs[c][:rows[0]] == -7.0) and np.all(outs[c][rows[1]:] == -7.0)


def test_auto_path_regenerates_when_samples_per_cell_are_few(ctx):
    """FG_PATH_AUTO: the cell table costs the same whatever N is, so with few samples per cell the engine
    regenerates per sample (k_pixelwise_direct, the reference's structure) and builds the table otherwise.
    Same pixels either way."""
    w, h = 160, 96
    img = noise_u8(w, h, seed=4)
    for n, expect_table in ((4, False), (64, True)):
        p = O.make_params(radius=0.1, n_samples=n, algo=O.ALGO_PIXEL)
        d, off, off_in = O.derive_common(p, w, h)
        lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
        got = ctx.render_pixelwise(fg_params_from(p, d, path=0), lam, off_in)
        st = ctx.stats()
        assert (st.tiles_total > 0) == expect_table, (n, st.tiles_total)
        assert np.array_equal(got, O.render_pixelwise(lam, p, d, off_in))


def test_full_size_config5_both_algorithms(ctx):
    """BASELINE.json configs[4] geometry (luma 1024^2, r = 0.12: delta = 1/9 with rm = 0.12, i.e. 3-4 cells per
    axis).  Pixel-wise at N = 1024 (four sample chunks): a 24-row band, staged == tiled == direct, and == the
    oracle on 3 rows.  Grain-wise at N = 256 (sub-pixel disks): the tile rasteriser, the dense global-mask
    rasteriser (sparse instance) and the engine's own pick agree bit for bit on the whole image, whose mean
    reproduces the input's (the Boolean-model identity E[pixel] = u)."""
    w = h = 1024
    img = noise_u8(w, h)[:, :, 0]
    p = O.make_params(radius=0.12, n_samples=1024, algo=O.ALGO_PIXEL, seed=5489)
    d, off, off_in = O.derive_common(p, w, h)
    assert abs(d.delta - 1.0 / 9.0) < 1e-7
    lam = lambda_from_u8(img, d.inv_e_pi_r2)
    a, b = 500, 524
    outs = {}
    for path in (1, 2, 3):
        outs[path] = np.zeros((h, w), np.float32)
        ctx.render_pixelwise(fg_params_from(p, d, path=path, rows=(a, b)), lam, off_in, out=outs[path])
    assert np.array_equal(outs[2][a:b], outs[1][a:b]) and np.array_equal(outs[3][a:b], outs[1][a:b])
    ref = O.render_pixelwise(lam, p, d, off_in, a, a + 3)
    assert np.array_equal(ref[a:a + 3], outs[3][a:a + 3])
    assert 0.3 < outs[3][a:b].mean() < 0.7

    pg = O.make_params(radius=0.12, n_samples=256, algo=O.ALGO_GRAIN, seed=5489)
    dg, offg, _ = O.derive_common(pg, w, h)
    lamg = lambda_from_u8(img, dg.inv_e_pi_r2)
    tile = ctx.render_grainwise(fg_params_from(pg, dg, path=3), lamg, offg)
    glob = ctx.render_grainwise(fg_params_from(pg, dg, path=1), lamg, offg)
    auto = ctx.render_grainwise(fg_params_from(pg, dg, path=0), lamg, offg)
    assert np.array_equal(tile, glob) and np.array_equal(auto, glob)
    assert abs(float(glob.mean()) - float(img.mean()) / 255.0) < 0.03  # E[pixel] ~ u (Boolean-model identity; exact for constant input)


def test_mixed_row_counts_never_read_rows_outside_the_window(ctx):
    """r = 0.12 (delta = 1/9, rm = 0.12): samples visit 3 or 4 cell rows, so a pair can mix row counts and the
    shorter sample idles while the longer one finishes.  It used to step on to the next ring row -- possibly
    one the window never loaded, whose arbitrary prefix value sent the unconditional slot loads outside the
    grain ring (an illegal shared-memory read on 1024^2 'natural', N = 128, seed 5490, found by the
    benchmarks/ sweep).  Same configuration: runs, and staged == tiled == direct on a band."""
    from tools.bench_sweep_b200 import intensity_field
    img = intensity_field("natural", 1024)
    for seed in (5489, 5490, 5491):
        p = O.make_params(radius=0.12, n_samples=128, algo=O.ALGO_PIXEL, seed=seed)
        d, off, off_in = O.derive_common(p, 1024, 1024)
        lam = lambda_from_u8(img[:, :, 0], d.inv_e_pi_r2)
        full = ctx.render_pixelwise(fg_params_from(p, d, path=3), lam, off_in)
        a, b = 300, 316
        band = {}
        for path in (1, 2):
            band[path] = np.zeros_like(full)
            ctx.render_pixelwise(fg_params_from(p, d, path=path, rows=(a, b)), lam, off_in, out=band[path])
        assert np.array_equal(full[a:b], band[1][a:b]) and np.array_equal(band[2][a:b], band[1][a:b])
