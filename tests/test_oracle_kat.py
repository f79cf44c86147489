"""Pin the CPU oracle against the third-party crates' own known-answer vectors.

The reference (joseph-wardle/film_grain) has no tests or fixtures, and its
stochastic arithmetic lives in un-vendored crates (rand 0.8.5, rand_core 0.6.4, rand_chacha
0.3.1, rand_distr 0.4.3; Cargo.lock:2478-2509).  The vectors below are the published
value-stability / reference vectors of those crates at those versions; the oracle's
restatement must reproduce each one exactly.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

L = O.lib()
INC = 11634580027462260723  # rand / rand_distr test::rng(seed) = Pcg32::new(seed, INC)


def test_xoshiro256plusplus_reference_vector():
    # rand 0.8.5 xoshiro256plusplus.rs `reference` test (= Vigna's xoshiro256plusplus.c)
    r = O.Rng()
    L.fgo_xoshiro_from_state(C.byref(r), (C.c_uint64 * 4)(1, 2, 3, 4))
    expected = [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
                9973669472204895162, 14011001112246962877, 12406186145184390807,
                15849039046786891736, 10450023813501588000]
    assert [L.fgo_next_u64(C.byref(r)) for _ in expected] == expected


def test_xoshiro_from_seed_le_and_next_u32_upper_half():
    seed = (C.c_uint8 * 32)(*([1, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0,
                               3, 0, 0, 0, 0, 0, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0]))
    r = O.Rng()
    L.fgo_xoshiro_from_seed(C.byref(r), seed)
    assert list(r.s) == [1, 2, 3, 4]
    r2 = O.Rng()
    L.fgo_xoshiro_from_seed(C.byref(r2), seed)
    for _ in range(8):
        assert L.fgo_next_u32(C.byref(r)) == (L.fgo_next_u64(C.byref(r2)) >> 32)


def test_pcg32_reference_vector():
    # PCG demo vector (rand_pcg Lcg64Xsh32 test_lcg64xsh32_reference)
    r = O.Rng()
    L.fgo_pcg32_new(C.byref(r), 42, 54)
    got = [L.fgo_next_u32(C.byref(r)) for _ in range(6)]
    assert got == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]


def test_stdrng_chacha12_construction_vector():
    # rand 0.8.5 rngs/std.rs test_stdrng_construction
    seed = (C.c_uint8 * 32)(*([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16))
    r = O.Rng()
    L.fgo_chacha12_from_seed(C.byref(r), seed)
    assert L.fgo_next_u64(C.byref(r)) == 10719222850664546238


def test_chacha_block_structure():
    # u64 = two consecutive LE u32 words; the 4-block buffer refills transparently
    seed = (C.c_uint8 * 32)(*range(32))
    a, b = O.Rng(), O.Rng()
    L.fgo_chacha12_from_seed(C.byref(a), seed)
    L.fgo_chacha12_from_seed(C.byref(b), seed)
    for _ in range(100):
        lo = L.fgo_next_u32(C.byref(b))
        hi = L.fgo_next_u32(C.byref(b))
        assert L.fgo_next_u64(C.byref(a)) == (hi << 32) | lo


def test_uniform_f32_value_stability():
    # rand 0.8.5 distributions/uniform.rs value_stability: test_samples(0f32, 1e-2f32, ...)
    r = O.Rng()
    L.fgo_pcg32_new(C.byref(r), 897, INC)
    scale = L.fgo_uniform_f32_scale(0.0, np.float32(1e-2))
    got = [L.fgo_uniform_f32_sample(C.byref(r), 0.0, scale) for _ in range(3)]
    want = [np.float32(0.0003070104), np.float32(0.0026630748), np.float32(0.00979833)]
    assert [np.float32(g) for g in got] == want


def test_poisson_value_stability_f32_both_branches():
    # rand_distr 0.4.3 tests/value_stability.rs: Poisson::new(7.0) and (27.0), f32, seed 223.
    r = O.Rng()
    L.fgo_pcg32_new(C.byref(r), 223, INC)
    assert [L.fgo_poisson_f32_sample(C.byref(r), 7.0) for _ in range(4)] == [5.0, 11.0, 6.0, 5.0]
    L.fgo_pcg32_new(C.byref(r), 223, INC)
    assert [L.fgo_poisson_f32_sample(C.byref(r), 27.0) for _ in range(4)] == [28.0, 32.0, 36.0, 36.0]


def test_poisson_f64_restatement_vectors():
    # same generic code instantiated for f64 (the type the reference uses, pixelwise.rs:81);
    # values from SURVEY.md 8(c)
    r = O.Rng()
    L.fgo_pcg32_new(C.byref(r), 223, INC)
    assert [L.fgo_poisson_f64_sample(C.byref(r), 7.0) for _ in range(4)] == [9.0, 5.0, 7.0, 6.0]
    L.fgo_pcg32_new(C.byref(r), 223, INC)
    assert [L.fgo_poisson_f64_sample(C.byref(r), 27.0) for _ in range(4)] == [28.0, 18.0, 34.0, 36.0]


def test_normal_f64_value_stability():
    # rand_distr 0.4.3 value_stability: Normal::new(0.0, 1.0), f64, seed 213
    r = O.Rng()
    L.fgo_pcg32_new(C.byref(r), 213, INC)
    got = [L.fgo_normal_f64_sample(C.byref(r), 0.0, 1.0) for _ in range(4)]
    assert got == [-0.11844188827977231, 0.7813779637772346, 0.06563993969580051, -1.1932899004186373]


def test_lognormal_f64_is_exp_of_normal():
    # rand_distr 0.4.3 LogNormal::sample = Normal::sample(rng).exp(); derived from the pinned
    # Normal vector above (math.exp and the oracle share this host's libm)
    import math
    r = O.Rng()
    L.fgo_pcg32_new(C.byref(r), 213, INC)
    got = [L.fgo_lognormal_f64_sample(C.byref(r), 0.0, 1.0) for _ in range(4)]
    want = [math.exp(v) for v in (-0.11844188827977231, 0.7813779637772346, 0.06563993969580051,
                                  -1.1932899004186373)]
    assert got == want


def test_log_gamma_lanczos():
    import math
    for x in (1.0, 2.5, 8.0, 13.0, 28.0, 441.0):
        assert abs(L.fgo_log_gamma_f64(x) - math.lgamma(x)) < 1e-9 * max(1.0, abs(math.lgamma(x)))


def test_seed_from_u64_pcg_fill_matches_pcg32_stream():
    # rand_core 0.6.4 seed_from_u64: word k is the PCG-XSH-RR output of the (k+1)-th LCG state
    MUL, INCR, M = 6364136223846793005, INC, (1 << 64) - 1
    state = 0x0123456789ABCDEF
    out = (C.c_uint8 * 32)()
    L.fgo_seed_bytes_from_u64(state, out)
    words = np.frombuffer(bytes(out), "<u4")
    for k in range(8):
        state = (state * MUL + INCR) & M
        xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        assert int(words[k]) == ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF


def test_reference_seeding_restatement_vectors():
    # surveyor-computed restatement vectors for src/rng.rs (seed 5489)
    s = 5489
    assert L.fgo_mix(s, L.fgo_stream_const(O.STREAM_OFFSET)) == 0x7403ef94804c0e9b
    cell, pix = L.fgo_stream_const(O.STREAM_CELL), L.fgo_stream_const(O.STREAM_PIXEL)
    assert L.fgo_mix3(s, cell, 0, 0) == 0x2bbaf50b9f534851
    assert L.fgo_mix3(s, cell, 1, 2) == 0x61cee92ea2cb2af1
    assert L.fgo_mix3(s, cell, -1, -1) == 0xfc0fcc352de4d27e
    assert L.fgo_mix3(s, cell, 38399, 21599) == 0xf565ca218ef14625
    assert L.fgo_mix3(s, pix, 0, 0) == 0x01016d7cf91e22aa
    assert L.fgo_mix3(s, pix, 1, 2) == 0xf027c0dca51626e5
    r = O.Rng()
    L.fgo_cell_rng(C.byref(r), s, 0, 0)
    assert list(r.s) == [0x1a696627c804d365, 0x97dd2c14503f9f52, 0x900afdf30af992bb, 0xb57a0e7f2934ae27]
    assert [L.fgo_next_u64(C.byref(r)) for _ in range(3)] == [0x6de202e88e6cc51f, 0x8998f2f549df6084,
                                                             0xe8c3514a0e9d1dc0]
    L.fgo_cell_rng(C.byref(r), s, 1, 2)
    assert list(r.s) == [0x687ffd717a138478, 0x3d09a48f2b728bd2, 0x57a45924afbf2299, 0xb4f1c8d53b678baf]
    assert L.fgo_next_u64(C.byref(r)) == 0x8bdabaf98da23d5b
    L.fgo_small_rng_seed_from_u64_variant_b(C.byref(r), L.fgo_mix3(s, cell, 0, 0))
    assert list(r.s) == [0x96aeaef080a0ed7c, 0x7eec8592d7c95d21, 0xa515e94978b9df6d, 0x47d650c042f8ddd8]
    assert L.fgo_next_u64(C.byref(r)) == 0x6f107bd62b102ffb


def test_gen_cell_restatement_vectors():
    # delta=0.1f, lambda*delta^2 = 0.7f, const radius
    p = O.make_params(radius=0.1, n_samples=1, seed=5489)
    d, _, _ = O.derive_common(p, 4, 4)
    assert np.float32(d.delta) == np.float32(0.1) and np.float32(d.rm) == np.float32(0.1)
    lam = np.float32(0.7) / np.float32(d.delta) / np.float32(d.delta)
    # make sure lam*delta*delta reproduces 0.7f exactly, else search neighbours
    for ulp in range(-4, 5):
        cand = np.nextafter(lam, np.float32(np.inf if ulp > 0 else -np.inf)) if ulp else lam
        for _ in range(abs(ulp) - 1):
            cand = np.nextafter(cand, np.float32(np.inf if ulp > 0 else -np.inf))
        if np.float32(np.float32(cand * np.float32(d.delta)) * np.float32(d.delta)) == np.float32(0.7):
            lam = cand
            break
    else:
        pytest.skip("no f32 lambda with lambda*delta*delta == 0.7f")
    q, cx, cy, r = O.gen_cell(p, d, O.STREAM_CELL, 1, 2, float(lam))
    assert q == 1
    assert (cx[0], cy[0]) == (np.float32(0.184573233127594), np.float32(0.2925030291080475))
    q, cx, cy, r = O.gen_cell(p, d, O.STREAM_CELL, -1, -1, float(lam))
    assert q == 1
    assert (cx[0], cy[0]) == (np.float32(-0.06866180896759033), np.float32(-0.0403667688369751))


def test_ziggurat_table_spot_values():
    # literals of rand_distr 0.4.3 ziggurat_tables.rs
    tab = open(O._HERE + "/zig_tables.h").read()
    x = [float.fromhex(t) for t in tab.split("FGO_ZIG_NORM_X[257] = {")[1].split("}")[0].replace("\n", "").split(",")]
    f = [float.fromhex(t) for t in tab.split("FGO_ZIG_NORM_F[257] = {")[1].split("}")[0].replace("\n", "").split(",")]
    assert len(x) == 257 and len(f) == 257
    assert x[0] == 3.910757959537090045 and x[1] == 3.654152885361008796
    assert x[2] == 3.449278298560964462 and x[255] == 0.215241895913273806 and x[256] == 0.0
    assert f[0] == 0.000477467764586655 and f[1] == 0.001260285930498598
    assert f[255] == 0.977101701282731328 and f[256] == 1.0
    assert all(x[i] > x[i + 1] for i in range(256)) and all(f[i] < f[i + 1] for i in range(256))


def test_small_rng_fast_seeding_equals_byte_path():
    rs = np.random.default_rng(1)
    for h in [0, 1, 2**64 - 1] + [int(v) for v in rs.integers(0, 2**63, 200)]:
        out = (C.c_uint8 * 32)()
        L.fgo_seed_bytes_from_u64(h, out)
        a, b = O.Rng(), O.Rng()
        L.fgo_xoshiro_from_seed(C.byref(a), out)
        L.fgo_small_rng_seed_from_u64(C.byref(b), h)
        assert list(a.s) == list(b.s)
