"""Host-side mirror (film_grain_b200/host/film_grain.cpp through include/fg_host.h) against the
oracle's restatement of the same reference functions: ParamsBuilder::build validation
(src/params.rs:141-180, 223-261), derive_common (src/model.rs:181-226), make_offsets
(src/rng.rs:9-24), choose_algorithm (src/choose.rs:4-26), normalize_plane + lambda_plane
(src/model.rs:228-265).  The two are independent implementations (C++ product code vs C checker)."""
import numpy as np
import pytest

from film_grain_b200 import build

build.build()

from film_grain_b200 import host as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    dict(radius_mean=0.1, n_samples=64),
    dict(radius_mean=0.05, n_samples=64, zoom=4.0),
    dict(radius_mean=0.12, n_samples=1024),
    dict(radius_mean=0.5, n_samples=128),
    dict(radius_mean=0.1, n_samples=256, algo=H.Algo.Pixel),
    dict(radius_mean=0.2, radius_dist=H.RadiusDist.Lognorm, radius_stddev=0.1, n_samples=40),
    dict(radius_mean=0.3, radius_dist=H.RadiusDist.Lognorm, radius_stddev=0.0, n_samples=8),
    dict(radius_mean=0.1, n_samples=8, max_radius=("absolute", 0.3), cell_delta=0.07),
    dict(radius_mean=2.0, n_samples=8, zoom=1.37, size=(77, None)),
    dict(radius_mean=0.1, n_samples=0, size=(40, 31), seed=2**63 + 12345, sigma_px=2.5),
]


def _oracle_params(kw):
    return O.make_params(radius=kw["radius_mean"], radius_dist=int(kw.get("radius_dist", 0)),
                         radius_stddev=kw.get("radius_stddev", 0.0), zoom=kw.get("zoom", 1.0),
                         sigma_px=kw.get("sigma_px", 0.8), n_samples=kw.get("n_samples", 32),
                         algo=int(kw.get("algo", 0)), max_radius=kw.get("max_radius", ("quantile", 0.999)),
                         cell_delta=kw.get("cell_delta"), size=kw.get("size"), seed=kw.get("seed", 5489))


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
@pytest.mark.parametrize("size", [(512, 512), (97, 33)])
def test_derive_common_matches_oracle(kw, size):
    p = H.ParamsBuilder(**kw).build()
    d = H.derive_common(p, size)
    po = _oracle_params(kw)
    do, off, off_in = O.derive_common(po, *size)
    assert (d.input_width, d.input_height, d.output_width, d.output_height) == \
           (do.input_width, do.input_height, do.output_width, do.output_height)
    assert np.float32(d.delta) == np.float32(do.delta)
    assert np.float32(d.rm) == np.float32(do.rm)
    assert np.float32(d.inv_e_pi_r2) == np.float32(do.inv_e_pi_r2)
    assert np.array_equal(d.offsets.view(np.uint32), off.view(np.uint32))
    assert np.array_equal(d.offsets_input.view(np.uint32), off_in.view(np.uint32))
    assert int(d.algorithm) == O.choose_algorithm(po, do)
    b = d.block
    assert (b.in_w, b.in_h, b.out_w, b.out_h, b.n_samples) == (size[0], size[1], do.output_width, do.output_height, po.n_samples)
    assert b.seed == po.seed and b.dist_kind == po.radius_dist and b.has_log == (1 if po.radius_dist == 1 else 0)
    assert b.radius_log_mu == float(np.float32(po.radius_log_mu)) and b.radius_log_sigma == float(np.float32(po.radius_log_sigma))


def test_choose_algorithm_on_baseline_configs():
    """C1 Auto->Pixel; C2 Auto->Grain (N>96, so the bench forces --algo pixel); C3 Auto->Grain;
    C4 Auto->Pixel; C5 Pixel for N<=64 and Grain above."""
    def auto(**kw):
        p = H.ParamsBuilder(**kw).build()
        return H.derive_common(p, (64, 64)).algorithm
    assert auto(radius_mean=0.1, n_samples=64) == H.Algo.Pixel
    assert auto(radius_mean=0.1, n_samples=256) == H.Algo.Grain
    assert auto(radius_mean=0.5, n_samples=128) == H.Algo.Grain
    assert auto(radius_mean=0.05, n_samples=64, zoom=4.0) == H.Algo.Pixel
    for n, want in [(16, H.Algo.Pixel), (64, H.Algo.Pixel), (65, H.Algo.Grain), (96, H.Algo.Grain), (4096, H.Algo.Grain)]:
        assert auto(radius_mean=0.12, n_samples=n) == want
    assert auto(radius_mean=0.1, n_samples=256, algo=H.Algo.Pixel) == H.Algo.Pixel


@pytest.mark.parametrize("kw,field", [
    (dict(radius_mean=0.0), "radius"), (dict(radius_mean=float("nan")), "radius"),
    (dict(radius_stddev=-1.0), "radius-stddev"), (dict(zoom=0.0), "zoom"), (dict(sigma_px=-0.1), "sigma"),
    (dict(max_radius=("quantile", 1.0)), "max-radius"), (dict(max_radius=("absolute", 0.0)), "max-radius"),
    (dict(cell_delta=0.0), "cell"), (dict(size=(0, None)), "size"), (dict(size=(10, 0)), "size"),
])
def test_params_validation_errors(kw, field):
    with pytest.raises(H.ParamsError) as e:
        H.ParamsBuilder(**kw).build()
    assert f"{field}:" in str(e.value)
    with pytest.raises(O.OracleError) as eo:
        O.make_params(radius=kw.get("radius_mean", 0.1), radius_stddev=kw.get("radius_stddev", 0.0), zoom=kw.get("zoom", 1.0),
                      sigma_px=kw.get("sigma_px", 0.8), max_radius=kw.get("max_radius", ("quantile", 0.999)),
                      cell_delta=kw.get("cell_delta"), size=kw.get("size"))
    assert str(eo.value).split(":")[0] == field


def test_default_cell_delta_quirks():
    # r in {0.05,0.1,0.2,0.5} give delta == r bit-exactly; 0.12 -> 1/9; r >= 1 -> 1
    for r in (0.05, 0.1, 0.2, 0.5):
        d = H.derive_common(H.ParamsBuilder(radius_mean=r).build(), (8, 8))
        assert np.float32(d.delta) == np.float32(r)
    assert np.float32(H.derive_common(H.ParamsBuilder(radius_mean=0.12).build(), (8, 8)).delta) == np.float32(1.0) / np.float32(9.0)
    assert H.derive_common(H.ParamsBuilder(radius_mean=3.0).build(), (8, 8)).delta == 1.0


def test_lambda_plane_matches_oracle():
    rng = np.random.default_rng(0)
    plane = rng.random((37, 53), dtype=np.float32)
    plane[0, :5] = [0.0, 1.0, 1.0 - 1e-7, 0.5, 254.0 / 255.0]
    for r in (0.05, 0.1, 0.5):
        d = H.derive_common(H.ParamsBuilder(radius_mean=r).build(), (53, 37))
        got = H.lambda_plane(plane, d.inv_e_pi_r2)
        want = O.lambda_plane(O.normalize_plane(plane), d.inv_e_pi_r2)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    big = plane * 3.0  # exercises normalize_plane's scaling branch
    d = H.derive_common(H.ParamsBuilder(radius_mean=0.1).build(), (53, 37))
    assert np.array_equal(H.lambda_plane(big, d.inv_e_pi_r2), O.lambda_plane(O.normalize_plane(big), d.inv_e_pi_r2))


def test_empty_input_is_an_error():
    p = H.ParamsBuilder().build()
    with pytest.raises(H.RenderError):
        H.derive_common(p, (0, 10))


def test_restated_logf_matches_the_platform_libm():
    """csrc/fg_logf.h (the logf the fused luma path runs on the device) vs this host's libm logf:
    the Rust host's f32::ln is that libm call (src/model.rs:261).  tools/check_logf.c is the
    exhaustive version (all positive normal floats, 0 differences on glibc 2.39)."""
    import ctypes as C
    from film_grain_b200 import _lib
    L = _lib.load()
    libm = C.CDLL("libm.so.6")
    libm.logf.restype = C.c_float
    libm.logf.argtypes = [C.c_float]
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.random(20000, dtype=np.float32), np.float32(1) - rng.random(5000, dtype=np.float32) * np.float32(1e-4),
                        np.float32(1e-6) * (1 + rng.random(2000, dtype=np.float32)),
                        (np.arange(1, 256, dtype=np.float32) / np.float32(255)),
                        np.array([1e-6, 1.0, 0.5, 2.0, 1e-30, 3e38, 1.0000001, 0.99999994], np.float32)]).astype(np.float32)
    x = x[x > 0]
    out = np.empty_like(x)
    L.fgh_logf_restated(C.c_void_p(x.ctypes.data), x.size, C.c_void_p(out.ctypes.data))
    ref = np.array([libm.logf(float(v)) for v in x], np.float32)
    assert np.array_equal(ref.view(np.uint32), out.view(np.uint32))
