// fg_rng.cuh -- device-side, bit-exact restatement of the random arithmetic the reference's
// integrators run on (src/rng.rs + rand 0.8.5 / rand_core 0.6.4 / rand_distr 0.4.3).
//
// Everything here must reproduce the CPU realisation bit for bit, so:
//   * every f32/f64 product-sum uses explicit round-to-nearest intrinsics (__fmul_rn,
//     __fadd_rn, __dmul_rn, ...) which nvcc never contracts into an FMA (Rust never fuses);
//   * divisions are IEEE (__fdiv_rn), never reciprocal multiplies;
//   * integer hashing follows src/rng.rs:36-52 and rand_core's PCG32 seed fill.
// The only non-bit-exact ingredients are the f64 transcendentals (exp/log/tan), which are
// <= 1-2 ulp from a correctly rounded libm; DESIGN.md states the residual.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fg {

// 64-bit rotate as two 32-bit funnel shifts (k is a compile-time constant at every call site)
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (k & 32) { const uint32_t t = lo; lo = hi; hi = t; }
    const uint32_t nhi = __funnelshift_l(lo, hi, k & 31), nlo = __funnelshift_l(hi, lo, k & 31);
    return ((uint64_t)nhi << 32) | nlo;
}

// src/rng.rs:46-52
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// first half of mix3 (src/rng.rs:40-42): depends on the column index only
__device__ __forceinline__ uint64_t mix3_col(uint64_t seed_xor_stream, int32_t a) {
    return splitmix64(rotl64(seed_xor_stream + (uint64_t)(int64_t)a, 17));
}
// second half (src/rng.rs:43)
__device__ __forceinline__ uint64_t mix3_row(uint64_t col_hash, int32_t b) {
    return splitmix64(rotl64(col_hash + (uint64_t)(int64_t)b, 41));
}

#define FG_PCG_MUL 6364136223846793005ULL
#define FG_PCG_INC 11634580027462260723ULL

// PCG-XSH-RR output of one LCG state (rand_core 0.6.4 seed_from_u64's inner pcg32())
__device__ __forceinline__ uint32_t pcg_out(uint64_t state) {
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    return __funnelshift_r(xorshifted, xorshifted, rot); // rotate_right
}

// Jump-ahead constants of the LCG: state_k = A_k * state_0 + C_k (k = 1..8), so the eight
// seed words are independent multiply-adds instead of a serial chain.
__host__ __device__ constexpr uint64_t pcg_jump_a(int k) {
    uint64_t a = 1;
    for (int i = 0; i < k; ++i) a = a * FG_PCG_MUL;
    return a;
}
__host__ __device__ constexpr uint64_t pcg_jump_c(int k) {
    uint64_t c = 0;
    for (int i = 0; i < k; ++i) c = c * FG_PCG_MUL + FG_PCG_INC;
    return c;
}
// seed word W (0..7) of SeedableRng::seed_from_u64(h)
template <int W>
__device__ __forceinline__ uint32_t pcg_seed_word(uint64_t h) {
    constexpr uint64_t A = pcg_jump_a(W + 1);
    constexpr uint64_t Cc = pcg_jump_c(W + 1);
    return pcg_out(A * h + Cc);
}
template <int K> // xoshiro state word s[K] = LE u64 of seed words 2K, 2K+1
__device__ __forceinline__ uint64_t pcg_word_pair(uint64_t h) {
    return (uint64_t)pcg_seed_word<2 * K>(h) | ((uint64_t)pcg_seed_word<2 * K + 1>(h) << 32);
}

struct Xoshiro { uint64_t s0, s1, s2, s3; };

// xoshiro256++'s own SplitMix64 seeding (rand >= 0.9 forwards SmallRng::seed_from_u64 here)
__device__ __forceinline__ void seed_splitmix(Xoshiro& r, uint64_t state) {
    uint64_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        state += 0x9e3779b97f4a7c15ULL;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        out[k] = z ^ (z >> 31);
    }
    r.s0 = out[0]; r.s1 = out[1]; r.s2 = out[2]; r.s3 = out[3];
}

// SmallRng::seed_from_u64(h): rand_core default PCG32 fill -> Xoshiro256PlusPlus::from_seed
__device__ __forceinline__ void seed_small_rng(Xoshiro& r, uint64_t h, uint32_t seeding) {
    if (seeding != 0) { seed_splitmix(r, h); return; }
    r.s0 = pcg_word_pair<0>(h);
    r.s1 = pcg_word_pair<1>(h);
    r.s2 = pcg_word_pair<2>(h);
    r.s3 = pcg_word_pair<3>(h);
    if ((r.s0 | r.s1 | r.s2 | r.s3) == 0) seed_splitmix(r, 0); // from_seed's all-zero guard
}

__device__ __forceinline__ uint64_t next_u64(Xoshiro& r) {
    uint64_t result = rotl64(r.s0 + r.s3, 23) + r.s0;
    uint64_t t = r.s1 << 17;
    r.s2 ^= r.s0;
    r.s3 ^= r.s1;
    r.s1 ^= r.s2;
    r.s0 ^= r.s3;
    r.s2 ^= t;
    r.s3 = rotl64(r.s3, 45);
    return result;
}
__device__ __forceinline__ uint32_t next_u32(Xoshiro& r) { return (uint32_t)(next_u64(r) >> 32); }
// the state update alone (skipping a draw whose value is not needed)
__device__ __forceinline__ void advance(Xoshiro& r) {
    const uint64_t t = r.s1 << 17;
    r.s2 ^= r.s0;
    r.s3 ^= r.s1;
    r.s1 ^= r.s2;
    r.s0 ^= r.s3;
    r.s2 ^= t;
    r.s3 = rotl64(r.s3, 45);
}
// a generator that counts its draws: k_gen_rows keeps one seed state per cell for all colour planes and
// re-derives a plane's state after its Poisson draw by skipping that many outputs
struct XoshiroCounted : Xoshiro { uint32_t n; };
__device__ __forceinline__ uint64_t next_u64(XoshiroCounted& r) { ++r.n; return next_u64(static_cast<Xoshiro&>(r)); }

// rand Standard: f64 in [0,1) with 53 bits
template <class R>
__device__ __forceinline__ double standard_f64(R& r) {
    return __dmul_rn((double)(next_u64(r) >> 11), 1.0 / 9007199254740992.0);
}
// rand Open01 f64
__device__ __forceinline__ double open01_f64(Xoshiro& r) {
    uint64_t bits = (next_u64(r) >> 12) | 0x3FF0000000000000ULL;
    return __dsub_rn(__longlong_as_double((long long)bits), 1.0 - 2.220446049250313e-16 / 2.0);
}
// rand UniformFloat<f32>::sample with low = 0: value0_1 * scale (+ 0.0)
__device__ __forceinline__ float uniform_f32(Xoshiro& r, float scale) {
    uint32_t bits = (next_u32(r) >> 9) | 0x3F800000u;
    float v = __fsub_rn(__uint_as_float(bits), 1.0f);
    return __fmul_rn(v, scale);
}

// rand_distr 0.4.3 utils::log_gamma (6-term Lanczos)
__device__ inline double log_gamma(double x) {
    const double coefficients[6] = {76.18009172947146,  -86.50532032941677,   24.01409824083091,
                                    -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
    double tmp = __dadd_rn(x, 5.5);
    double lg = __dsub_rn(__dmul_rn(__dadd_rn(x, 0.5), log(tmp)), tmp);
    double a = 1.000000000190015;
    double denom = x;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        denom = __dadd_rn(denom, 1.0);
        a = __dadd_rn(a, __ddiv_rn(coefficients[k], denom));
    }
    return __dadd_rn(lg, log(__ddiv_rn(__dmul_rn(2.5066282746310005, a), x)));
}

// f64 -> u32 like Rust's `as u32` (saturating, NaN -> 0)
__device__ __forceinline__ uint32_t sat_u32(double v) {
    if (!(v > 0.0)) return 0u;
    if (v >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

// rand_distr 0.4.3 Poisson<f64>::new(lambda).sample(rng) as u32; lambda > 0.
template <class R>
__device__ inline uint32_t poisson_f64(R& r, double lambda) {
    // Poisson::new rejects lambda <= 0 / NaN (the reference unwrap()s, i.e. panics); absurd means
    // would never terminate in reasonable time either.  Draw nothing instead of hanging the GPU.
    if (!(lambda > 0.0) || !(lambda < 1.0e15)) return 0u;
    if (lambda < 12.0) { // Knuth
        double exp_lambda = exp(-lambda);
        double result = 0.0, p = 1.0;
        while (p > exp_lambda) {
            p = __dmul_rn(p, standard_f64(r));
            result = __dadd_rn(result, 1.0);
        }
        return sat_u32(__dsub_rn(result, 1.0));
    }
    // rejection against a Cauchy envelope (Numerical Recipes)
    double log_lambda = log(lambda);
    double sqrt_2lambda = __dsqrt_rn(__dmul_rn(2.0, lambda));
    double magic_val = __dsub_rn(__dmul_rn(lambda, log_lambda), log_gamma(__dadd_rn(1.0, lambda)));
    double result;
    for (;;) {
        double comp_dev;
        for (;;) {
            double x = standard_f64(r);
            comp_dev = tan(__dmul_rn(3.14159265358979323846264338327950288, x));
            comp_dev = __dadd_rn(0.0, __dmul_rn(1.0, comp_dev));
            result = __dadd_rn(__dmul_rn(sqrt_2lambda, comp_dev), lambda);
            if (result >= 0.0) break;
        }
        result = floor(result);
        double e = __dsub_rn(__dsub_rn(__dmul_rn(result, log_lambda), log_gamma(__dadd_rn(1.0, result))), magic_val);
        double check = __dmul_rn(__dmul_rn(0.9, __dadd_rn(1.0, __dmul_rn(comp_dev, comp_dev))), exp(e));
        if (standard_f64(r) <= check) break;
    }
    return sat_u32(result);
}

// ziggurat tables of rand_distr 0.4.3 (257 entries each), filled once per context
// (global, read through the read-only path: lanes index different layers, which would serialise in
// the constant cache)
__device__ double kZigX[257];
__device__ double kZigF[257];
#define FG_ZIG_R 3.654152885361008796

// rand_distr 0.4.3 StandardNormal f64 (ziggurat, symmetric)
__device__ inline double standard_normal(Xoshiro& r) {
    for (;;) {
        uint64_t bits = next_u64(r);
        int i = (int)(bits & 0xff);
        double u = __dsub_rn(__longlong_as_double((long long)((bits >> 12) | 0x4000000000000000ULL)), 3.0);
        double xi = __ldg(&kZigX[i]), xi1 = __ldg(&kZigX[i + 1]);
        double x = __dmul_rn(u, xi);
        if (fabs(x) < xi1) return x;
        if (i == 0) { // tail
            double tx = 1.0, ty = 0.0;
            while (__dmul_rn(-2.0, ty) < __dmul_rn(tx, tx)) {
                double x_ = open01_f64(r);
                double y_ = open01_f64(r);
                tx = __ddiv_rn(log(x_), FG_ZIG_R);
                ty = log(y_);
            }
            return (u < 0.0) ? __dsub_rn(tx, FG_ZIG_R) : __dsub_rn(FG_ZIG_R, tx);
        }
        double f0 = __ldg(&kZigF[i]), f1 = __ldg(&kZigF[i + 1]);
        double lhs = __dadd_rn(f1, __dmul_rn(__dsub_rn(f0, f1), standard_f64(r)));
        double pdf = exp(__ddiv_rn(__dmul_rn(-x, x), 2.0));
        if (lhs < pdf) return x;
    }
}

// per-render radius model (RadiusProfile, src/model.rs:101-148)
struct RadiusModel {
    uint32_t lognorm;   // dist == Lognorm && lognormal is Some
    float mean_linear;
    float rm;
    double mu, sigma;
};
// RadiusProfile::sample + the clamp to rm (src/pixelwise.rs:89-92, src/grainwise.rs:54-57)
__device__ __forceinline__ float radius_sample_clamped(const RadiusModel& m, Xoshiro& r) {
    float radius = m.mean_linear;
    if (m.lognorm) {
        double n = standard_normal(r);
        radius = __double2float_rn(exp(__dadd_rn(m.mu, __dmul_rn(m.sigma, n))));
    }
    if (radius > m.rm) radius = m.rm;
    return radius;
}

// saturating conversions with floor, like `x.floor() as i32` / `as isize`
__device__ __forceinline__ int32_t floor_i32(float v) { return __float2int_rd(v); }
__device__ __forceinline__ long long floor_i64(float v) { return __float2ll_rd(v); }

} // namespace fg
