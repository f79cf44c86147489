/* fg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See fg_oracle.h.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference) or the third-party crate item it restates (crate sources are not
 * vendored; versions pinned in Cargo.lock:2478-2509).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp).
 * Rust never contracts a*b+c into an FMA and evaluates f32 expressions in f32, so the
 * build flags forbid contraction and every expression below keeps the reference's
 * operand order and type.
 */
#include "fg_oracle.h"
#include "zig_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ helpers ---- */
static inline uint64_t rotl64(uint64_t x, unsigned k) { return (x << k) | (x >> (64 - k)); }
static inline uint32_t rotr32(uint32_t x, unsigned k) { k &= 31; return (x >> k) | (x << ((32 - k) & 31)); }
static inline uint32_t rotl32(uint32_t x, unsigned k) { return (x << k) | (x >> (32 - k)); }

/* Rust `as` casts from float saturate and map NaN to 0. */
static inline int32_t sat_i32_f32(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;
}
static inline int64_t sat_i64_f32(float v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0f) return INT64_MAX;
    if (v <= -9223372036854775808.0f) return INT64_MIN;
    return (int64_t)v;
}
static inline uint32_t sat_u32_f64(double v) {
    if (v != v) return 0;
    if (v <= 0.0) return 0;
    if (v >= 4294967295.0) return UINT32_MAX;
    return (uint32_t)v;
}
static inline float clampf(float v, float lo, float hi) { /* f32::clamp */
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
static inline float maxf_rust(float a, float b) { /* f32::max: NaN-ignoring */
    if (a != a) return b;
    if (b != b) return a;
    return a > b ? a : b;
}
static inline float minf_rust(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}
static void set_msg(char* msg, size_t len, const char* text) {
    if (msg && len) { snprintf(msg, len, "%s", text); }
}

/* ------------------------------------------------------------- src/rng.rs ---- */
static const uint64_t OFFSET_STREAM = 0x9E3779B97F4A7C15ULL; /* rng.rs:5 */
static const uint64_t CELL_STREAM = 0xA24B1C30BEBCCF59ULL;   /* rng.rs:6 */
static const uint64_t PIXEL_STREAM = 0x6935FA5C55F65F1BULL;  /* rng.rs:7 */

uint64_t fgo_stream_const(int which) {
    return which == 0 ? OFFSET_STREAM : which == 1 ? CELL_STREAM : PIXEL_STREAM;
}

uint64_t fgo_splitmix64(uint64_t x) { /* rng.rs:46-52 */
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t fgo_mix(uint64_t seed, uint64_t stream) { return fgo_splitmix64(seed ^ stream); } /* :36-38 */
uint64_t fgo_mix3(uint64_t seed, uint64_t stream, int64_t a, int64_t b) { /* :40-44 */
    uint64_t state = seed ^ stream;
    state = fgo_splitmix64(rotl64(state + (uint64_t)a, 17));
    return fgo_splitmix64(rotl64(state + (uint64_t)b, 41));
}

/* -------------------------------- rand_core 0.6.4 SeedableRng::seed_from_u64 ---- */
void fgo_seed_bytes_from_u64(uint64_t state, uint8_t out[32]) {
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    for (int w = 0; w < 8; ++w) {
        state = state * MUL + INC; /* advance first */
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        uint32_t x = rotr32(xorshifted, rot);
        out[4 * w + 0] = (uint8_t)x;        /* to_le_bytes */
        out[4 * w + 1] = (uint8_t)(x >> 8);
        out[4 * w + 2] = (uint8_t)(x >> 16);
        out[4 * w + 3] = (uint8_t)(x >> 24);
    }
}

/* -------------------------------------- rand 0.8.5 Xoshiro256PlusPlus (SmallRng) ---- */
static void xoshiro_seed_splitmix(fgo_rng* r, uint64_t state) { /* xoshiro's own seed_from_u64 */
    const uint64_t PHI = 0x9e3779b97f4a7c15ULL;
    for (int k = 0; k < 4; ++k) {
        state += PHI;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        z = z ^ (z >> 31);
        r->s[k] = z;
    }
}
void fgo_xoshiro_from_seed(fgo_rng* r, const uint8_t seed[32]) {
    r->kind = FGO_RNG_XOSHIRO256PP;
    int allzero = 1;
    for (int k = 0; k < 32; ++k) allzero &= (seed[k] == 0);
    if (allzero) { xoshiro_seed_splitmix(r, 0); return; }
    for (int k = 0; k < 4; ++k) {
        uint64_t v = 0;
        for (int b = 7; b >= 0; --b) v = (v << 8) | seed[8 * k + b]; /* from_le_bytes */
        r->s[k] = v;
    }
}
void fgo_xoshiro_from_state(fgo_rng* r, const uint64_t s[4]) {
    r->kind = FGO_RNG_XOSHIRO256PP;
    memcpy(r->s, s, 32);
}
void fgo_small_rng_seed_from_u64(fgo_rng* r, uint64_t state) { /* variant A: rand_core default */
    /* same as fgo_seed_bytes_from_u64 + fgo_xoshiro_from_seed, without the byte round trip
     * (tests/test_oracle_kat.py checks the two agree) */
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    uint32_t w[8];
    for (int k = 0; k < 8; ++k) {
        state = state * MUL + INC;
        w[k] = rotr32((uint32_t)(((state >> 18) ^ state) >> 27), (uint32_t)(state >> 59));
    }
    r->kind = FGO_RNG_XOSHIRO256PP;
    for (int k = 0; k < 4; ++k) r->s[k] = (uint64_t)w[2 * k] | ((uint64_t)w[2 * k + 1] << 32);
    if ((r->s[0] | r->s[1] | r->s[2] | r->s[3]) == 0) xoshiro_seed_splitmix(r, 0);
}
void fgo_small_rng_seed_from_u64_variant_b(fgo_rng* r, uint64_t state) {
    r->kind = FGO_RNG_XOSHIRO256PP;
    xoshiro_seed_splitmix(r, state);
}
static inline uint64_t xoshiro_next(fgo_rng* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* ------------------------------------------------- rand_pcg Lcg64Xsh32 (tests) ---- */
void fgo_pcg32_new(fgo_rng* r, uint64_t state, uint64_t stream) {
    memset(r, 0, sizeof *r);
    r->kind = FGO_RNG_PCG32;
    r->pcg_inc = (stream << 1) | 1;
    r->pcg_state = state + r->pcg_inc;
    r->pcg_state = r->pcg_state * 6364136223846793005ULL + r->pcg_inc;
}
static inline uint32_t pcg32_next(fgo_rng* r) {
    uint64_t state = r->pcg_state;
    r->pcg_state = state * 6364136223846793005ULL + r->pcg_inc;
    uint32_t rot = (uint32_t)(state >> 59);
    uint32_t xsh = (uint32_t)(((state >> 18) ^ state) >> 27);
    return rotr32(xsh, rot);
}

/* --------------------------- rand_chacha 0.3.1 ChaCha12Rng over BlockRng (StdRng) ---- */
#define CC_QR(a, b, c, d)                                     \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
static void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    for (int round = 0; round < 6; ++round) { /* 12 rounds = 6 double rounds */
        CC_QR(x[0], x[4], x[8], x[12]) CC_QR(x[1], x[5], x[9], x[13])
        CC_QR(x[2], x[6], x[10], x[14]) CC_QR(x[3], x[7], x[11], x[15])
        CC_QR(x[0], x[5], x[10], x[15]) CC_QR(x[1], x[6], x[11], x[12])
        CC_QR(x[2], x[7], x[8], x[13]) CC_QR(x[3], x[4], x[9], x[14])
    }
    for (int k = 0; k < 16; ++k) out[k] = x[k] + in[k];
}
static void chacha_refill(fgo_rng* r) { /* 4 consecutive blocks per BlockRng refill */
    for (int b = 0; b < 4; ++b) chacha12_block(r->cc_key, r->cc_counter + (uint64_t)b, r->cc_buf + 16 * b);
    r->cc_counter += 4;
}
void fgo_chacha12_from_seed(fgo_rng* r, const uint8_t seed[32]) {
    memset(r, 0, sizeof *r);
    r->kind = FGO_RNG_CHACHA12;
    for (int k = 0; k < 8; ++k)
        r->cc_key[k] = (uint32_t)seed[4 * k] | ((uint32_t)seed[4 * k + 1] << 8) |
                       ((uint32_t)seed[4 * k + 2] << 16) | ((uint32_t)seed[4 * k + 3] << 24);
    r->cc_counter = 0;
    r->cc_index = 64; /* empty buffer */
}
void fgo_std_rng_seed_from_u64(fgo_rng* r, uint64_t state) {
    uint8_t seed[32];
    fgo_seed_bytes_from_u64(state, seed);
    fgo_chacha12_from_seed(r, seed);
}
static uint64_t chacha_next_u64(fgo_rng* r) { /* rand_core BlockRng::next_u64 */
    const int len = 64;
    int index = r->cc_index;
    if (index < len - 1) {
        r->cc_index += 2;
        return ((uint64_t)r->cc_buf[index + 1] << 32) | r->cc_buf[index];
    } else if (index >= len) {
        chacha_refill(r);
        r->cc_index = 2;
        return ((uint64_t)r->cc_buf[1] << 32) | r->cc_buf[0];
    } else {
        uint64_t x = r->cc_buf[len - 1];
        chacha_refill(r);
        r->cc_index = 1;
        uint64_t y = r->cc_buf[0];
        return (y << 32) | x;
    }
}
static uint32_t chacha_next_u32(fgo_rng* r) {
    if (r->cc_index >= 64) { chacha_refill(r); r->cc_index = 0; }
    return r->cc_buf[r->cc_index++];
}

uint64_t fgo_next_u64(fgo_rng* r) {
    if (__builtin_expect(r->kind == FGO_RNG_XOSHIRO256PP, 1)) return xoshiro_next(r);
    switch (r->kind) {
    case FGO_RNG_XOSHIRO256PP: return xoshiro_next(r);
    case FGO_RNG_CHACHA12: return chacha_next_u64(r);
    default: { uint64_t x = pcg32_next(r); uint64_t y = pcg32_next(r); return (y << 32) | x; }
    }
}
uint32_t fgo_next_u32(fgo_rng* r) {
    if (__builtin_expect(r->kind == FGO_RNG_XOSHIRO256PP, 1)) return (uint32_t)(xoshiro_next(r) >> 32);
    switch (r->kind) {
    case FGO_RNG_XOSHIRO256PP: return (uint32_t)(xoshiro_next(r) >> 32); /* upper half */
    case FGO_RNG_CHACHA12: return chacha_next_u32(r);
    default: return pcg32_next(r);
    }
}

static int g_seeding_variant = 1;
void fgo_set_seeding_variant(int v) { g_seeding_variant = (v == 2) ? 2 : 1; }

void fgo_cell_rng(fgo_rng* r, uint64_t seed, int32_t i, int32_t j) { /* rng.rs:26-29 */
    uint64_t hashed = fgo_mix3(seed, CELL_STREAM, (int64_t)i, (int64_t)j);
    if (g_seeding_variant == 2) fgo_small_rng_seed_from_u64_variant_b(r, hashed);
    else fgo_small_rng_seed_from_u64(r, hashed);
}
void fgo_pixel_rng(fgo_rng* r, uint64_t seed, int32_t i, int32_t j) { /* rng.rs:31-34 */
    uint64_t hashed = fgo_mix3(seed, PIXEL_STREAM, (int64_t)i, (int64_t)j);
    if (g_seeding_variant == 2) fgo_small_rng_seed_from_u64_variant_b(r, hashed);
    else fgo_small_rng_seed_from_u64(r, hashed);
}

/* ----------------------------------------------- rand 0.8.5 Standard / Uniform ---- */
double fgo_standard_f64(fgo_rng* r) { return (double)(fgo_next_u64(r) >> 11) * (1.0 / 9007199254740992.0); }
float fgo_standard_f32(fgo_rng* r) { return (float)(fgo_next_u32(r) >> 8) * (1.0f / 16777216.0f); }
double fgo_open01_f64(fgo_rng* r) {
    uint64_t bits = (fgo_next_u64(r) >> 12) | 0x3FF0000000000000ULL;
    double v; memcpy(&v, &bits, 8);
    return v - (1.0 - 2.220446049250313e-16 / 2.0);
}
float fgo_uniform_f32_scale(float low, float high) { /* UniformFloat<f32>::new */
    uint32_t mr = (0xFFFFFFFFu >> 9) | 0x3F800000u;
    float max_rand; memcpy(&max_rand, &mr, 4);
    max_rand = max_rand - 1.0f;
    float scale = high - low;
    for (;;) {
        float t = scale * max_rand;
        t = t + low;
        if (!(t >= high)) break;
        uint32_t b; memcpy(&b, &scale, 4); b -= 1; memcpy(&scale, &b, 4); /* decrease by one ulp */
    }
    return scale;
}
float fgo_uniform_f32_sample(fgo_rng* r, float low, float scale) { /* UniformFloat<f32>::sample */
    uint32_t bits = (fgo_next_u32(r) >> 9) | 0x3F800000u;
    float value1_2; memcpy(&value1_2, &bits, 4);
    float value0_1 = value1_2 - 1.0f;
    float t = value0_1 * scale; /* separate multiply and add -- no FMA */
    return t + low;
}

/* -------------------------------------------- rand_distr 0.4.3 Poisson / utils ---- */
#define DEFINE_POISSON(T, SUF, LN, EXP, SQRT, TAN, FLOOR, GEN, PI_CONST)                          \
    static T log_gamma_##SUF(T x) {                                                               \
        const T coefficients[6] = {(T)76.18009172947146, (T)-86.50532032941677,                   \
                                   (T)24.01409824083091, (T)-1.231739572450155,                   \
                                   (T)0.1208650973866179e-2, (T)-0.5395239384953e-5};             \
        T tmp = x + (T)5.5;                                                                       \
        T lg = (x + (T)0.5) * LN(tmp) - tmp;                                                      \
        T a = (T)1.000000000190015;                                                               \
        T denom = x;                                                                              \
        for (int k = 0; k < 6; ++k) { denom = denom + (T)1.0; a = a + (coefficients[k] / denom); }\
        return lg + LN((T)2.5066282746310005 * a / x);                                            \
    }                                                                                             \
    static T poisson_sample_##SUF(fgo_rng* rng, T lambda) {                                       \
        T exp_lambda = EXP(-lambda);                                                              \
        if (lambda < (T)12.0) { /* Knuth */                                                       \
            T result = (T)0.0, p = (T)1.0;                                                        \
            while (p > exp_lambda) { p = p * GEN(rng); result = result + (T)1.0; }                \
            return result - (T)1.0;                                                               \
        }                                                                                         \
        /* Poisson::new's remaining fields; only this branch reads them */                        \
        T log_lambda = LN(lambda);                                                                \
        T sqrt_2lambda = SQRT((T)2.0 * lambda);                                                   \
        T magic_val = lambda * log_lambda - log_gamma_##SUF((T)1.0 + lambda);                     \
        T result;                                                                                 \
        for (;;) {                                                                                \
            T comp_dev;                                                                           \
            for (;;) {                                                                            \
                T x = GEN(rng);              /* Cauchy::new(0,1).sample */                       \
                comp_dev = TAN(PI_CONST * x);                                                     \
                comp_dev = (T)0.0 + (T)1.0 * comp_dev;                                            \
                result = sqrt_2lambda * comp_dev + lambda;                                        \
                if (result >= (T)0.0) break;                                                      \
            }                                                                                     \
            result = FLOOR(result);                                                               \
            T check = (T)0.9 * ((T)1.0 + comp_dev * comp_dev) *                                   \
                      EXP(result * log_lambda - log_gamma_##SUF((T)1.0 + result) - magic_val);    \
            if (GEN(rng) <= check) break;                                                         \
        }                                                                                         \
        return result;                                                                            \
    }
DEFINE_POISSON(double, f64, log, exp, sqrt, tan, floor, fgo_standard_f64, 3.14159265358979323846264338327950288)
DEFINE_POISSON(float, f32, logf, expf, sqrtf, tanf, floorf, fgo_standard_f32, 3.14159265358979323846264338327950288f)

double fgo_log_gamma_f64(double x) { return log_gamma_f64(x); }
double fgo_poisson_f64_sample(fgo_rng* r, double lambda) { return poisson_sample_f64(r, lambda); }
float fgo_poisson_f32_sample(fgo_rng* r, float lambda) { return poisson_sample_f32(r, lambda); }

/* --------------------------------- rand_distr 0.4.3 StandardNormal / Normal / LogNormal ---- */
static double zig_pdf(double x) { return exp(-x * x / 2.0); }
static double zig_zero_case(fgo_rng* rng, double u) {
    double x = 1.0, y = 0.0;
    while (-2.0 * y < x * x) {
        double x_ = fgo_open01_f64(rng);
        double y_ = fgo_open01_f64(rng);
        x = log(x_) / FGO_ZIG_NORM_R;
        y = log(y_);
    }
    return (u < 0.0) ? x - FGO_ZIG_NORM_R : FGO_ZIG_NORM_R - x;
}
double fgo_standard_normal_f64(fgo_rng* rng) {
    for (;;) {
        uint64_t bits = fgo_next_u64(rng);
        size_t i = (size_t)(bits & 0xff);
        uint64_t fb = (bits >> 12) | 0x4000000000000000ULL; /* exponent 1: [2,4) */
        double f; memcpy(&f, &fb, 8);
        double u = f - 3.0;
        double x = u * FGO_ZIG_NORM_X[i];
        double test_x = fabs(x);
        if (test_x < FGO_ZIG_NORM_X[i + 1]) return x;
        if (i == 0) return zig_zero_case(rng, u);
        if (FGO_ZIG_NORM_F[i + 1] + (FGO_ZIG_NORM_F[i] - FGO_ZIG_NORM_F[i + 1]) * fgo_standard_f64(rng) < zig_pdf(x))
            return x;
    }
}
double fgo_normal_f64_sample(fgo_rng* r, double mean, double std_dev) {
    double z = fgo_standard_normal_f64(r);
    return mean + std_dev * z;
}
double fgo_lognormal_f64_sample(fgo_rng* r, double mu, double sigma) {
    return exp(fgo_normal_f64_sample(r, mu, sigma));
}

/* statrs 0.16.1 Normal::inverse_cdf stand-in (host only, model.rs:159-161).  statrs's
 * rational erfc_inv is not restated; this is Acklam's approximation polished by two
 * Halley steps on erfc, accurate to ~1 ulp of f64.  Its only consumer rounds
 * exp(mu+sigma*z) to f32, so the two agree unless that value sits within ~1e-16
 * (relative) of an f32 rounding boundary. */
double fgo_norm_inv_cdf(double p) {
    if (!(p > 0.0)) return -INFINITY;
    if (!(p < 1.0)) return INFINITY;
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double x, q, r;
    if (p < 0.02425) {
        q = sqrt(-2 * log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        q = p - 0.5; r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        q = sqrt(-2 * log(1 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    for (int it = 0; it < 2; ++it) {
        double e = 0.5 * erfc(-x / sqrt(2.0)) - p;
        double u = e * sqrt(2.0 * 3.14159265358979323846) * exp(x * x / 2.0);
        x = x - u / (1.0 + x * u / 2.0);
    }
    return x;
}

/* ----------------------------------------------------------- src/params.rs ---- */
float fgo_default_cell_delta(float radius_mean) { /* params.rs:255-261 */
    if (radius_mean <= 0.0f) return 1.0f;
    float inv = maxf_rust(ceilf(1.0f / radius_mean), 1.0f);
    return 1.0f / inv;
}

int fgo_params_build(fgo_params* p, char* msg, size_t msg_len) { /* params.rs:141-180 */
    if (!isfinite(p->radius_mean)) { set_msg(msg, msg_len, "radius: value must be finite"); return -1; }
    if (p->radius_mean <= 0.0f) { set_msg(msg, msg_len, "radius: value must be greater than 0"); return -1; }
    if (!isfinite(p->radius_stddev)) { set_msg(msg, msg_len, "radius-stddev: value must be finite"); return -1; }
    if (p->radius_stddev < 0.0f) { set_msg(msg, msg_len, "radius-stddev: value must be >= 0"); return -1; }
    if (!isfinite(p->zoom)) { set_msg(msg, msg_len, "zoom: value must be finite"); return -1; }
    if (p->zoom <= 0.0f) { set_msg(msg, msg_len, "zoom: value must be greater than 0"); return -1; }
    if (!isfinite(p->sigma_px)) { set_msg(msg, msg_len, "sigma: value must be finite"); return -1; }
    if (p->sigma_px <= 0.0f) { set_msg(msg, msg_len, "sigma: value must be greater than 0"); return -1; }
    if (p->n_samples < 1) p->n_samples = 1; /* :147 */
    if (p->max_radius_kind == FGO_MAXR_ABSOLUTE) { /* :298-328 */
        if (!isfinite(p->max_radius_value)) { set_msg(msg, msg_len, "max-radius: absolute radius must be finite"); return -1; }
        if (p->max_radius_value <= 0.0f) { set_msg(msg, msg_len, "max-radius: absolute radius must be > 0"); return -1; }
    } else {
        if (!isfinite(p->max_radius_value)) { set_msg(msg, msg_len, "max-radius: quantile must be finite"); return -1; }
        if (!(0.0f < p->max_radius_value && p->max_radius_value < 1.0f)) {
            set_msg(msg, msg_len, "max-radius: quantile must lie in the open interval (0,1)"); return -1; }
    }
    if (p->has_size) { /* :344-358 */
        if (p->size_w == 0) { set_msg(msg, msg_len, "size: output width must be > 0"); return -1; }
        if (p->has_size_h && p->size_h == 0) { set_msg(msg, msg_len, "size: output height must be > 0"); return -1; }
    }
    if (p->has_cell_delta) { /* :283-296 */
        if (!isfinite(p->cell_delta)) { set_msg(msg, msg_len, "cell: cell size must be finite"); return -1; }
        if (p->cell_delta <= 0.0f) { set_msg(msg, msg_len, "cell: cell size must be > 0"); return -1; }
    } else {
        p->has_cell_delta = 1;
        p->cell_delta = fgo_default_cell_delta(p->radius_mean);
    }
    /* derive_radius_parameters, params.rs:223-253 */
    if (p->radius_dist == FGO_DIST_CONST) {
        p->radius_stddev = 0.0f; p->has_log = 0; p->radius_log_mu = 0.0f; p->radius_log_sigma = 0.0f;
    } else {
        float mean = p->radius_mean, sd = p->radius_stddev;
        if (sd == 0.0f) {
            p->radius_stddev = 0.0f; p->has_log = 1; p->radius_log_mu = logf(mean); p->radius_log_sigma = 0.0f;
        } else {
            float variance_ratio = (sd * sd) / (mean * mean);
            float sigma_sq = logf(1.0f + variance_ratio);
            float sigma = sqrtf(sigma_sq);
            float mu = logf(mean) - 0.5f * sigma_sq;
            p->has_log = 1; p->radius_log_mu = mu; p->radius_log_sigma = sigma;
        }
    }
    return 0;
}

/* ------------------------------------------------------------ src/model.rs ---- */
static const float EPSILON_F = 1e-6f;   /* model.rs:10 */
static const float MAX_LAMBDA = 1.0e6f; /* model.rs:11 */

void fgo_make_offsets(uint64_t seed, size_t n, float sigma, float* out) { /* rng.rs:9-24 */
    if (n == 0) return;
    double sd = (double)maxf_rust(sigma, 1.1920929e-7f /* f32::EPSILON */);
    fgo_rng rng;
    fgo_std_rng_seed_from_u64(&rng, fgo_mix(seed, OFFSET_STREAM));
    for (size_t k = 0; k < n; ++k) {
        out[2 * k + 0] = (float)fgo_normal_f64_sample(&rng, 0.0, sd);
        out[2 * k + 1] = (float)fgo_normal_f64_sample(&rng, 0.0, sd);
    }
}

static float radius_quantile(const fgo_params* p, float q) { /* model.rs:150-164 */
    if (p->radius_dist == FGO_DIST_CONST) return p->radius_mean;
    double mu = p->has_log ? (double)p->radius_log_mu : (double)p->radius_mean;
    double sigma = p->has_log ? (double)p->radius_log_sigma : 0.0;
    if (sigma == 0.0) return (float)exp(mu);
    double z = fgo_norm_inv_cdf((double)q);
    return (float)exp(mu + sigma * z);
}

int fgo_derive_common(const fgo_params* p, int64_t in_w, int64_t in_h, fgo_derived* d,
                      float* offsets, float* offsets_input, char* msg, size_t msg_len) { /* model.rs:181-226 */
    if (in_w <= 0 || in_h <= 0) { set_msg(msg, msg_len, "input image is empty after ROI"); return -1; }
    int64_t ow, oh; /* resolve_output_size, model.rs:267-298 */
    if (p->has_size) {
        ow = (int64_t)p->size_w;
        if (ow == 0) { set_msg(msg, msg_len, "output width must be > 0"); return -1; }
        if (p->has_size_h) oh = (int64_t)p->size_h;
        else {
            float fw = (float)in_w, fh = (float)in_h;
            float aspect = fh / fw;
            float computed = maxf_rust(roundf((float)ow * aspect), 1.0f);
            oh = (int64_t)computed;
        }
        if (oh == 0) { set_msg(msg, msg_len, "output height must be > 0"); return -1; }
    } else {
        ow = (int64_t)maxf_rust(ceilf((float)in_w * p->zoom), 1.0f);
        oh = (int64_t)maxf_rust(ceilf((float)in_h * p->zoom), 1.0f);
    }
    if (ow == 0 || oh == 0) { set_msg(msg, msg_len, "output dimensions must be positive"); return -1; }
    float mean_sq = p->radius_mean * p->radius_mean;
    float variance = p->radius_stddev * p->radius_stddev;
    const float PI_F = 3.14159265358979323846f;
    float inv_e_pi_r2 = 1.0f / (PI_F * maxf_rust(mean_sq + variance, EPSILON_F));
    if (p->radius_dist == FGO_DIST_LOGNORM && !p->has_log) {
        set_msg(msg, msg_len, "missing log-normal mean; parameters were not derived"); return -1; }
    float rm = (p->max_radius_kind == FGO_MAXR_ABSOLUTE) ? p->max_radius_value
                                                         : radius_quantile(p, p->max_radius_value);
    rm = maxf_rust(rm, EPSILON_F);
    float delta = p->has_cell_delta ? p->cell_delta : fgo_default_cell_delta(p->radius_mean);
    delta = maxf_rust(delta, EPSILON_F);
    if (offsets) {
        fgo_make_offsets(p->seed, p->n_samples, p->sigma_px, offsets);
        if (offsets_input)
            for (size_t k = 0; k < 2 * (size_t)p->n_samples; ++k) offsets_input[k] = offsets[k] / p->zoom;
    }
    d->input_width = in_w; d->input_height = in_h; d->output_width = ow; d->output_height = oh;
    d->inv_e_pi_r2 = inv_e_pi_r2; d->rm = rm; d->delta = delta;
    d->radius_dist = p->radius_dist; d->mean_linear = p->radius_mean; d->has_log = p->has_log;
    d->log_mu = (double)p->radius_log_mu; d->log_sigma = (double)p->radius_log_sigma;
    return 0;
}

int fgo_choose_algorithm(const fgo_params* p, const fgo_derived* d) { /* choose.rs:4-26 */
    if (p->algo != FGO_ALGO_AUTO) return p->algo;
    float mean = maxf_rust(p->radius_mean, 1e-6f);
    float sigma_ratio = (mean > 0.0f) ? p->radius_stddev / mean : 0.0f;
    float rm_ratio = d->rm / mean;
    if (sigma_ratio < 0.1f && mean < 0.5f && p->n_samples <= 64) return FGO_ALGO_PIXEL;
    if (rm_ratio > 8.0f || sigma_ratio > 0.6f || p->n_samples > 96) return FGO_ALGO_GRAIN;
    if (p->n_samples <= 24 && rm_ratio < 5.0f) return FGO_ALGO_PIXEL;
    return FGO_ALGO_GRAIN;
}

float fgo_normalize_plane(const float* src, size_t n, float* dst) { /* model.rs:228-250 */
    float max_value = 0.0f;
    for (size_t k = 0; k < n; ++k) max_value = maxf_rust(max_value, src[k]);
    int already = max_value <= 1.0f + EPSILON_F;
    for (size_t k = 0; k < n; ++k) {
        float v;
        if (already) v = src[k];
        else { float denom = maxf_rust(max_value + EPSILON_F, EPSILON_F); v = src[k] / denom; }
        dst[k] = clampf(v, 0.0f, 1.0f - EPSILON_F);
    }
    return max_value;
}
void fgo_lambda_plane(const float* norm, size_t n, float inv_e_pi_r2, float* dst) { /* model.rs:252-265 */
    for (size_t k = 0; k < n; ++k) {
        float clamped = clampf(norm[k], 0.0f, 1.0f - EPSILON_F);
        float safe = maxf_rust(1.0f - clamped, EPSILON_F);
        float activity = -inv_e_pi_r2 * logf(safe);
        dst[k] = minf_rust(activity, MAX_LAMBDA);
    }
}
void fgo_resize_nearest(const float* src, int64_t w, int64_t h, int64_t nw, int64_t nh, float* dst) { /* model.rs:77-98 */
    if (nw == w && nh == h) { memcpy(dst, src, sizeof(float) * (size_t)(w * h)); return; }
    if (nw == 0 || nh == 0) return;
    if (w == 0 || h == 0) { memset(dst, 0, sizeof(float) * (size_t)(nw * nh)); return; }
    float scale_x = (float)w / (float)nw;
    float scale_y = (float)h / (float)nh;
    for (int64_t y = 0; y < nh; ++y) {
        float src_y = clampf(((float)y + 0.5f) * scale_y - 0.5f, 0.0f, (float)(h - 1));
        int64_t sy = (int64_t)roundf(src_y);
        for (int64_t x = 0; x < nw; ++x) {
            float src_x = clampf(((float)x + 0.5f) * scale_x - 0.5f, 0.0f, (float)(w - 1));
            int64_t sx = (int64_t)roundf(src_x);
            dst[y * nw + x] = src[sy * w + sx];
        }
    }
}

static inline float radius_sample(const fgo_derived* d, fgo_rng* rng) { /* model.rs:137-148 */
    if (d->radius_dist == FGO_DIST_CONST) return d->mean_linear;
    if (d->has_log) return (float)fgo_lognormal_f64_sample(rng, d->log_mu, d->log_sigma);
    return d->mean_linear;
}
static inline float get_clamped(const float* data, int64_t w, int64_t h, int64_t x, int64_t y) { /* model.rs:60-67 */
    if (w * h == 0) return 0.0f;
    int64_t xi = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
    int64_t yi = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
    return data[yi * w + xi];
}

/* -------------------------------------------------------- src/pixelwise.rs ---- */
static float evaluate_indicator(float xg, float yg, const float* lambda, const fgo_params* p,
                                const fgo_derived* d, fgo_counters* c) { /* pixelwise.rs:47-106 */
    float rm = d->rm, delta = d->delta;
    if (rm <= 0.0f) return 0.0f;
    int32_t i0 = sat_i32_f32(floorf((xg - rm) / delta));
    int32_t i1 = sat_i32_f32(floorf((xg + rm) / delta));
    int32_t j0 = sat_i32_f32(floorf((yg - rm) / delta));
    int32_t j1 = sat_i32_f32(floorf((yg + rm) / delta));
    if (i0 > i1 || j0 > j1) return 0.0f;
    float uscale = fgo_uniform_f32_scale(0.0f, delta); /* Uniform::new(0.0f32, delta) :64 */
    for (int64_t i_delta = i0; i_delta <= i1; ++i_delta) {
        for (int64_t j_delta = j0; j_delta <= j1; ++j_delta) {
            c->cell_visits++;
            fgo_rng rng;
            fgo_cell_rng(&rng, p->seed, (int32_t)i_delta, (int32_t)j_delta);
            float sample_x = (float)(int32_t)i_delta * delta;
            float sample_y = (float)(int32_t)j_delta * delta;
            int64_t ix = sat_i64_f32(floorf(sample_x));
            int64_t iy = sat_i64_f32(floorf(sample_y));
            float lambda_cell = get_clamped(lambda, d->input_width, d->input_height, ix, iy);
            if (lambda_cell <= 0.0f) continue;
            float expected = lambda_cell * delta * delta;
            if (expected <= 0.0f) continue;
            uint32_t q = sat_u32_f64(poisson_sample_f64(&rng, (double)expected));
            if (q == 0) continue;
            for (uint32_t g = 0; g < q; ++g) {
                float cx = sample_x + fgo_uniform_f32_sample(&rng, 0.0f, uscale);
                float cy = sample_y + fgo_uniform_f32_sample(&rng, 0.0f, uscale);
                float radius = radius_sample(d, &rng);
                c->grains_drawn++;
                if (radius > d->rm) radius = d->rm;
                if (radius <= 0.0f) continue;
                float dx = xg - cx;
                float dy = yg - cy;
                c->grain_tests++;
                float dx2 = dx * dx, dy2 = dy * dy, r2 = radius * radius;
                if (dx2 + dy2 <= r2) return 1.0f;
            }
        }
    }
    return 0.0f;
}

int fgo_render_pixelwise(const float* lambda, const fgo_params* p, const fgo_derived* d,
                         const float* offsets_input, float* out, int64_t y0, int64_t y1,
                         int nthreads, fgo_counters* cnt) { /* pixelwise.rs:11-45 */
    int64_t out_w = d->output_width, out_h = d->output_height;
    if (y0 < 0) y0 = 0;
    if (y1 > out_h) y1 = out_h;
    uint32_t n = p->n_samples < 1 ? 1 : p->n_samples;
    float inv_samples = 1.0f / (float)n;
    float inv_zoom = 1.0f / p->zoom;
    uint64_t t_se = 0, t_cv = 0, t_gt = 0, t_gd = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : t_se, t_cv, t_gt, t_gd)
    for (int64_t y = y0; y < y1; ++y) {
        fgo_counters c = {0, 0, 0, 0};
        float* row = out + y * out_w;
        for (int64_t x = 0; x < out_w; ++x) {
            float sum = 0.0f;
            for (uint32_t k = 0; k < p->n_samples; ++k) {
                float xg = (((float)x + 0.5f) * inv_zoom) - offsets_input[2 * k + 0];
                float yg = (((float)y + 0.5f) * inv_zoom) - offsets_input[2 * k + 1];
                sum += evaluate_indicator(xg, yg, lambda, p, d, &c);
                c.sample_evals++;
            }
            row[x] = sum * inv_samples;
        }
        t_se += c.sample_evals; t_cv += c.cell_visits; t_gt += c.grain_tests; t_gd += c.grains_drawn;
    }
    if (cnt) { cnt->sample_evals += t_se; cnt->cell_visits += t_cv; cnt->grain_tests += t_gt; cnt->grains_drawn += t_gd; }
    return 0;
}

/* -------------------------------------------------------- src/grainwise.rs ---- */
static int gw_bounds(float center, float radius, int32_t limit, int32_t* lo, int32_t* hi) { /* grainwise.rs:126-142 */
    int32_t mn = sat_i32_f32(ceilf((center - radius) - 0.5f));
    int32_t mx = sat_i32_f32(floorf((center + radius) - 0.5f));
    if (mx < mn) return 0;
    if (limit <= 0) return 0;
    int32_t last = limit - 1;
    if (mn > last || mx < 0) return 0;
    mn = mn < 0 ? 0 : (mn > last ? last : mn);
    mx = mx < 0 ? 0 : (mx > last ? last : mx);
    if (mn > mx) return 0;
    *lo = mn; *hi = mx;
    return 1;
}

int fgo_render_grainwise(const float* lambda, const fgo_params* p, const fgo_derived* d,
                         const float* offsets, float* out, int nthreads, fgo_counters* cnt) { /* grainwise.rs:12-124 */
    int64_t out_w = d->output_width, out_h = d->output_height;
    size_t total = (size_t)(out_w * out_h);
    size_t lanes = ((size_t)p->n_samples + 63) / 64;
    uint64_t* bitsets = (uint64_t*)calloc((total * lanes) > 0 ? total * lanes : 1, sizeof(uint64_t));
    if (!bitsets) return -2;
    float zoom = p->zoom;
    uint32_t n = p->n_samples < 1 ? 1 : p->n_samples;
    float inv_samples = 1.0f / (float)n;
    uint64_t t_gd = 0, t_gt = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : t_gd, t_gt)
    for (int64_t y = 0; y < d->input_height; ++y) {
        for (int64_t x = 0; x < d->input_width; ++x) {
            float lambda_val = lambda[y * d->input_width + x];
            if (lambda_val <= 0.0f) continue;
            fgo_rng rng;
            fgo_pixel_rng(&rng, p->seed, (int32_t)x, (int32_t)y);
            uint32_t q = sat_u32_f64(poisson_sample_f64(&rng, (double)lambda_val));
            if (q == 0) continue;
            float uscale = fgo_uniform_f32_scale(0.0f, 1.0f);
            for (uint32_t g = 0; g < q; ++g) {
                float cx = (float)x + fgo_uniform_f32_sample(&rng, 0.0f, uscale);
                float cy = (float)y + fgo_uniform_f32_sample(&rng, 0.0f, uscale);
                float radius = radius_sample(d, &rng);
                t_gd++;
                if (radius > d->rm) radius = d->rm;
                if (radius <= 0.0f) continue;
                float radius_out = radius * zoom;
                if (radius_out <= 0.0f) continue;
                float radius_sq = radius_out * radius_out;
                for (uint32_t k = 0; k < p->n_samples; ++k) {
                    float tx = (cx * zoom) + offsets[2 * k + 0];
                    float ty = (cy * zoom) + offsets[2 * k + 1];
                    int32_t x_min, x_max, y_min, y_max;
                    t_gt++;
                    if (!gw_bounds(tx, radius_out, (int32_t)out_w, &x_min, &x_max)) continue;
                    if (!gw_bounds(ty, radius_out, (int32_t)out_h, &y_min, &y_max)) continue;
                    size_t lane_idx = k / 64;
                    uint64_t bit_mask = 1ULL << (k % 64);
                    for (int32_t oy = y_min; oy <= y_max; ++oy) {
                        float center_y = (float)oy + 0.5f;
                        float dy = center_y - ty;
                        float dy_sq = dy * dy;
                        if (dy_sq > radius_sq) continue;
                        for (int32_t ox = x_min; ox <= x_max; ++ox) {
                            float center_x = (float)ox + 0.5f;
                            float dx = center_x - tx;
                            float dx2 = dx * dx;
                            if (dx2 + dy_sq <= radius_sq) {
                                size_t idx = (size_t)oy * (size_t)out_w + (size_t)ox;
                                __atomic_fetch_or(&bitsets[idx * lanes + lane_idx], bit_mask, __ATOMIC_RELAXED);
                            }
                        }
                    }
                }
            }
        }
    }
    for (size_t pix = 0; pix < total; ++pix) { /* grainwise.rs:114-122 */
        uint32_t count = 0;
        for (size_t lane = 0; lane < lanes; ++lane) count += (uint32_t)__builtin_popcountll(bitsets[pix * lanes + lane]);
        out[pix] = (float)count * inv_samples;
    }
    free(bitsets);
    if (cnt) { cnt->grains_drawn += t_gd; cnt->grain_tests += t_gt; cnt->sample_evals += (uint64_t)total * p->n_samples; }
    return 0;
}

uint32_t fgo_gen_cell(const fgo_params* p, const fgo_derived* d, int which_stream, int32_t i,
                      int32_t j, float lambda_cell, float* cx, float* cy, float* rad, uint32_t cap) {
    fgo_rng rng;
    float ox, oy, hi, mean;
    if (which_stream == 2) { /* grainwise.rs:37-57 */
        if (lambda_cell <= 0.0f) return 0;
        fgo_pixel_rng(&rng, p->seed, i, j);
        ox = (float)i; oy = (float)j; hi = 1.0f; mean = lambda_cell;
    } else { /* pixelwise.rs:68-95 */
        fgo_cell_rng(&rng, p->seed, i, j);
        ox = (float)i * d->delta; oy = (float)j * d->delta; hi = d->delta;
        if (lambda_cell <= 0.0f) return 0;
        mean = lambda_cell * d->delta * d->delta;
        if (mean <= 0.0f) return 0;
    }
    uint32_t q = sat_u32_f64(poisson_sample_f64(&rng, (double)mean));
    float uscale = fgo_uniform_f32_scale(0.0f, hi);
    for (uint32_t g = 0; g < q; ++g) {
        float x = ox + fgo_uniform_f32_sample(&rng, 0.0f, uscale);
        float y = oy + fgo_uniform_f32_sample(&rng, 0.0f, uscale);
        float radius = radius_sample(d, &rng);
        if (radius > d->rm) radius = d->rm;
        if (g < cap) { cx[g] = x; cy[g] = y; rad[g] = radius; }
    }
    return q;
}

void fgo_gen_cells(const fgo_params* p, const fgo_derived* d, int which_stream, const int32_t* ij,
                   const float* lambda_cell, size_t n, uint32_t cap, uint32_t* q_out, float* grains_out) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)n; ++t) {
        float cx[64], cy[64], r[64];
        uint32_t c2 = cap < 64 ? cap : 64;
        uint32_t q = fgo_gen_cell(p, d, which_stream, ij[2 * t], ij[2 * t + 1], lambda_cell[t], cx, cy, r, c2);
        q_out[t] = q;
        for (uint32_t g = 0; g < cap; ++g) {
            float* o = grains_out + ((size_t)t * cap + g) * 3;
            if (g < q && g < c2) { o[0] = cx[g]; o[1] = cy[g]; o[2] = r[g]; }
            else { o[0] = o[1] = o[2] = 0.0f; }
        }
    }
}

/* ------------------------------------------------------------ src/color.rs ---- */
static const float Y_COEFF_R = 0.2126f, Y_COEFF_G = 0.7152f, Y_COEFF_B = 0.0722f; /* color.rs:9-11 */
static const float CB_DENOM = 1.8556f, CR_DENOM = 1.5748f;                        /* color.rs:12-13 */
static inline float clamp01(float v) { return clampf(v, 0.0f, 1.0f); }
uint8_t fgo_to_u8(float v) { return (uint8_t)floorf(clamp01(v) * 255.0f + 0.5f); } /* color.rs:237-239 */

void fgo_load_rgb_u8(const uint8_t* rgb, size_t npix, float* r, float* g, float* b) { /* color.rs:158-179 */
    for (size_t k = 0; k < npix; ++k) { /* image 0.25.8 to_rgb32f: c as f32 / 255.0 */
        r[k] = clamp01((float)rgb[3 * k + 0] / 255.0f);
        g[k] = clamp01((float)rgb[3 * k + 1] / 255.0f);
        b[k] = clamp01((float)rgb[3 * k + 2] / 255.0f);
    }
}
void fgo_load_luma_u8(const uint8_t* rgb, size_t npix, float* y, float* cb, float* cr) { /* color.rs:181-213 */
    for (size_t k = 0; k < npix; ++k) {
        float r = clamp01((float)rgb[3 * k + 0] / 255.0f);
        float g = clamp01((float)rgb[3 * k + 1] / 255.0f);
        float b = clamp01((float)rgb[3 * k + 2] / 255.0f);
        float t0 = Y_COEFF_R * r, t1 = Y_COEFF_G * g, t2 = Y_COEFF_B * b;
        float luma = (t0 + t1) + t2;
        cb[k] = (b - luma) / CB_DENOM;
        cr[k] = (r - luma) / CR_DENOM;
        y[k] = clamp01(luma);
    }
}
void fgo_store_rgb_u8(const float* r, const float* g, const float* b, size_t npix, uint8_t* rgb) { /* color.rs:98-112 */
    for (size_t k = 0; k < npix; ++k) {
        rgb[3 * k + 0] = fgo_to_u8(r[k]); rgb[3 * k + 1] = fgo_to_u8(g[k]); rgb[3 * k + 2] = fgo_to_u8(b[k]);
    }
}
void fgo_store_luma_u8(const float* y, const float* cb, const float* cr, size_t npix, uint8_t* rgb) { /* color.rs:68-97 */
    for (size_t k = 0; k < npix; ++k) {
        float y_val = y[k], cb_val = cb[k], cr_val = cr[k];
        float tr = CR_DENOM * cr_val; float r = clamp01(y_val + tr);
        float tb = CB_DENOM * cb_val; float b = clamp01(y_val + tb);
        float m0 = Y_COEFF_R * r, m1 = Y_COEFF_B * b;
        float g_unclamped = ((y_val - m0) - m1) / Y_COEFF_G;
        float g = clamp01(g_unclamped);
        rgb[3 * k + 0] = fgo_to_u8(r); rgb[3 * k + 1] = fgo_to_u8(g); rgb[3 * k + 2] = fgo_to_u8(b);
    }
}

/* ------------------------------------------------- src/lib.rs:134-173 pipeline ---- */
int fgo_render_rgb8(const uint8_t* rgb, int64_t in_w, int64_t in_h, const fgo_params* p, int color_mode,
                    uint8_t* out_rgb, int nthreads, int* algo_out, fgo_counters* cnt, char* msg, size_t msg_len) {
    size_t npix = (size_t)(in_w * in_h);
    fgo_derived d;
    float* offsets = (float*)malloc(sizeof(float) * 2 * (size_t)p->n_samples);
    float* offsets_input = (float*)malloc(sizeof(float) * 2 * (size_t)p->n_samples);
    int rc = fgo_derive_common(p, in_w, in_h, &d, offsets, offsets_input, msg, msg_len);
    if (rc) { free(offsets); free(offsets_input); return rc; }
    int algo = fgo_choose_algorithm(p, &d);
    if (algo_out) *algo_out = algo;
    size_t nout = (size_t)(d.output_width * d.output_height);
    float* planes[3]; float* outp[3];
    for (int c = 0; c < 3; ++c) { planes[c] = (float*)malloc(sizeof(float) * npix); outp[c] = (float*)malloc(sizeof(float) * nout); }
    float* norm = (float*)malloc(sizeof(float) * npix);
    float* lam = (float*)malloc(sizeof(float) * npix);
    int nplanes;
    if (color_mode == 0) { fgo_load_luma_u8(rgb, npix, planes[0], planes[1], planes[2]); nplanes = 1; }
    else { fgo_load_rgb_u8(rgb, npix, planes[0], planes[1], planes[2]); nplanes = 3; }
    for (int c = 0; c < nplanes && rc == 0; ++c) { /* for_each_plane, color.rs:47-64 */
        fgo_normalize_plane(planes[c], npix, norm);
        fgo_lambda_plane(norm, npix, d.inv_e_pi_r2, lam);
        if (algo == FGO_ALGO_PIXEL) rc = fgo_render_pixelwise(lam, p, &d, offsets_input, outp[c], 0, d.output_height, nthreads, cnt);
        else rc = fgo_render_grainwise(lam, p, &d, offsets, outp[c], nthreads, cnt);
    }
    if (rc == 0) {
        if (color_mode == 0) { /* into_rgb_image luma branch, color.rs:68-97 */
            fgo_resize_nearest(planes[1], in_w, in_h, d.output_width, d.output_height, outp[1]);
            fgo_resize_nearest(planes[2], in_w, in_h, d.output_width, d.output_height, outp[2]);
            fgo_store_luma_u8(outp[0], outp[1], outp[2], nout, out_rgb);
        } else {
            fgo_store_rgb_u8(outp[0], outp[1], outp[2], nout, out_rgb);
        }
    }
    for (int c = 0; c < 3; ++c) { free(planes[c]); free(outp[c]); }
    free(norm); free(lam); free(offsets); free(offsets_input);
    return rc;
}

int fgo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
