/* fg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the Monte-Carlo hot path of joseph-wardle/film_grain
 * (src/rng.rs, src/model.rs, src/pixelwise.rs, src/grainwise.rs, src/choose.rs,
 * src/params.rs:223-261, src/color.rs) and of the third-party arithmetic that path
 * runs on (rand 0.8.5, rand_core 0.6.4, rand_chacha 0.3.1, rand_distr 0.4.3 -- pinned in
 * Cargo.lock:2478-2509; their sources are NOT under /root/reference, so the published
 * algorithms are restated here and pinned by the crates' own known-answer vectors in
 * tests/test_oracle_kat.py).
 *
 * PARITY STATUS: the reference holds no tests, fixtures or golden vectors for this path
 * and cannot be built here (no cargo/rustc), so the reference-level functions
 * (render_pixelwise, render_grainwise, ...) are "parity unpinned": they are pinned only
 * through the third-party KATs below them and by reading the src .rs files line by line.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library -- as the checker or the timed CPU baseline, never as a
 * product path.
 */
#ifndef FG_ORACLE_H
#define FG_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- generic RNG handle (xoshiro256++ = SmallRng, ChaCha12 = StdRng, PCG32 = the test
 *      generator rand/rand_distr use in their value-stability tests) ---------------- */
enum { FGO_RNG_XOSHIRO256PP = 0, FGO_RNG_CHACHA12 = 1, FGO_RNG_PCG32 = 2 };
typedef struct fgo_rng {
    int kind;
    uint64_t s[4];              /* xoshiro256++ */
    uint64_t pcg_state, pcg_inc;/* Lcg64Xsh32 */
    uint32_t cc_key[8];         /* ChaCha12 */
    uint64_t cc_counter;
    uint32_t cc_buf[64];
    int cc_index;
} fgo_rng;

uint64_t fgo_splitmix64(uint64_t x);                                   /* rng.rs:46-52 */
uint64_t fgo_mix(uint64_t seed, uint64_t stream);                      /* rng.rs:36-38 */
uint64_t fgo_mix3(uint64_t seed, uint64_t stream, int64_t a, int64_t b);/* rng.rs:40-44 */
uint64_t fgo_stream_const(int which);  /* 0 OFFSET, 1 CELL, 2 PIXEL   rng.rs:5-7 */

void fgo_seed_bytes_from_u64(uint64_t state, uint8_t out[32]);  /* rand_core seed_from_u64 */
void fgo_xoshiro_from_seed(fgo_rng* r, const uint8_t seed[32]);
void fgo_xoshiro_from_state(fgo_rng* r, const uint64_t s[4]);
void fgo_small_rng_seed_from_u64(fgo_rng* r, uint64_t state);   /* variant A (default) */
void fgo_small_rng_seed_from_u64_variant_b(fgo_rng* r, uint64_t state);
void fgo_chacha12_from_seed(fgo_rng* r, const uint8_t seed[32]);
void fgo_std_rng_seed_from_u64(fgo_rng* r, uint64_t state);
void fgo_pcg32_new(fgo_rng* r, uint64_t state, uint64_t stream);
uint64_t fgo_next_u64(fgo_rng* r);
uint32_t fgo_next_u32(fgo_rng* r);

void fgo_cell_rng(fgo_rng* r, uint64_t seed, int32_t i, int32_t j);   /* rng.rs:26-29 */
void fgo_pixel_rng(fgo_rng* r, uint64_t seed, int32_t i, int32_t j);  /* rng.rs:31-34 */
/* 1 = SmallRng::seed_from_u64 is rand_core's PCG32 fill (variant A, rand 0.8.5),
 * 2 = xoshiro's own SplitMix64 seeding (variant B, rand >= 0.9).  Process-global. */
void fgo_set_seeding_variant(int v);

/* ---- distributions ------------------------------------------------------------- */
double fgo_standard_f64(fgo_rng* r);
float fgo_standard_f32(fgo_rng* r);
double fgo_open01_f64(fgo_rng* r);
float fgo_uniform_f32_scale(float low, float high);         /* UniformFloat::new */
float fgo_uniform_f32_sample(fgo_rng* r, float low, float scale);
double fgo_log_gamma_f64(double x);
double fgo_poisson_f64_sample(fgo_rng* r, double lambda);
float fgo_poisson_f32_sample(fgo_rng* r, float lambda);    /* KAT pinning only */
double fgo_standard_normal_f64(fgo_rng* r);
double fgo_normal_f64_sample(fgo_rng* r, double mean, double std_dev);
double fgo_lognormal_f64_sample(fgo_rng* r, double mu, double sigma);
double fgo_norm_inv_cdf(double p);                         /* statrs stand-in, see .c */

/* ---- model / params ------------------------------------------------------------ */
enum { FGO_DIST_CONST = 0, FGO_DIST_LOGNORM = 1 };
enum { FGO_ALGO_AUTO = 0, FGO_ALGO_GRAIN = 1, FGO_ALGO_PIXEL = 2 };
enum { FGO_MAXR_ABSOLUTE = 0, FGO_MAXR_QUANTILE = 1 };

typedef struct fgo_params {      /* the subset of params.rs:45-68 the path reads */
    int radius_dist;
    float radius_mean;
    float radius_stddev;         /* as derived (0 for const) */
    int has_log;                 /* radius_log_mu / radius_log_sigma are Some */
    float radius_log_mu, radius_log_sigma;
    float zoom, sigma_px;
    uint32_t n_samples;
    int algo;
    int max_radius_kind; float max_radius_value;
    int has_cell_delta; float cell_delta;
    int has_size; uint32_t size_w; int has_size_h; uint32_t size_h;
    uint64_t seed;
} fgo_params;

typedef struct fgo_derived {     /* model.rs:167-179 (offsets held separately) */
    int64_t input_width, input_height, output_width, output_height;
    float inv_e_pi_r2, rm, delta;
    int radius_dist; float mean_linear; int has_log; double log_mu, log_sigma;
} fgo_derived;

float fgo_default_cell_delta(float radius_mean);                   /* params.rs:255-261 */
/* params.rs:141-180 + 223-253; returns 0 ok, <0 error (msg filled) */
int fgo_params_build(fgo_params* p, char* msg, size_t msg_len);
int fgo_derive_common(const fgo_params* p, int64_t in_w, int64_t in_h, fgo_derived* d,
                      float* offsets /*2N*/, float* offsets_input /*2N*/, char* msg,
                      size_t msg_len);                             /* model.rs:181-226 */
void fgo_make_offsets(uint64_t seed, size_t n, float sigma, float* out /*2n*/);/* rng.rs:9-24 */
int fgo_choose_algorithm(const fgo_params* p, const fgo_derived* d);/* choose.rs:4-26 */
float fgo_normalize_plane(const float* src, size_t n, float* dst); /* model.rs:228-250 */
void fgo_lambda_plane(const float* norm, size_t n, float inv_e_pi_r2, float* dst);/* :252-265 */
void fgo_resize_nearest(const float* src, int64_t w, int64_t h, int64_t nw, int64_t nh,
                        float* dst);                               /* model.rs:77-98 */

/* ---- integrators ------------------------------------------------------------- */
typedef struct fgo_counters {    /* work counters for SURVEY 8(d)'s op model */
    uint64_t sample_evals;       /* S */
    uint64_t cell_visits;        /* sum n_cell (with the reference's early exit) */
    uint64_t grain_tests;        /* sum n_test */
    uint64_t grains_drawn;
} fgo_counters;

/* render rows [y0,y1) of the output (full plane: 0,out_h); out is the full plane buffer.
 * nthreads <= 0: all OpenMP threads.  cnt may be NULL. */
int fgo_render_pixelwise(const float* lambda, const fgo_params* p, const fgo_derived* d,
                         const float* offsets_input, float* out, int64_t y0, int64_t y1,
                         int nthreads, fgo_counters* cnt);         /* pixelwise.rs:11-106 */
int fgo_render_grainwise(const float* lambda, const fgo_params* p, const fgo_derived* d,
                         const float* offsets, float* out, int nthreads,
                         fgo_counters* cnt);                       /* grainwise.rs:12-142 */

/* grain realisation of ONE cell, exactly as the integrators draw it.
 * which_stream 1 = CELL (pixel-wise: origin (i*delta, j*delta), uniform [0,delta),
 * mean = lambda*delta*delta), 2 = PIXEL (grain-wise: origin (i,j), uniform [0,1), mean = lambda).
 * Returns the Poisson count q (grains with radius<=0 are still counted and written). */
uint32_t fgo_gen_cell(const fgo_params* p, const fgo_derived* d, int which_stream, int32_t i,
                      int32_t j, float lambda_cell, float* cx, float* cy, float* rad,
                      uint32_t cap);

/* batched fgo_gen_cell: ij n x [i,j]; q_out n; grains_out n x cap x [cx,cy,r] (zero-filled) */
void fgo_gen_cells(const fgo_params* p, const fgo_derived* d, int which_stream, const int32_t* ij,
                   const float* lambda_cell, size_t n, uint32_t cap, uint32_t* q_out,
                   float* grains_out);

/* ---- colour (color.rs) ---------------------------------------------------------- */
void fgo_load_rgb_u8(const uint8_t* rgb, size_t npix, float* r, float* g, float* b);/* :158-179 */
void fgo_load_luma_u8(const uint8_t* rgb, size_t npix, float* y, float* cb, float* cr);/* :181-213 */
void fgo_store_rgb_u8(const float* r, const float* g, const float* b, size_t npix,
                      uint8_t* rgb);                                /* :98-112 */
void fgo_store_luma_u8(const float* y, const float* cb, const float* cr, size_t npix,
                       uint8_t* rgb);                               /* :68-97 */
uint8_t fgo_to_u8(float v);                                         /* :237-239 */

/* Whole pipeline lib.rs:134-173 on an 8-bit RGB image (the decode/encode outside it).
 * color_mode 0 luma, 1 rgb.  algo_out receives the algorithm actually used.
 * out_rgb must hold 3*out_w*out_h bytes where (out_w,out_h) = derive_common's size. */
int fgo_render_rgb8(const uint8_t* rgb, int64_t in_w, int64_t in_h, const fgo_params* p,
                    int color_mode, uint8_t* out_rgb, int nthreads, int* algo_out,
                    fgo_counters* cnt, char* msg, size_t msg_len);

int fgo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
