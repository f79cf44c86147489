// film_grain.cpp -- host-side mirror of the reference's library API (see film_grain.hpp).
// Compiled with -ffp-contract=off: every f32 expression keeps the reference's operand order and
// is never fused, so Derived (delta, rm, inv_e_pi_r2, offsets) is bit-identical to the Rust host's.
#include "film_grain.hpp"
#include "image_io.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <mutex>

#include "../csrc/fg_logf.h"
#include "../csrc/fg_zig_tables.h"

namespace film_grain {

namespace {

constexpr float EPSILON = 1e-6f;    // src/model.rs:10
constexpr float MAX_LAMBDA = 1.0e6f; // src/model.rs:11

inline float f32_max(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
inline float f32_min(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
inline float f32_clamp(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline uint64_t rotl64(uint64_t x, unsigned k) { return (x << k) | (x >> (64 - k)); }
inline uint32_t rotl32(uint32_t x, unsigned k) { return (x << k) | (x >> (32 - k)); }

// ---- src/rng.rs:36-52 ----
uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
constexpr uint64_t OFFSET_STREAM = 0x9E3779B97F4A7C15ULL; // src/rng.rs:5

// ---- StdRng = ChaCha12 (rand 0.8.5 / rand_chacha 0.3.1) seeded by rand_core's seed_from_u64 ----
class StdRng {
  public:
    explicit StdRng(uint64_t state) { // SeedableRng::seed_from_u64: PCG32 fill of the 32-byte key
        const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
        for (int w = 0; w < 8; ++w) {
            state = state * MUL + INC;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            key_[w] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        }
    }
    uint64_t next_u64() { // BlockRng::next_u64 over a 4-block (64-word) buffer
        if (index_ < 63) {
            uint64_t v = ((uint64_t)buf_[index_ + 1] << 32) | buf_[index_];
            index_ += 2;
            return v;
        }
        if (index_ >= 64) {
            refill();
            index_ = 2;
            return ((uint64_t)buf_[1] << 32) | buf_[0];
        }
        uint64_t x = buf_[63];
        refill();
        index_ = 1;
        return ((uint64_t)buf_[0] << 32) | x;
    }

  private:
    static void qr(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
        a += b; d ^= a; d = rotl32(d, 16);
        c += d; b ^= c; b = rotl32(b, 12);
        a += b; d ^= a; d = rotl32(d, 8);
        c += d; b ^= c; b = rotl32(b, 7);
    }
    void block(uint64_t counter, uint32_t* out) const {
        uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key_[0], key_[1], key_[2], key_[3],
                           key_[4], key_[5], key_[6], key_[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
        uint32_t x[16];
        std::memcpy(x, in, sizeof x);
        for (int r = 0; r < 6; ++r) {
            qr(x[0], x[4], x[8], x[12]); qr(x[1], x[5], x[9], x[13]); qr(x[2], x[6], x[10], x[14]); qr(x[3], x[7], x[11], x[15]);
            qr(x[0], x[5], x[10], x[15]); qr(x[1], x[6], x[11], x[12]); qr(x[2], x[7], x[8], x[13]); qr(x[3], x[4], x[9], x[14]);
        }
        for (int k = 0; k < 16; ++k) out[k] = x[k] + in[k];
    }
    void refill() {
        for (int b = 0; b < 4; ++b) block(counter_ + (uint64_t)b, buf_ + 16 * b);
        counter_ += 4;
    }
    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t buf_[64];
    int index_ = 64;
};

double standard_f64(StdRng& r) { return (double)(r.next_u64() >> 11) * (1.0 / 9007199254740992.0); }
double open01_f64(StdRng& r) {
    uint64_t bits = (r.next_u64() >> 12) | 0x3FF0000000000000ULL;
    double v;
    std::memcpy(&v, &bits, 8);
    return v - (1.0 - 2.220446049250313e-16 / 2.0);
}
// rand_distr 0.4.3 StandardNormal (ziggurat, 256 layers)
double standard_normal(StdRng& rng) {
    for (;;) {
        uint64_t bits = rng.next_u64();
        size_t i = (size_t)(bits & 0xff);
        uint64_t fb = (bits >> 12) | 0x4000000000000000ULL;
        double f;
        std::memcpy(&f, &fb, 8);
        double u = f - 3.0;
        double x = u * FG_ZIG_NORM_X_INIT[i];
        if (std::fabs(x) < FG_ZIG_NORM_X_INIT[i + 1]) return x;
        if (i == 0) {
            double tx = 1.0, ty = 0.0;
            while (-2.0 * ty < tx * tx) {
                double x_ = open01_f64(rng);
                double y_ = open01_f64(rng);
                tx = std::log(x_) / FG_ZIG_NORM_R;
                ty = std::log(y_);
            }
            return (u < 0.0) ? tx - FG_ZIG_NORM_R : FG_ZIG_NORM_R - tx;
        }
        if (FG_ZIG_NORM_F_INIT[i + 1] + (FG_ZIG_NORM_F_INIT[i] - FG_ZIG_NORM_F_INIT[i + 1]) * standard_f64(rng) <
            std::exp(-x * x / 2.0))
            return x;
    }
}

// statrs 0.16.1 Normal::inverse_cdf stand-in (src/model.rs:159-161): Acklam + two Halley steps on
// erfc, ~1 ulp of f64; its only consumer rounds exp(mu + sigma*z) to f32.
double norm_inv_cdf(double p) {
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    double x, q, r;
    if (p < 0.02425) {
        q = std::sqrt(-2 * std::log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        q = p - 0.5;
        r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        q = std::sqrt(-2 * std::log(1 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    for (int it = 0; it < 2; ++it) {
        double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
        double u = e * std::sqrt(2.0 * 3.14159265358979323846) * std::exp(x * x / 2.0);
        x = x - u / (1.0 + x * u / 2.0);
    }
    return x;
}

float ensure_positive(float v, const char* field) { // src/params.rs:263-271
    if (!std::isfinite(v)) throw ParamsError(field, "value must be finite");
    if (v <= 0.0f) throw ParamsError(field, "value must be greater than 0");
    return v;
}

} // namespace

// ------------------------------------------------------------------ src/params.rs
float default_cell_delta(float radius_mean) { // :255-261
    if (radius_mean <= 0.0f) return 1.0f;
    float inv = f32_max(std::ceil(1.0f / radius_mean), 1.0f);
    return 1.0f / inv;
}

Params ParamsBuilder::build() const { // :141-180
    Params p;
    p.radius_mean = ensure_positive(radius_mean, "radius");
    if (!std::isfinite(radius_stddev)) throw ParamsError("radius-stddev", "value must be finite");
    if (radius_stddev < 0.0f) throw ParamsError("radius-stddev", "value must be >= 0");
    p.zoom = ensure_positive(zoom, "zoom");
    p.sigma_px = ensure_positive(sigma_px, "sigma");
    p.n_samples = std::max<uint32_t>(n_samples, 1);
    if (max_radius.kind == MaxRadius::Absolute) { // :298-328
        if (!std::isfinite(max_radius.value)) throw ParamsError("max-radius", "absolute radius must be finite");
        if (max_radius.value <= 0.0f) throw ParamsError("max-radius", "absolute radius must be > 0");
    } else {
        if (!std::isfinite(max_radius.value)) throw ParamsError("max-radius", "quantile must be finite");
        if (!(0.0f < max_radius.value && max_radius.value < 1.0f))
            throw ParamsError("max-radius", "quantile must lie in the open interval (0,1)");
    }
    p.max_radius = max_radius;
    if (size) { // :344-358
        if (size->first == 0) throw ParamsError("size", "output width must be > 0");
        if (size->second && *size->second == 0) throw ParamsError("size", "output height must be > 0");
    }
    p.size = size;
    if (cell_delta) { // :283-296
        if (!std::isfinite(*cell_delta)) throw ParamsError("cell", "cell size must be finite");
        if (*cell_delta <= 0.0f) throw ParamsError("cell", "cell size must be > 0");
        p.cell_delta = cell_delta;
    } else {
        p.cell_delta = default_cell_delta(p.radius_mean);
    }
    p.radius_dist = radius_dist;
    if (radius_dist == RadiusDist::Const) { // derive_radius_parameters, :223-253
        p.radius_stddev = 0.0f;
    } else if (radius_stddev == 0.0f) {
        p.radius_stddev = 0.0f;
        p.radius_log_mu = std::log(p.radius_mean);
        p.radius_log_sigma = 0.0f;
    } else {
        float variance_ratio = (radius_stddev * radius_stddev) / (p.radius_mean * p.radius_mean);
        float sigma_sq = std::log(1.0f + variance_ratio);
        float sigma = std::sqrt(sigma_sq);
        float mu = std::log(p.radius_mean) - 0.5f * sigma_sq;
        p.radius_stddev = radius_stddev;
        p.radius_log_mu = mu;
        p.radius_log_sigma = sigma;
    }
    p.algo = algo;
    p.color_mode = color_mode;
    p.seed = seed;
    return p;
}

// ------------------------------------------------------------------ src/model.rs
Plane Plane::resize_nearest(size_t new_width, size_t new_height) const { // :77-98
    if (new_width == width && new_height == height) return *this;
    Plane result(new_width, new_height);
    if (new_width == 0 || new_height == 0 || width == 0 || height == 0) return result;
    float scale_x = (float)width / (float)new_width;
    float scale_y = (float)height / (float)new_height;
    for (size_t y = 0; y < new_height; ++y) {
        float src_y = f32_clamp(((float)y + 0.5f) * scale_y - 0.5f, 0.0f, (float)(height - 1));
        size_t sy = (size_t)std::round(src_y);
        for (size_t x = 0; x < new_width; ++x) {
            float src_x = f32_clamp(((float)x + 0.5f) * scale_x - 0.5f, 0.0f, (float)(width - 1));
            size_t sx = (size_t)std::round(src_x);
            result.data[y * new_width + x] = get(sx, sy);
        }
    }
    return result;
}

float RadiusProfile::quantile(float p) const { // :150-164
    if (dist == RadiusDist::Const) return mean_linear;
    double mu = log_mu ? *log_mu : (double)mean_linear;
    double sigma = log_sigma ? *log_sigma : 0.0;
    if (sigma == 0.0) return (float)std::exp(mu);
    double z = norm_inv_cdf((double)p);
    return (float)std::exp(mu + sigma * z);
}

std::vector<std::array<float, 2>> make_offsets(uint64_t seed, size_t n, float sigma) { // src/rng.rs:9-24
    std::vector<std::array<float, 2>> out;
    if (n == 0) return out;
    double sd = (double)f32_max(sigma, 1.1920929e-7f);
    StdRng rng(splitmix64(seed ^ OFFSET_STREAM));
    out.resize(n);
    for (size_t k = 0; k < n; ++k) {
        double a = 0.0 + sd * standard_normal(rng); // Normal::sample = mean + std_dev * z
        double b = 0.0 + sd * standard_normal(rng);
        out[k] = {(float)a, (float)b};
    }
    return out;
}

Derived derive_common(const Params& params, size_t input_width, size_t input_height) { // :181-226
    if (input_width == 0 || input_height == 0) throw RenderError(RenderError::Message, "input image is empty after ROI");
    size_t ow, oh; // resolve_output_size, :267-298
    if (params.size) {
        ow = params.size->first;
        if (ow == 0) throw RenderError(RenderError::Message, "output width must be > 0");
        if (params.size->second) oh = *params.size->second;
        else {
            float aspect = (float)input_height / (float)input_width;
            oh = (size_t)f32_max(std::round((float)ow * aspect), 1.0f);
        }
        if (oh == 0) throw RenderError(RenderError::Message, "output height must be > 0");
    } else {
        ow = (size_t)f32_max(std::ceil((float)input_width * params.zoom), 1.0f);
        oh = (size_t)f32_max(std::ceil((float)input_height * params.zoom), 1.0f);
    }
    if (ow == 0 || oh == 0) throw RenderError(RenderError::Message, "output dimensions must be positive");
    Derived d;
    d.input_width = input_width; d.input_height = input_height; d.output_width = ow; d.output_height = oh;
    const float PI = 3.14159265358979323846f;
    float mean_sq = params.radius_mean * params.radius_mean;
    float variance = params.radius_stddev * params.radius_stddev;
    d.inv_e_pi_r2 = 1.0f / (PI * f32_max(mean_sq + variance, EPSILON));
    d.radius.dist = params.radius_dist; // RadiusProfile::new, :110-135
    d.radius.mean_linear = params.radius_mean;
    if (params.radius_log_mu) d.radius.log_mu = (double)*params.radius_log_mu;
    if (params.radius_log_sigma) d.radius.log_sigma = (double)*params.radius_log_sigma;
    if (params.radius_dist == RadiusDist::Lognorm) {
        if (!d.radius.log_mu) throw RenderError(RenderError::Message, "missing log-normal mean; parameters were not derived");
        if (!d.radius.log_sigma) throw RenderError(RenderError::Message, "missing log-normal sigma; parameters were not derived");
        d.radius.lognormal = true;
    }
    float rm = params.max_radius.kind == MaxRadius::Absolute ? params.max_radius.value : d.radius.quantile(params.max_radius.value);
    d.rm = f32_max(rm, EPSILON);
    d.delta = f32_max(params.cell_delta ? *params.cell_delta : default_cell_delta(params.radius_mean), EPSILON);
    d.offsets = make_offsets(params.seed, params.n_samples, params.sigma_px);
    d.offsets_input.resize(d.offsets.size());
    for (size_t k = 0; k < d.offsets.size(); ++k)
        d.offsets_input[k] = {d.offsets[k][0] / params.zoom, d.offsets[k][1] / params.zoom};
    return d;
}

Algo choose_algorithm(const Params& params, const Derived& derived) { // src/choose.rs:4-26
    if (params.algo != Algo::Auto) return params.algo;
    float mean = f32_max(params.radius_mean, 1e-6f);
    float sigma_ratio = mean > 0.0f ? params.radius_stddev / mean : 0.0f;
    float rm_ratio = derived.rm / mean;
    if (sigma_ratio < 0.1f && mean < 0.5f && params.n_samples <= 64) return Algo::Pixel;
    if (rm_ratio > 8.0f || sigma_ratio > 0.6f || params.n_samples > 96) return Algo::Grain;
    if (params.n_samples <= 24 && rm_ratio < 5.0f) return Algo::Pixel;
    return Algo::Grain;
}

std::pair<Plane, float> normalize_plane(const Plane& plane) { // :228-250
    float max_value = 0.0f;
    for (float v : plane.data) max_value = f32_max(max_value, v);
    bool already = max_value <= 1.0f + EPSILON;
    Plane out(plane.width, plane.height);
    for (size_t k = 0; k < plane.data.size(); ++k) {
        float v = plane.data[k];
        if (!already) v = v / f32_max(max_value + EPSILON, EPSILON);
        out.data[k] = f32_clamp(v, 0.0f, 1.0f - EPSILON);
    }
    return {std::move(out), max_value};
}

Plane lambda_plane(const Plane& normalized, float inv_e_pi_r2) { // :252-265
    Plane out(normalized.width, normalized.height);
    for (size_t k = 0; k < normalized.data.size(); ++k) {
        float clamped = f32_clamp(normalized.data[k], 0.0f, 1.0f - EPSILON);
        float safe = f32_max(1.0f - clamped, EPSILON);
        float activity = -inv_e_pi_r2 * std::log(safe);
        out.data[k] = f32_min(activity, MAX_LAMBDA);
    }
    return out;
}

// ------------------------------------------------------------------ cuda module (the seam)
namespace cuda {

GpuContext::GpuContext(int device) : device_(device) {
    int rc = fg_context_create(&ctx_, device);
    if (rc != FG_OK) throw RenderError(RenderError::Gpu, std::string("gpu unavailable: ") + fg_error_string(rc), rc);
}
GpuContext::~GpuContext() { fg_context_destroy(ctx_); }

namespace {
std::mutex g_mu;
std::shared_ptr<GpuContext> g_ctx;
} // namespace

std::shared_ptr<GpuContext> context(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_ctx && g_ctx->device() == device) return g_ctx;
    g_ctx = std::make_shared<GpuContext>(device);
    return g_ctx;
}
void invalidate_context() {
    std::lock_guard<std::mutex> lock(g_mu);
    g_ctx.reset();
}

fg_params build_params_block(const Params& params, const Derived& d) {
    fg_params q;
    std::memset(&q, 0, sizeof q);
    q.struct_size = sizeof q;
    q.in_w = (uint32_t)d.input_width; q.in_h = (uint32_t)d.input_height;
    q.out_w = (uint32_t)d.output_width; q.out_h = (uint32_t)d.output_height;
    q.n_samples = params.n_samples;
    q.dist_kind = params.radius_dist == RadiusDist::Const ? FG_DIST_CONST : FG_DIST_LOGNORM;
    q.seeding = FG_SEEDING_RAND_0_8;
    q.seed = params.seed;
    q.zoom = params.zoom; q.delta = d.delta; q.rm = d.rm; q.inv_e_pi_r2 = d.inv_e_pi_r2;
    q.radius_mean = params.radius_mean;
    q.has_log = d.radius.lognormal ? 1u : 0u;
    q.radius_log_mu = d.radius.log_mu ? *d.radius.log_mu : 0.0;
    q.radius_log_sigma = d.radius.log_sigma ? *d.radius.log_sigma : 0.0;
    return q;
}

namespace {
[[noreturn]] void throw_gpu(const GpuContext& ctx, int rc, const char* label) {
    if (rc == FG_ERR_CANCELLED) throw RenderError(RenderError::Cancelled, "cancelled", rc);
    std::string msg = std::string(label) + ": " + fg_last_error(ctx.raw());
    if (rc == FG_ERR_OOM || rc == FG_ERR_CUDA_STICKY) invalidate_context(); // handle_gpu_error's `fatal`, src/wgpu/mod.rs:727-752
    throw RenderError(RenderError::Gpu, msg, rc);
}
} // namespace

Plane render_pixelwise_gpu(const GpuContext& ctx, const Plane& lambda, const Params& params, const Derived& d) {
    if (d.offsets_input.size() != params.n_samples) // src/wgpu/mod.rs:353-358
        throw RenderError(RenderError::Gpu, "offset count does not match sample count");
    fg_params q = build_params_block(params, d);
    Plane out(d.output_width, d.output_height);
    int rc = fg_render_pixelwise(ctx.raw(), &q, lambda.data.data(), &d.offsets_input[0][0], out.data.data());
    if (rc != FG_OK) throw_gpu(ctx, rc, "Pixel renderer");
    return out;
}

Plane render_grainwise_gpu(const GpuContext& ctx, const Plane& lambda, const Params& params, const Derived& d) {
    if (d.offsets.size() != params.n_samples) // src/wgpu/mod.rs:490-495
        throw RenderError(RenderError::Gpu, "offset count does not match sample count");
    fg_params q = build_params_block(params, d);
    Plane out(d.output_width, d.output_height);
    int rc = fg_render_grainwise(ctx.raw(), &q, lambda.data.data(), &d.offsets[0][0], out.data.data());
    if (rc != FG_OK) throw_gpu(ctx, rc, "Grain renderer");
    return out;
}

} // namespace cuda

// ------------------------------------------------------------------ src/color.rs + src/lib.rs
namespace {
constexpr float Y_COEFF_R = 0.2126f, Y_COEFF_G = 0.7152f, Y_COEFF_B = 0.0722f, CB_DENOM = 1.8556f, CR_DENOM = 1.5748f;
inline float clamp01(float v) { return f32_clamp(v, 0.0f, 1.0f); }
inline uint8_t to_u8(float v) { return (uint8_t)std::floor(clamp01(v) * 255.0f + 0.5f); } // :237-239

RenderStats make_stats(const Params& params, const Derived& d, Algo algo) { // src/lib.rs:205-222
    float mean = f32_max(params.radius_mean, 1e-6f);
    RenderStats s;
    s.algorithm = algo;
    s.input_w = d.input_width; s.input_h = d.input_height; s.output_w = d.output_width; s.output_h = d.output_height;
    s.n_samples = params.n_samples;
    s.sigma_ratio = mean > 0.0f ? params.radius_stddev / mean : 0.0f;
    s.rm_ratio = d.rm / mean;
    return s;
}
void check_cancel(const volatile int* cancel) {
    if (cancel && *cancel) throw RenderError(RenderError::Cancelled, "cancelled");
}
} // namespace

RenderStats dry_run_with_input_image(const InputImage& input, const Params& params) {
    Derived d = derive_common(params, input.width, input.height);
    return make_stats(params, d, choose_algorithm(params, d));
}

std::pair<RgbImage, RenderStats> render_with_input_image(const InputImage& input, const Params& params,
                                                         const volatile int* cancel, int device) {
    check_cancel(cancel);
    if (input.rgb.size() != input.width * input.height * 3) throw RenderError(RenderError::Message, "image buffer size mismatch");
    Derived d = derive_common(params, input.width, input.height);
    Algo algo = choose_algorithm(params, d);
    auto ctx = cuda::context(device); // shared_ptr: the context outlives this call even if another thread invalidates it
    const size_t npix = input.width * input.height, nout = d.output_width * d.output_height;
    const bool luma = params.color_mode == ColorMode::Luma;
    const int n_planes = luma ? 1 : 3;
    std::vector<Plane> planes(3, Plane(input.width, input.height));
    for (size_t k = 0; k < npix; ++k) { // load_luma_workspace / load_rgb_workspace, src/color.rs:158-213
        float r = clamp01((float)input.rgb[3 * k + 0] / 255.0f);
        float g = clamp01((float)input.rgb[3 * k + 1] / 255.0f);
        float b = clamp01((float)input.rgb[3 * k + 2] / 255.0f);
        if (luma) {
            float t0 = Y_COEFF_R * r, t1 = Y_COEFF_G * g, t2 = Y_COEFF_B * b;
            float l = (t0 + t1) + t2;
            planes[1].data[k] = (b - l) / CB_DENOM;
            planes[2].data[k] = (r - l) / CR_DENOM;
            planes[0].data[k] = clamp01(l);
        } else {
            planes[0].data[k] = r; planes[1].data[k] = g; planes[2].data[k] = b;
        }
    }
    // for_each_plane (src/color.rs:47-64): same seed/offsets for every plane -> one batched call
    std::vector<Plane> lambdas, outs(n_planes, Plane(d.output_width, d.output_height));
    for (int c = 0; c < n_planes; ++c) {
        check_cancel(cancel);
        lambdas.push_back(lambda_plane(normalize_plane(planes[c]).first, d.inv_e_pi_r2));
    }
    fg_params q = cuda::build_params_block(params, d);
    const float* lp[3];
    float* op[3];
    for (int c = 0; c < n_planes; ++c) { lp[c] = lambdas[c].data.data(); op[c] = outs[c].data.data(); }
    const bool pixel = algo == Algo::Pixel;
    // the cancel flag travels with the call (never stored in the shared context beyond it)
    int rc = fg_render_planes_cancelable(ctx->raw(), &q, pixel ? FG_ALGO_PIXEL : FG_ALGO_GRAIN, n_planes, lp,
                                         pixel ? &d.offsets_input[0][0] : &d.offsets[0][0], op, cancel);
    if (rc != FG_OK) cuda::throw_gpu(*ctx, rc, pixel ? "Pixel renderer" : "Grain renderer");
    check_cancel(cancel);
    RgbImage img;
    img.width = d.output_width; img.height = d.output_height;
    img.rgb.resize(nout * 3);
    if (luma) { // into_rgb_image, src/color.rs:68-97
        Plane cb = planes[1].resize_nearest(d.output_width, d.output_height);
        Plane cr = planes[2].resize_nearest(d.output_width, d.output_height);
        for (size_t k = 0; k < nout; ++k) {
            float y_val = outs[0].data[k];
            float tr = CR_DENOM * cr.data[k], tb = CB_DENOM * cb.data[k];
            float r = clamp01(y_val + tr), b = clamp01(y_val + tb);
            float m0 = Y_COEFF_R * r, m1 = Y_COEFF_B * b;
            float g = clamp01(((y_val - m0) - m1) / Y_COEFF_G);
            img.rgb[3 * k + 0] = to_u8(r); img.rgb[3 * k + 1] = to_u8(g); img.rgb[3 * k + 2] = to_u8(b);
        }
    } else {
        for (size_t k = 0; k < nout; ++k)
            for (int c = 0; c < 3; ++c) img.rgb[3 * k + c] = to_u8(outs[c].data[k]);
    }
    return {std::move(img), make_stats(params, d, algo)};
}

std::pair<RgbImage, RenderStats> render_with_input_image_fused(const InputImage& input, const Params& params, int device) {
    if (input.rgb.size() != input.width * input.height * 3) throw RenderError(RenderError::Message, "image buffer size mismatch");
    Derived d = derive_common(params, input.width, input.height);
    Algo algo = choose_algorithm(params, d);
    auto ctx = cuda::context(device);
    fg_params q = cuda::build_params_block(params, d);
    RgbImage img;
    img.width = d.output_width; img.height = d.output_height;
    img.rgb.resize(d.output_width * d.output_height * 3);
    const bool pixel = algo == Algo::Pixel;
    int rc = fg_render_rgb8(ctx->raw(), &q, pixel ? FG_ALGO_PIXEL : FG_ALGO_GRAIN,
                            params.color_mode == ColorMode::Luma ? FG_COLOR_LUMA : FG_COLOR_RGB, input.rgb.data(),
                            pixel ? &d.offsets_input[0][0] : &d.offsets[0][0], img.rgb.data());
    if (rc != FG_OK) cuda::throw_gpu(*ctx, rc, pixel ? "Pixel renderer" : "Grain renderer");
    return {std::move(img), make_stats(params, d, algo)};
}

} // namespace film_grain

// =============================================================================================
// C wrappers over the C++ mirror, so tests and bench.py (Python, ctypes) drive the same host code
// a C++ caller links against.  Declared in include/fg_host.h.
// =============================================================================================
#include "../../include/fg_host.h"

namespace {
using namespace film_grain;

thread_local std::string g_host_err;

ParamsBuilder to_builder(const fgh_params* p) {
    ParamsBuilder b;
    b.radius_dist = p->radius_dist == 1 ? RadiusDist::Lognorm : RadiusDist::Const;
    b.radius_mean = p->radius_mean;
    b.radius_stddev = p->radius_stddev;
    b.zoom = p->zoom;
    b.sigma_px = p->sigma_px;
    b.n_samples = p->n_samples;
    b.algo = p->algo == 1 ? Algo::Grain : (p->algo == 2 ? Algo::Pixel : Algo::Auto);
    b.max_radius.kind = p->max_radius_kind == 0 ? MaxRadius::Absolute : MaxRadius::Quantile;
    b.max_radius.value = p->max_radius_value;
    if (p->has_cell_delta) b.cell_delta = p->cell_delta;
    b.color_mode = p->color_mode == 1 ? ColorMode::Rgb : ColorMode::Luma;
    if (p->has_size) b.size = std::make_pair(p->size_w, p->has_size_h ? std::optional<uint32_t>(p->size_h) : std::nullopt);
    b.seed = p->seed;
    return b;
}

template <typename F>
int guarded(F&& f) {
    try {
        f();
        g_host_err.clear();
        return 0;
    } catch (const ParamsError& e) {
        g_host_err = std::string("parameter error: ") + e.what();
        return FGH_ERR_PARAMS;
    } catch (const RenderError& e) {
        g_host_err = e.what();
        switch (e.kind) {
        case RenderError::Gpu: return FGH_ERR_GPU;
        case RenderError::Cancelled: return FGH_ERR_CANCELLED;
        default: return FGH_ERR_MESSAGE;
        }
    } catch (const std::exception& e) {
        g_host_err = e.what();
        return FGH_ERR_MESSAGE;
    }
}

void fill_derived(const Params& params, const Derived& d, fgh_derived* out) {
    out->input_width = d.input_width; out->input_height = d.input_height;
    out->output_width = d.output_width; out->output_height = d.output_height;
    out->inv_e_pi_r2 = d.inv_e_pi_r2; out->rm = d.rm; out->delta = d.delta;
    out->radius_stddev = params.radius_stddev;
    out->has_log = d.radius.lognormal ? 1 : 0;
    out->log_mu = d.radius.log_mu ? *d.radius.log_mu : 0.0;
    out->log_sigma = d.radius.log_sigma ? *d.radius.log_sigma : 0.0;
    out->algorithm = (int)choose_algorithm(params, d);
    out->block = cuda::build_params_block(params, d);
}
} // namespace

extern "C" {

const char* fgh_last_error(void) { return g_host_err.c_str(); }

int fgh_derive(const fgh_params* p, uint64_t in_w, uint64_t in_h, fgh_derived* out, float* offsets, float* offsets_input) {
    return guarded([&] {
        Params params = to_builder(p).build();
        Derived d = derive_common(params, in_w, in_h);
        fill_derived(params, d, out);
        for (size_t k = 0; k < d.offsets.size(); ++k) {
            if (offsets) { offsets[2 * k] = d.offsets[k][0]; offsets[2 * k + 1] = d.offsets[k][1]; }
            if (offsets_input) { offsets_input[2 * k] = d.offsets_input[k][0]; offsets_input[2 * k + 1] = d.offsets_input[k][1]; }
        }
    });
}

int fgh_lambda_from_plane(const float* plane, uint64_t w, uint64_t h, float inv_e_pi_r2, float* lambda_out) {
    return guarded([&] {
        Plane pl(w, h);
        std::memcpy(pl.data.data(), plane, sizeof(float) * w * h);
        Plane lam = lambda_plane(normalize_plane(pl).first, inv_e_pi_r2);
        std::memcpy(lambda_out, lam.data.data(), sizeof(float) * w * h);
    });
}

int fgh_render_with_input_image(const fgh_params* p, const uint8_t* rgb, uint64_t w, uint64_t h, int fused, int device,
                                const volatile int* cancel, uint8_t* rgb_out, uint64_t out_capacity, fgh_derived* info) {
    return guarded([&] {
        Params params = to_builder(p).build();
        InputImage in;
        in.width = w; in.height = h;
        in.rgb.assign(rgb, rgb + w * h * 3);
        auto res = fused ? render_with_input_image_fused(in, params, device) : render_with_input_image(in, params, cancel, device);
        if (res.first.rgb.size() > out_capacity) throw RenderError(RenderError::Message, "output buffer too small");
        std::memcpy(rgb_out, res.first.rgb.data(), res.first.rgb.size());
        if (info) {
            Derived d = derive_common(params, w, h);
            fill_derived(params, d, info);
        }
    });
}

fg_ctx* fgh_context(int device) {
    // A raw pointer cannot carry the shared_ptr's reference: contexts handed out here are kept alive until the
    // process ends, so the pointer stays valid across fgh_invalidate_context() and device switches.
    static std::mutex mu;
    static std::vector<std::shared_ptr<cuda::GpuContext>> handed_out;
    fg_ctx* out = nullptr;
    guarded([&] {
        auto sp = cuda::context(device);
        std::lock_guard<std::mutex> lock(mu);
        bool known = false;
        for (auto& h : handed_out) known = known || h.get() == sp.get();
        if (!known) handed_out.push_back(sp);
        out = sp->raw();
    });
    return out;
}

void fgh_invalidate_context(void) { cuda::invalidate_context(); }

int fgh_render_file(const fgh_params* p, const char* input_path, const char* output_path, const char* format_token,
                    const uint32_t* roi4, int fused, int device, const volatile int* cancel, fgh_derived* info) {
    return guarded([&] {
        if (!p || !input_path || !output_path) throw RenderError(RenderError::Message, "NULL argument");
        Params params = to_builder(p).build();
        Roi roi{};
        if (roi4) { // ensure_roi, src/params.rs:330-341
            roi = Roi{roi4[0], roi4[1], roi4[2], roi4[3]};
            if (roi.x1 <= roi.x0 || roi.y1 <= roi.y0) throw ParamsError("roi", "roi end must be greater than start (exclusive bounds)");
        }
        const RenderStats st = render_file(params, input_path, output_path, format_token, roi4 ? &roi : nullptr, fused != 0, cancel, device);
        if (info) {
            Derived d = derive_common(params, st.input_w, st.input_h);
            fill_derived(params, d, info);
        }
    });
}

int fgh_load_image(const char* path, uint8_t** rgb, uint64_t* w, uint64_t* h) {
    return guarded([&] {
        if (!path || !rgb || !w || !h) throw RenderError(RenderError::Message, "NULL argument");
        InputImage img = load_image(path);
        uint8_t* buf = (uint8_t*)std::malloc(img.rgb.size() ? img.rgb.size() : 1);
        if (!buf) throw RenderError(RenderError::Message, "out of host memory");
        std::memcpy(buf, img.rgb.data(), img.rgb.size());
        *rgb = buf; *w = img.width; *h = img.height;
    });
}

void fgh_free(void* p) { std::free(p); }

int fgh_save_image(const char* path, const uint8_t* rgb, uint64_t w, uint64_t h, const char* format_token) {
    return guarded([&] {
        if (!path || !rgb) throw RenderError(RenderError::Message, "NULL argument");
        save_image(path, rgb, w, h, resolve_format(path, format_token));
    });
}

void fgh_logf_restated(const float* x, uint64_t n, float* out) {
    for (uint64_t k = 0; k < n; ++k) out[k] = fg::logf_libm(x[k]);
}

} // extern "C"
