// film_grain.hpp -- host-side mirror of the reference's library API for the hot path, in C++
// (the reference is compiled Rust; there is no Rust toolchain in this image).  Same names,
// argument meaning and error behaviour as joseph-wardle/film_grain:
//
//   Params / ParamsBuilder::build / MaxRadius / RadiusDist / Algo / ColorMode   src/params.rs:6-180
//   default_cell_delta                                                          src/params.rs:255-261
//   Plane, RadiusProfile, Derived, derive_common, normalize_plane, lambda_plane src/model.rs
//   make_offsets                                                                src/rng.rs:9-24
//   choose_algorithm                                                            src/choose.rs:4-26
//   Workspace (load_luma / load_rgb / for_each_plane / into_rgb_image)          src/color.rs
//   render_with_input_image, RenderError, RenderStats                           src/lib.rs:28-173
//   cuda::context / render_pixelwise_gpu / render_grainwise_gpu                 src/wgpu/mod.rs:84-86, 336-345, 473-482
//
// The `cuda` module is the drop-in for the reference's `wgpu` module: it marshals Params/Derived
// into fg_params exactly where build_uniforms does (src/wgpu/mod.rs:661-692) and calls the C ABI
// (include/fg.h).  Device::Gpu has no CPU fallback; Device::Cpu is not offered here at all.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fg.h"

namespace film_grain {

enum class RadiusDist { Const, Lognorm };
enum class Algo { Auto, Grain, Pixel };
enum class ColorMode { Luma, Rgb };

struct MaxRadius {
    enum Kind { Absolute, Quantile } kind = Quantile;
    float value = 0.999f;
};

struct ParamsError : std::runtime_error { // src/params.rs:116-139
    std::string field, message;
    ParamsError(std::string f, std::string m) : std::runtime_error(f + ": " + m), field(std::move(f)), message(std::move(m)) {}
};

struct RenderError : std::runtime_error { // src/lib.rs:28-44
    enum Kind { Params, Gpu, Unsupported, Cancelled, Message } kind;
    int code; // C-ABI return code for Kind::Gpu
    RenderError(Kind k, const std::string& m, int c = 0) : std::runtime_error(m), kind(k), code(c) {}
};

struct Params { // src/params.rs:45-68 (path, roi, output format and flags stay with the caller)
    RadiusDist radius_dist = RadiusDist::Const;
    float radius_mean = 0.10f;
    float radius_stddev = 0.0f;
    std::optional<float> radius_log_mu, radius_log_sigma;
    float zoom = 1.0f;
    float sigma_px = 0.8f;
    uint32_t n_samples = 32;
    Algo algo = Algo::Auto;
    MaxRadius max_radius;
    std::optional<float> cell_delta;
    ColorMode color_mode = ColorMode::Luma;
    std::optional<std::pair<uint32_t, std::optional<uint32_t>>> size;
    uint64_t seed = 5489;
};

struct ParamsBuilder { // src/params.rs:70-91, defaults of src/main.rs:94-246
    RadiusDist radius_dist = RadiusDist::Const;
    float radius_mean = 0.10f;
    float radius_stddev = 0.0f;
    float zoom = 1.0f;
    float sigma_px = 0.8f;
    uint32_t n_samples = 32;
    Algo algo = Algo::Auto;
    MaxRadius max_radius;
    std::optional<float> cell_delta;
    ColorMode color_mode = ColorMode::Luma;
    std::optional<std::pair<uint32_t, std::optional<uint32_t>>> size;
    uint64_t seed = 5489;
    Params build() const; // throws ParamsError (src/params.rs:141-180)
};

float default_cell_delta(float radius_mean);

struct Plane { // src/model.rs:13-99
    size_t width = 0, height = 0;
    std::vector<float> data;
    Plane() = default;
    Plane(size_t w, size_t h) : width(w), height(h), data(w * h, 0.0f) {}
    float get(size_t x, size_t y) const { return data[y * width + x]; }
    Plane resize_nearest(size_t new_width, size_t new_height) const;
};

struct RadiusProfile { // src/model.rs:101-165
    RadiusDist dist = RadiusDist::Const;
    float mean_linear = 0.0f;
    std::optional<double> log_mu, log_sigma;
    bool lognormal = false;
    float quantile(float p) const;
};

struct Derived { // src/model.rs:167-179
    size_t input_width = 0, input_height = 0, output_width = 0, output_height = 0;
    float inv_e_pi_r2 = 0, rm = 0, delta = 0;
    std::vector<std::array<float, 2>> offsets, offsets_input;
    RadiusProfile radius;
};

std::vector<std::array<float, 2>> make_offsets(uint64_t seed, size_t n, float sigma); // src/rng.rs:9-24
Derived derive_common(const Params& params, size_t input_width, size_t input_height); // throws RenderError::Message
Algo choose_algorithm(const Params& params, const Derived& derived);
std::pair<Plane, float> normalize_plane(const Plane& plane);
Plane lambda_plane(const Plane& normalized, float inv_e_pi_r2);

struct RenderStats { // src/lib.rs:46-55
    Algo algorithm;
    size_t input_w, input_h, output_w, output_h;
    uint32_t n_samples;
    float sigma_ratio, rm_ratio;
};

namespace cuda { // drop-in for `pub mod wgpu` (src/lib.rs:18)

class GpuContext { // src/wgpu/mod.rs:40-49
  public:
    explicit GpuContext(int device);
    ~GpuContext();
    GpuContext(const GpuContext&) = delete;
    GpuContext& operator=(const GpuContext&) = delete;
    fg_ctx* raw() const { return ctx_; }
    int device() const { return device_; }

  private:
    fg_ctx* ctx_ = nullptr;
    int device_ = 0;
};

// process-global cached context (src/wgpu/mod.rs:15, 51-92); Err -> RenderError::Gpu
std::shared_ptr<GpuContext> context(int device = 0);
void invalidate_context(); // src/wgpu/mod.rs:88-92

fg_params build_params_block(const Params& params, const Derived& d); // build_uniforms, src/wgpu/mod.rs:661-692
Plane render_pixelwise_gpu(const GpuContext& ctx, const Plane& lambda, const Params& params, const Derived& d);
Plane render_grainwise_gpu(const GpuContext& ctx, const Plane& lambda, const Params& params, const Derived& d);

} // namespace cuda

// Decoded 8-bit interleaved RGB image (what image::open + to_rgb32f see, src/color.rs:158-213)
struct InputImage {
    size_t width = 0, height = 0;
    std::vector<uint8_t> rgb;
};
struct RgbImage {
    size_t width = 0, height = 0;
    std::vector<uint8_t> rgb;
};

// render_with_input_image (src/lib.rs:78-84, 134-173) with Device::Gpu: host load/normalize/lambda,
// device integrator per plane (batched), host store.  `cancel` mirrors render_with_input_image_cancelable.
std::pair<RgbImage, RenderStats> render_with_input_image(const InputImage& input, const Params& params,
                                                         const volatile int* cancel = nullptr, int device = 0);
// Same result with load/lambda/store fused on the device (u8 over PCIe instead of f32 planes).
std::pair<RgbImage, RenderStats> render_with_input_image_fused(const InputImage& input, const Params& params,
                                                               int device = 0);
RenderStats dry_run_with_input_image(const InputImage& input, const Params& params); // src/lib.rs:100-103

} // namespace film_grain
