// image_io.hpp -- image file I/O of the reference's `render(params)`: `image::open` (src/color.rs:26-29)
// and `save_with_format` (src/lib.rs:60-68) for the formats this engine reads and writes itself:
//   PNG  8-bit (and 1/2/4-bit grey / palette), colour types 0, 2, 3, 4, 6, non-interlaced; alpha is dropped as
//        DynamicImage::to_rgb32f drops it.  Written as 8-bit RGB, zlib level 1 (at B200 speeds the encoder, not the
//        render, is what a caller waits for).
//   PNM  P5 / P6 with maxval 255 (the raw fast path).
// 16-bit sources are refused: the reference keeps their precision (to_rgb32f), the 8-bit pipeline would not.
#pragma once
#include <cstdint>
#include <string>

#include "film_grain.hpp"

namespace film_grain {

enum class ImageFormat { Png, Pnm };

struct Roi { uint32_t x0, y0, x1, y1; }; // src/params.rs:32-37 (exclusive end)

InputImage load_image(const std::string& path);                               // throws RenderError::Message
InputImage decode_image(const uint8_t* bytes, size_t n, const std::string& what);
void save_image(const std::string& path, const uint8_t* rgb, size_t width, size_t height, ImageFormat format);
std::vector<uint8_t> encode_image(const uint8_t* rgb, size_t width, size_t height, ImageFormat format);
ImageFormat parse_format_token(const std::string& token);                     // src/lib.rs:198-203
ImageFormat resolve_format(const std::string& output_path, const char* format_token); // src/lib.rs:188-196
InputImage apply_roi(const InputImage& image, const Roi* roi);                // src/color.rs:215-231

// render(params) (src/lib.rs:57-71): load, crop, render on the device, create the output directory, save.
RenderStats render_file(const Params& params, const std::string& input_path, const std::string& output_path, const char* format_token,
                        const Roi* roi, bool fused, const volatile int* cancel, int device);

} // namespace film_grain
