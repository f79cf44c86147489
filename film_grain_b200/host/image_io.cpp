// image_io.cpp -- see image_io.hpp.  PNG per the PNG specification (ISO/IEC 15948): chunk CRCs checked, IDAT stream
// inflated with zlib, the five scanline filters undone, samples expanded to 8-bit RGB.
#include "image_io.hpp"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <sys/stat.h>

#include <zlib.h>

namespace film_grain {
namespace {

[[noreturn]] void fail(const std::string& msg) { throw RenderError(RenderError::Message, msg); }

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) fail("cannot open '" + path + "': " + std::strerror(errno)); // RenderError::Io / Image in the reference
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    const bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) fail("read error on '" + path + "'");
    return buf;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

InputImage decode_png(const uint8_t* d, size_t n, const std::string& what) {
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1;
    bool have_ihdr = false, have_iend = false;
    std::vector<uint8_t> idat, plte;
    while (pos + 12 <= n && !have_iend) {
        const uint32_t len = be32(d + pos);
        if ((size_t)len > n - pos - 12) fail(what + ": truncated PNG chunk");
        const uint8_t* type = d + pos + 4;
        const uint8_t* data = d + pos + 8;
        if (be32(data + len) != (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4)) fail(what + ": PNG chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) fail(what + ": bad IHDR");
            w = be32(data); h = be32(data + 4);
            depth = data[8]; ctype = data[9];
            if (data[10] != 0 || data[11] != 0) fail(what + ": unknown PNG compression / filter method");
            if (data[12] != 0) fail(what + ": interlaced PNG is not supported");
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            have_iend = true;
        } else if (!(type[0] & 0x20)) {
            fail(what + ": unknown critical PNG chunk");
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !have_iend || idat.empty()) fail(what + ": incomplete PNG");
    if (w == 0 || h == 0) fail(what + ": empty image");
    int channels;
    switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: fail(what + ": unknown PNG colour type");
    }
    if (depth == 16) fail(what + ": 16-bit PNG is not supported (the reference keeps 16-bit precision; this path is 8-bit)");
    const bool low = depth == 1 || depth == 2 || depth == 4;
    if (!(depth == 8 || (low && (ctype == 0 || ctype == 3)))) fail(what + ": unsupported PNG bit depth");
    if (ctype == 3 && (plte.empty() || plte.size() % 3)) fail(what + ": palette PNG without a valid PLTE");
    const size_t row_bytes = ((size_t)w * channels * depth + 7) / 8;
    const size_t bpp = std::max<size_t>(1, (size_t)channels * depth / 8);
    if (row_bytes > ((size_t)1 << 31) || (size_t)h > ((size_t)1 << 33) / (row_bytes + 1)) fail(what + ": image too large");
    std::vector<uint8_t> raw((row_bytes + 1) * h);
    {
        z_stream zs{};
        if (inflateInit(&zs) != Z_OK) fail("zlib: inflateInit failed");
        zs.next_in = idat.data();
        zs.avail_in = (uInt)std::min<size_t>(idat.size(), 0xFFFFFFFFu);
        size_t in_done = zs.avail_in, out_done = 0;
        int zr = Z_OK;
        while (zr != Z_STREAM_END) {
            if (zs.avail_in == 0 && in_done < idat.size()) {
                const size_t take = std::min<size_t>(idat.size() - in_done, 0xFFFFFFFFu);
                zs.next_in = idat.data() + in_done;
                zs.avail_in = (uInt)take;
                in_done += take;
            }
            const size_t room = std::min<size_t>(raw.size() - out_done, 0x40000000u);
            zs.next_out = raw.data() + out_done;
            zs.avail_out = (uInt)room;
            zr = inflate(&zs, Z_NO_FLUSH);
            out_done += room - zs.avail_out;
            if (zr != Z_OK && zr != Z_STREAM_END) { inflateEnd(&zs); fail(what + ": corrupt PNG data stream"); }
            if (zr == Z_OK && out_done == raw.size() && zs.avail_out == 0) { // more data than the image holds
                uint8_t extra;
                zs.next_out = &extra; zs.avail_out = 1;
                zr = inflate(&zs, Z_NO_FLUSH);
                if (zr != Z_STREAM_END || zs.avail_out == 0) { inflateEnd(&zs); fail(what + ": PNG data stream longer than the image"); }
            }
        }
        inflateEnd(&zs);
        if (out_done != raw.size()) fail(what + ": PNG data stream shorter than the image");
    }
    // undo the scanline filters in place
    std::vector<uint8_t> zero(row_bytes, 0);
    for (size_t y = 0; y < h; ++y) {
        uint8_t* cur = raw.data() + y * (row_bytes + 1) + 1;
        const uint8_t* up = y ? raw.data() + (y - 1) * (row_bytes + 1) + 1 : zero.data();
        const int ft = cur[-1];
        switch (ft) {
        case 0: break;
        case 1: for (size_t i = bpp; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]); break;
        case 2: for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
        case 3:
            for (size_t i = 0; i < row_bytes; ++i) cur[i] = (uint8_t)(cur[i] + (((i >= bpp ? cur[i - bpp] : 0) + up[i]) >> 1));
            break;
        case 4:
            for (size_t i = 0; i < row_bytes; ++i)
                cur[i] = (uint8_t)(cur[i] + paeth(i >= bpp ? cur[i - bpp] : 0, up[i], i >= bpp ? up[i - bpp] : 0));
            break;
        default: fail(what + ": unknown PNG filter type");
        }
    }
    InputImage img;
    img.width = w; img.height = h;
    img.rgb.resize((size_t)w * h * 3);
    const int maxv = (1 << depth) - 1;
    for (size_t y = 0; y < h; ++y) {
        const uint8_t* row = raw.data() + y * (row_bytes + 1) + 1;
        uint8_t* o = img.rgb.data() + y * (size_t)w * 3;
        for (size_t x = 0; x < w; ++x, o += 3) {
            if (ctype == 2 || ctype == 6) {
                const uint8_t* s = row + x * channels;
                o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
                continue;
            }
            int v;
            if (depth == 8) v = row[x * channels];
            else {
                const size_t bit = x * depth;
                v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
            }
            if (ctype == 3) {
                if ((size_t)v * 3 + 2 >= plte.size()) fail(what + ": palette index out of range");
                o[0] = plte[v * 3]; o[1] = plte[v * 3 + 1]; o[2] = plte[v * 3 + 2];
            } else {
                const uint8_t g = depth == 8 ? (uint8_t)v : (uint8_t)(v * 255 / maxv); // 1 / 2 / 4-bit grey: bit replication
                o[0] = o[1] = o[2] = g;
            }
        }
    }
    return img;
}

// P5 / P6 header: magic, width, height, maxval separated by whitespace, '#' comments, then ONE whitespace byte
InputImage decode_pnm(const uint8_t* d, size_t n, const std::string& what) {
    const bool grey = d[1] == '5';
    size_t pos = 2;
    auto next_int = [&]() -> unsigned long {
        for (;;) {
            while (pos < n && std::isspace(d[pos])) ++pos;
            if (pos < n && d[pos] == '#') { while (pos < n && d[pos] != '\n') ++pos; continue; }
            break;
        }
        if (pos >= n || !std::isdigit(d[pos])) fail(what + ": bad PNM header");
        unsigned long v = 0;
        while (pos < n && std::isdigit(d[pos])) { v = v * 10 + (unsigned long)(d[pos++] - '0'); if (v > (1ul << 31)) fail(what + ": bad PNM header"); }
        return v;
    };
    const unsigned long w = next_int(), h = next_int(), maxv = next_int();
    if (pos >= n || !std::isspace(d[pos])) fail(what + ": bad PNM header");
    ++pos;
    if (w == 0 || h == 0) fail(what + ": empty image");
    if (maxv != 255) fail(what + ": only maxval 255 PNM is supported");
    const size_t need = (size_t)w * h * (grey ? 1 : 3);
    if (n - pos < need) fail(what + ": truncated PNM");
    InputImage img;
    img.width = w; img.height = h;
    img.rgb.resize((size_t)w * h * 3);
    if (!grey) std::memcpy(img.rgb.data(), d + pos, need);
    else
        for (size_t k = 0; k < need; ++k) img.rgb[3 * k] = img.rgb[3 * k + 1] = img.rgb[3 * k + 2] = d[pos + k];
    return img;
}

void png_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
    put_be32(out, (uint32_t)len);
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    put_be32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), out.data() + at, (uInt)(len + 4)));
}

std::vector<uint8_t> encode_png(const uint8_t* rgb, size_t w, size_t h) {
    if (w == 0 || h == 0 || w > 0x7FFFFFFFu || h > 0x7FFFFFFFu) fail("PNG: image dimensions out of range");
    std::vector<uint8_t> out(kPngSig, kPngSig + 8);
    uint8_t ihdr[13];
    std::vector<uint8_t> tmp;
    put_be32(tmp, (uint32_t)w); put_be32(tmp, (uint32_t)h);
    std::memcpy(ihdr, tmp.data(), 8);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    png_chunk(out, "IHDR", ihdr, 13);
    // filter type 0 on every row: film grain is noise, prediction filters buy nothing
    const size_t row = w * 3;
    z_stream zs{};
    if (deflateInit(&zs, 1) != Z_OK) fail("zlib: deflateInit failed");
    std::vector<uint8_t> buf(1 << 20);
    std::vector<uint8_t> line(row + 1);
    for (size_t y = 0; y <= h; ++y) {
        const bool last = y == h;
        if (!last) { line[0] = 0; std::memcpy(line.data() + 1, rgb + y * row, row); }
        zs.next_in = last ? nullptr : line.data();
        zs.avail_in = last ? 0u : (uInt)(row + 1);
        for (;;) {
            zs.next_out = buf.data();
            zs.avail_out = (uInt)buf.size();
            const int zr = deflate(&zs, last ? Z_FINISH : Z_NO_FLUSH);
            if (zr == Z_STREAM_ERROR) { deflateEnd(&zs); fail("zlib: deflate failed"); }
            const size_t got = buf.size() - zs.avail_out;
            if (got) png_chunk(out, "IDAT", buf.data(), got);
            if (last ? zr == Z_STREAM_END : (zs.avail_in == 0 && zs.avail_out != 0)) break;
        }
    }
    deflateEnd(&zs);
    png_chunk(out, "IEND", nullptr, 0);
    return out;
}

std::string lower(std::string s) {
    for (auto& ch : s) ch = (char)std::tolower((unsigned char)ch);
    return s;
}

void create_dir_all(const std::string& dir) { // fs::create_dir_all (src/lib.rs:60-64)
    if (dir.empty()) return;
    std::string cur;
    for (size_t i = 0; i <= dir.size(); ++i) {
        if (i == dir.size() || dir[i] == '/') {
            if (!cur.empty() && cur != "/" && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) fail("cannot create directory '" + cur + "': " + std::strerror(errno));
        }
        if (i < dir.size()) cur.push_back(dir[i]);
    }
}

} // namespace

InputImage decode_image(const uint8_t* bytes, size_t n, const std::string& what) {
    if (n >= 8 && !std::memcmp(bytes, kPngSig, 8)) return decode_png(bytes, n, what);
    if (n >= 2 && bytes[0] == 'P' && (bytes[1] == '5' || bytes[1] == '6')) return decode_pnm(bytes, n, what);
    fail(what + ": unsupported image format (this build reads PNG and binary PNM)");
}

InputImage load_image(const std::string& path) {
    const std::vector<uint8_t> bytes = read_file(path);
    return decode_image(bytes.data(), bytes.size(), path);
}

std::vector<uint8_t> encode_image(const uint8_t* rgb, size_t width, size_t height, ImageFormat format) {
    if (format == ImageFormat::Png) return encode_png(rgb, width, height);
    std::string head = "P6\n" + std::to_string(width) + " " + std::to_string(height) + "\n255\n";
    std::vector<uint8_t> out(head.begin(), head.end());
    out.insert(out.end(), rgb, rgb + width * height * 3);
    return out;
}

void save_image(const std::string& path, const uint8_t* rgb, size_t width, size_t height, ImageFormat format) {
    const std::vector<uint8_t> bytes = encode_image(rgb, width, height, format);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) fail("cannot create '" + path + "': " + std::strerror(errno));
    const bool ok = std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    if (std::fclose(f) != 0 || !ok) fail("write error on '" + path + "'");
}

ImageFormat parse_format_token(const std::string& token) { // src/lib.rs:198-203
    size_t b = 0, e = token.size();
    while (b < e && std::isspace((unsigned char)token[b])) ++b;
    while (e > b && std::isspace((unsigned char)token[e - 1])) --e;
    while (b < e && token[b] == '.') ++b;
    const std::string t = lower(token.substr(b, e - b));
    if (t == "png") return ImageFormat::Png;
    if (t == "ppm" || t == "pnm" || t == "pgm" || t == "pbm" || t == "pam") return ImageFormat::Pnm; // image::ImageFormat::Pnm
    static const char* known[] = {"jpg", "jpeg", "gif", "webp", "tif", "tiff", "tga", "dds", "bmp", "ico", "hdr", "exr", "ff", "avif", "qoi"};
    for (const char* k : known)
        if (t == k) fail("image format '" + t + "' is not built into this engine (PNG and PNM are)");
    fail("unsupported or unknown image format '" + token.substr(b, e - b) + "'");
}

ImageFormat resolve_format(const std::string& output_path, const char* format_token) { // src/lib.rs:188-196
    if (format_token && *format_token) return parse_format_token(format_token);
    const size_t slash = output_path.find_last_of('/');
    const std::string name = slash == std::string::npos ? output_path : output_path.substr(slash + 1);
    const size_t dot = name.find_last_of('.');
    if (dot != std::string::npos && dot > 0) return parse_format_token(name.substr(dot + 1));
    return ImageFormat::Png;
}

InputImage apply_roi(const InputImage& image, const Roi* roi) { // src/color.rs:215-231
    if (!roi) return image;
    if (roi->x1 > image.width || roi->y1 > image.height) fail("ROI exceeds image bounds");
    if (roi->x1 <= roi->x0 || roi->y1 <= roi->y0) fail("ROI width and height must be positive");
    InputImage out;
    out.width = roi->x1 - roi->x0; out.height = roi->y1 - roi->y0;
    out.rgb.resize(out.width * out.height * 3);
    for (size_t y = 0; y < out.height; ++y)
        std::memcpy(out.rgb.data() + y * out.width * 3, image.rgb.data() + ((roi->y0 + y) * image.width + roi->x0) * 3, out.width * 3);
    return out;
}

RenderStats render_file(const Params& params, const std::string& input_path, const std::string& output_path, const char* format_token,
                        const Roi* roi, bool fused, const volatile int* cancel, int device) {
    const InputImage input = apply_roi(load_image(input_path), roi);             // Workspace::load, src/color.rs:26-29
    auto res = fused ? render_with_input_image_fused(input, params, device) : render_with_input_image(input, params, cancel, device);
    const size_t slash = output_path.find_last_of('/');
    if (slash != std::string::npos && slash > 0) create_dir_all(output_path.substr(0, slash));
    const ImageFormat format = resolve_format(output_path, format_token);       // after the render, like the reference
    save_image(output_path, res.first.rgb.data(), res.first.width, res.first.height, format);
    return res.second;
}

} // namespace film_grain
