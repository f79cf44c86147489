// fg_logf.h -- logf exactly as the host libm computes it, usable on the device.
//
// lambda_plane (src/model.rs:252-265) takes f32::ln of every pixel on the Rust host, i.e. the
// platform libm's logf.  glibc (>= 2.27) and musl (>= 1.1.22) both ship the ARM optimized-routines
// logf: a 16-entry table of (1/c, ln c), a degree-3 polynomial, all in binary64, one final rounding
// to binary32.  This header restates that algorithm with explicit, never-contracted operations.
// tools/check_logf.c compares it with the box's libm over all 2 130 706 432 positive normal floats
// (0 differences on glibc 2.39, with or without FMA contraction), and tests/test_host_mirror.py
// spot-checks it on every run.  Inputs on this path are in [1e-6, 1]; zero/negative/inf/NaN and
// subnormals are not handled here (callers clamp first, exactly like the reference).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#else
#define FG_HD inline
#endif

namespace fg {

struct LogfEntry { double invc, logc; };

#define FG_LOGF_TABLE_INIT                                                                          \
    {{0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2}, \
     {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3}, \
     {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},    \
     {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4}, \
     {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},                               \
     {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},   \
     {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},   \
     {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}}

static const LogfEntry kLogfTabHost[16] = FG_LOGF_TABLE_INIT;
#if defined(__CUDACC__)
__constant__ LogfEntry kLogfTabDev[16] = FG_LOGF_TABLE_INIT;
#endif

FG_HD float logf_libm(float x) {
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    const double Ln2 = 0x1.62e42fefa39efp-1;
    uint32_t ix;
#if defined(__CUDA_ARCH__)
    ix = __float_as_uint(x);
#else
    __builtin_memcpy(&ix, &x, 4);
#endif
    if (ix == 0x3f800000u) return 0.0f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    float zf;
#if defined(__CUDA_ARCH__)
    zf = __uint_as_float(iz);
    const double invc = kLogfTabDev[i].invc, logc = kLogfTabDev[i].logc;
    const double z = (double)zf;
    const double r = __dadd_rn(__dmul_rn(z, invc), -1.0);
    const double y0 = __dadd_rn(logc, __dmul_rn((double)k, Ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
#else
    __builtin_memcpy(&zf, &iz, 4);
    const double invc = kLogfTabHost[i].invc, logc = kLogfTabHost[i].logc;
    const double z = (double)zf;
    volatile double t0 = z * invc;          // volatile: keep the host compiler from contracting
    const double r = t0 - 1.0;
    volatile double t1 = (double)k * Ln2;
    const double y0 = logc + t1;
    const double r2 = r * r;
    volatile double t2 = A1 * r;
    double y = t2 + A2;
    volatile double t3 = A0 * r2;
    y = t3 + y;
    volatile double t4 = y * r2;
    y = t4 + (y0 + r);
    return (float)y;
#endif
}

} // namespace fg
