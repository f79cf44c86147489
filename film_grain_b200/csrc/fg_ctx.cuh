// fg_ctx.cuh -- context object and host-side plumbing shared by fg_api.cu, fg_tile.cuh and
// fg_color.cuh (buffer pools, error mapping).  Host code only.
#pragma once
#include "../../include/fg.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

// ----------------------------------------------------------------------------- context
namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

} // namespace

struct fg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[9] = {}; // [0..6] timings of a render, [7] / [8] start / done markers of a multi-device render
    std::mutex mu;
    std::string err;
    const volatile int* cancel = nullptr;      // fg_set_cancel_flag: stays until replaced
    const volatile int* call_cancel = nullptr; // the flag of the running *_cancelable call (set and cleared under mu)
    fg_stats stats{};
    int sm_count = 0;
    size_t smem_optin = 0;
    // pools
    DevBuf lambda, out, offsets, bits, counts, scan_out, scan_tmp, grains, misc, tiles, thr, bitmap, rowinfo, ptab, gtab, fbtotal, rgb_in, rgb_out, chroma, lut, gw_states;
    bool tables_ready = false;
    uint32_t fb_count_host = 0; // tiled path: fallback-list length of the last render (valid after a stream sync)
    bool fb_pending = false;
    size_t table_max = (size_t)48 << 30; // cell-table budget per band (FG_B200_TABLE_MAX_BYTES overrides; tests)
    double table_slack_sigma = 8.0; // row capacity = expected grains + this many sigma + 64 (FG_B200_TABLE_SLACK_SIGMA: tests)
    const char* eval_kernel = ""; // name of the kernel that evaluated / rasterised the last render (fg_last_eval_kernel)
    uint32_t strip_launches = 0;
    // multi-device context (fg_context_create_multi): this object only routes; subs[g] is a complete context on its own
    // device, subs[0] on the device that owns device-resident inputs and outputs.  peer[g]: device g can store into subs[0]'s memory.
    std::vector<fg_ctx*> subs;
    std::vector<char> peer; // strip-kernel launches of the last pixel-wise render (row sub-bands)
};

namespace {

struct ScopedDevice {
    int prev = -1;
    explicit ScopedDevice(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~ScopedDevice() { if (prev >= 0) cudaSetDevice(prev); }
};

int set_err(fg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int map_cuda_error(fg_ctx* ctx, cudaError_t e, const char* what) {
    cudaGetLastError(); // clear non-sticky state
    std::string msg = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    int code;
    switch (e) {
    case cudaErrorMemoryAllocation: code = FG_ERR_OOM; break;
    case cudaErrorNoDevice:
    case cudaErrorInvalidDevice:
    case cudaErrorInsufficientDriver: code = FG_ERR_NO_DEVICE; break;
    case cudaErrorInvalidValue:
    case cudaErrorInvalidConfiguration:
    case cudaErrorLaunchOutOfResources: code = FG_ERR_CUDA; break;
    default: code = FG_ERR_CUDA_STICKY; break; // illegal address, launch failure, ECC, ...
    }
    return set_err(ctx, code, msg);
}

#define FG_CUDA(ctx, expr)                                              \
    do {                                                                \
        cudaError_t e__ = (expr);                                       \
        if (e__ != cudaSuccess) return map_cuda_error((ctx), e__, #expr); \
    } while (0)

int ensure(fg_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return FG_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256; // a little headroom so sweeps do not realloc every call
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) { b.p = nullptr; return map_cuda_error(ctx, e, "cudaMalloc"); }
    b.cap = want;
    return FG_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

bool cancelled(const fg_ctx* ctx) {
    return (ctx->cancel && *ctx->cancel != 0) || (ctx->call_cancel && *ctx->call_cancel != 0);
}
bool cancel_armed(const fg_ctx* ctx) { return ctx->cancel || ctx->call_cancel; }

// the per-call cancel flag lives in the context only while the call holds the context mutex
struct ScopedCallCancel {
    fg_ctx* ctx;
    ScopedCallCancel(fg_ctx* c, const volatile int* flag) : ctx(c) { ctx->call_cancel = flag; }
    ~ScopedCallCancel() { ctx->call_cancel = nullptr; }
};

} // namespace
