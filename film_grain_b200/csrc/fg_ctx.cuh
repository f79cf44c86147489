// fg_ctx.cuh -- context object and host-side plumbing shared by fg_api.cu, fg_tile.cuh and
// fg_color.cuh (buffer pools, error mapping).  Host code only.
#pragma once
#include "../../include/fg.h"

#include <chrono>
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

// Identity of a cell table (fg_set_table_cache): everything the table's contents depend on.  N, the sample offsets and
// the zoom are NOT part of it -- they only decide which cells a render visits, i.e. the rectangle the table must cover.
struct TableKey {
    uint64_t seed_cell = 0, h0 = 0, h1 = 0;   // h0, h1: 128-bit content hash of the lambda planes
    uint32_t seeding = 0, in_w = 0, in_h = 0, n_planes = 0, lognorm = 0;
    float delta = 0, rm = 0, mean_linear = 0;
    double mu = 0, sigma = 0, slack = 0;
    bool operator==(const TableKey& o) const {
        return seed_cell == o.seed_cell && h0 == o.h0 && h1 == o.h1 && seeding == o.seeding && in_w == o.in_w && in_h == o.in_h &&
               n_planes == o.n_planes && lognorm == o.lognorm && delta == o.delta && rm == o.rm && mean_linear == o.mean_linear &&
               mu == o.mu && sigma == o.sigma && slack == o.slack;
    }
};
struct TableCache {
    bool enabled = false, valid = false;
    TableKey key;
    int i0 = 0, j0 = 0, cols = 0, rows = 0; // cell rectangle of the table held in the ptab / gtab / rowinfo / bitmap / thr pools
    uint64_t total = 0;                     // grain slots (sum of the row capacities)
    double dens = 0.0;
};

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
    DevBuf lambda, out, offsets, bits, counts, scan_out, scan_tmp, grains, misc, tiles, thr, bitmap, rowinfo, ptab, gtab, fbtotal, rgb_in, rgb_out, chroma, lut, gw_states, acc, part;
    bool tables_ready = false;
    // Page-locked landing area of the small device->host reads (a copy into pageable memory would block the host inside
    // cudaMemcpyAsync until the stream gets there, i.e. for the whole kernel in front of it -- no polling, no cancel):
    // [0] table / grain total, [1] overflow flag, [2..3] plane hash, [4] fallback-list length of the last render
    uint64_t* h_pin = nullptr;
    uint32_t& fb_count_host() const { return *(uint32_t*)(h_pin + 4); } // valid after a stream sync
    bool fb_pending = false;
    size_t table_max = (size_t)48 << 30; // cell-table budget per band (FG_B200_TABLE_MAX_BYTES overrides; tests)
    double table_slack_sigma = 8.0; // row capacity = expected grains + this many sigma + 64 (FG_B200_TABLE_SLACK_SIGMA: tests)
    // in-launch cancel: a device word the kernels poll (RenderConsts::abort), raised from `abort_stream` by the host
    // thread that waits for the render (wait_stream) when it sees the caller's cancel flag
    int* d_abort = nullptr;
    int* h_one = nullptr; // page-locked 1: the source of the abort copy
    cudaStream_t abort_stream = nullptr;
    cudaEvent_t ev_wait = nullptr;
    bool abort_sent = false;
    // A whole-frame lambda upload the host entry point has NOT issued yet: the staged pixel-wise pipeline issues it in row
    // chunks on `copy_stream` and runs the thresholds + first-draw bitmap of chunk k while chunk k + 1 is still crossing
    // PCIe (or, for pageable planes, being staged by the host).  Whoever reaches the data first consumes it
    // (flush_upload: plain copies on the main stream).
    struct {
        bool pending = false;
        const float* const* host = nullptr; // the caller's planes
        int n_planes = 0;
        size_t in_w = 0, in_h = 0;
        float* dev = nullptr;               // ctx->lambda
    } up;
    // Pageable output planes: the pipeline evaluates the band in row slices and records an event after each, the host
    // entry point copies slice e on `copy_stream` while slice e + 1 is being evaluated.
    struct {
        bool want = false;
        int n = 0;
        int row_end[4] = {};
        cudaEvent_t ev[4] = {};
    } outp;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up_ev[5] = {};
    TableCache tcache;
    // progressive refinement (fg_refine_planes): `acc` holds the render of samples [0, acc_k) of the geometry below
    struct { bool valid = false; uint32_t k = 0, n_planes = 0, out_w = 0, out_h = 0, row_begin = 0, row_end = 0; int algo = 0; } prog;
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

// Wait for the context's stream.  With a cancel flag armed the host polls instead of blocking: the moment the caller's
// flag goes up, the device word the kernels test is raised from a second stream (a 4-byte copy that overtakes the
// running kernels), so a render stops within one CTA lifetime (~1.5 ms on a 4K frame) instead of at the next call
// boundary.  The caller checks cancelled() afterwards.
cudaError_t wait_stream(fg_ctx* ctx) {
    if (!cancel_armed(ctx) || !ctx->d_abort) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->ev_wait, ctx->stream);
    if (e != cudaSuccess) return e;
    for (;;) {
        e = cudaEventQuery(ctx->ev_wait);
        if (e != cudaErrorNotReady) return e;
        if (!ctx->abort_sent && cancelled(ctx)) {
            cudaMemcpyAsync(ctx->d_abort, ctx->h_one, sizeof(int), cudaMemcpyHostToDevice, ctx->abort_stream);
            ctx->abort_sent = true;
        }
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
}

// issue a still-pending lambda upload in one piece on the main stream (every path but the chunked one below)
cudaError_t flush_upload(fg_ctx* ctx) {
    if (!ctx->up.pending) return cudaSuccess;
    ctx->up.pending = false;
    const size_t elems = ctx->up.in_w * ctx->up.in_h;
    for (int pl = 0; pl < ctx->up.n_planes; ++pl) {
        const cudaError_t e = cudaMemcpyAsync(ctx->up.dev + elems * pl, ctx->up.host[pl], elems * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// the per-call cancel flag lives in the context only while the call holds the context mutex
struct ScopedCallCancel {
    fg_ctx* ctx;
    ScopedCallCancel(fg_ctx* c, const volatile int* flag) : ctx(c) { ctx->call_cancel = flag; }
    ~ScopedCallCancel() { ctx->call_cancel = nullptr; }
};

} // namespace
