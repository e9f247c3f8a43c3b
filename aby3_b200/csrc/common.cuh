// common.cuh -- context object, error plumbing and launch helpers shared by the
// kernels behind include/aby3cu.h.  sm_100a only; no CPU fallback anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/aby3cu.h"

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

namespace aby3cu {

// B200: 148 SMs.  Grids of the HBM-bound kernels are sized as a multiple of the
// SM count the device reports (queried once per context).
struct GemmWorkspace {
    void* ptr = nullptr;
    size_t bytes = 0;
};

}  // namespace aby3cu

struct aby3cu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 148;
    u64 launches = 0;
    int last_gemm_algo = 0;
    int corun = 0;                       // aby3cu_ctx_set_corun: AES kernels of this context are shaped to fit next to a running GEMM
    cudaEvent_t c_ready = nullptr;       // aby3cu_gemm_cross_after: waited for before the first kernel that touches C
    // aby3cu_gemm_cross_blocks: rows per output block and the events recorded when each block of C is final
    u64 blk_rows = 0;
    cudaEvent_t* blk_events = nullptr;
    u32 blk_n = 0, blk_done = 0;
    // one launch, many blocks: the GEMM counts finished tiles per raster group in `progress` (device memory) and a helper
    // stream turns "group complete" into the caller's block events (cuStreamWaitValue32)
    u32* progress = nullptr;
    cudaStream_t progress_stream = nullptr;
    cudaEvent_t progress_reset = nullptr;
    aby3cu::GemmWorkspace gemm_ws;   // limb planes for the tcgen05 GEMM
    cudaEvent_t ev_gemm0 = nullptr, ev_gemm1 = nullptr;   // bracket the main GEMM kernel
};

namespace aby3cu {

void set_error(const char* fmt, ...);

#define ABY3CU_CHECK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            aby3cu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                              __FILE__, __LINE__);                                      \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

#define ABY3CU_REQUIRE(cond, msg)                                                       \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            aby3cu::set_error("%s (%s:%d)", msg, __FILE__, __LINE__);                   \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        want = dev;
    }
    ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
    int want = -1;
};

// ABY3CU_TRACE=1: every launch leaves an event behind it on its stream; aby3cu_trace_dump() writes "stream, kernel, ms since
// aby3cu_trace_begin" -- the END time of every kernel on every stream of the device, enough to reconstruct how the streams of
// the three parties interleave (there is no nsys in this image).  Off: one predictable branch per launch.
struct TraceRec { cudaStream_t stream; const char* name; cudaEvent_t end; };
struct TraceState {
    bool on = false;
    std::mutex mtx;
    cudaEvent_t base = nullptr;
    std::vector<TraceRec> recs;
};
inline TraceState& trace_state() {
    static TraceState* t = [] { auto* s = new TraceState; const char* e = getenv("ABY3CU_TRACE"); s->on = e && e[0] == '1'; return s; }();
    return *t;
}

// a mark on the stream BEFORE a launch: its time is when the stream became ready for the kernel (all waits satisfied)
inline void trace_mark(aby3cu_ctx* ctx, const char* name) {
    TraceState& t = trace_state();
    if (t.on && t.base) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) == cudaSuccess && cudaEventRecord(ev, ctx->stream) == cudaSuccess) {
            std::lock_guard<std::mutex> g(t.mtx);
            t.recs.push_back(TraceRec{ctx->stream, name, ev});
        }
    }
}

inline int post_launch(aby3cu_ctx* ctx, const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", name, cudaGetErrorString(e));
        return 1;
    }
    ctx->launches++;
    TraceState& t = trace_state();
    if (t.on && t.base) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) == cudaSuccess && cudaEventRecord(ev, ctx->stream) == cudaSuccess) {
            std::lock_guard<std::mutex> g(t.mtx);
            t.recs.push_back(TraceRec{ctx->stream, name, ev});
        }
    }
    return 0;
}

// Kernels that are meant to share an SM with another party's kernel (the 64 KiB T-table AES kernels next to the
// 145 KiB tcgen05 GEMM) ask for the largest shared-memory carve-out: with the default preference the driver sizes the
// carve-out for the kernel that got there first and the second one has to wait for the SM to drain.
// ABY3CU_NO_CARVEOUT=1 keeps the driver's default (for A/B measurements).
template <class K>
inline int prefer_max_smem(K kernel) {
    static const bool off = [] { const char* e = getenv("ABY3CU_NO_CARVEOUT"); return e && e[0] == '1'; }();
    if (off) return 0;
    ABY3CU_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    return 0;
}

// grid for a grid-stride elementwise kernel: enough CTAs to fill the machine a
// few times over, never more than the work needs.
inline unsigned ew_grid(const aby3cu_ctx* ctx, u64 work_items, unsigned threads, unsigned ctas_per_sm) {
    u64 need = (work_items + threads - 1) / threads;
    u64 cap = (u64)ctx->sm_count * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

}  // namespace aby3cu
