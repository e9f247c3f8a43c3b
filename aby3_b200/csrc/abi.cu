// abi.cu -- context, memory, event and dispatch entry points of include/aby3cu.h,
// plus the small host-side AES used for key draws and key schedules.
#include <stdarg.h>

#include <mutex>

#include "aes.cuh"

namespace aby3cu {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- host AES (FIPS-197), used for key schedules and <= 4 KiB key draws ---------
void host_sbox(u8 sbox[256]) {
    // multiplicative inverse in GF(2^8) via the generator 3, then the affine map
    u8 p = 1, q = 1;
    do {
        p = (u8)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1B : 0));
        q ^= (u8)(q << 1); q ^= (u8)(q << 2); q ^= (u8)(q << 4);
        if (q & 0x80) q ^= 0x09;
        const u8 x = (u8)(q ^ (u8)((q << 1) | (q >> 7)) ^ (u8)((q << 2) | (q >> 6)) ^
                          (u8)((q << 3) | (q >> 5)) ^ (u8)((q << 4) | (q >> 4)));
        sbox[p] = (u8)(x ^ 0x63);
    } while (p != 1);
    sbox[0] = 0x63;
}

static const u8* sbox_table() {
    static u8 sb[256];
    static std::once_flag once;
    std::call_once(once, [] { host_sbox(sb); });
    return sb;
}

static inline u8 xt(u8 x) { return (u8)((x << 1) ^ ((x & 0x80) ? 0x1B : 0)); }

void host_expand_key(const u8 key[16], AesKey* out) {
    const u8* sb = sbox_table();
    u8 rk[11][16];
    memcpy(rk[0], key, 16);
    u8 rcon = 1;
    for (int r = 1; r <= 10; ++r) {
        const u8* pv = rk[r - 1];
        u8 t[4] = {(u8)(sb[pv[13]] ^ rcon), sb[pv[14]], sb[pv[15]], sb[pv[12]]};
        rcon = xt(rcon);
        for (int i = 0; i < 4; ++i) rk[r][i] = (u8)(pv[i] ^ t[i]);
        for (int i = 4; i < 16; ++i) rk[r][i] = (u8)(pv[i] ^ rk[r][i - 4]);
    }
    for (int r = 0; r < 11; ++r)
        for (int c = 0; c < 4; ++c) {
            const u8* b = &rk[r][4 * c];
            out->rk[4 * r + c] = (u32)b[0] | ((u32)b[1] << 8) | ((u32)b[2] << 16) | ((u32)b[3] << 24);
        }
}

void host_encrypt_block(const AesKey& k, const u8 in[16], u8 out[16]) {
    const u8* sb = sbox_table();
    u8 s[16];
    auto add = [&](int r) {
        for (int c = 0; c < 4; ++c) {
            const u32 w = k.rk[4 * r + c];
            s[4 * c] ^= (u8)w; s[4 * c + 1] ^= (u8)(w >> 8); s[4 * c + 2] ^= (u8)(w >> 16); s[4 * c + 3] ^= (u8)(w >> 24);
        }
    };
    memcpy(s, in, 16);
    add(0);
    for (int r = 1; r <= 10; ++r) {
        u8 t[16];
        for (int c = 0; c < 4; ++c)
            for (int row = 0; row < 4; ++row) t[4 * c + row] = sb[s[4 * ((c + row) & 3) + row]];
        if (r < 10) {
            for (int c = 0; c < 4; ++c) {
                const u8 a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = (u8)(xt(a0) ^ xt(a1) ^ a1 ^ a2 ^ a3);
                s[4 * c + 1] = (u8)(a0 ^ xt(a1) ^ xt(a2) ^ a2 ^ a3);
                s[4 * c + 2] = (u8)(a0 ^ a1 ^ xt(a2) ^ xt(a3) ^ a3);
                s[4 * c + 3] = (u8)(xt(a0) ^ a0 ^ a1 ^ a2 ^ xt(a3));
            }
        } else {
            memcpy(s, t, 16);
        }
        add(r);
    }
    memcpy(out, s, 16);
}

int gemm_cross_imad(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                    u64 M, u64 K, u64 N, i64* C, int accumulate);
int gemm_cross_tc(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                  u64 M, u64 K, u64 N, i64* C, int accumulate);
bool gemm_tc_profitable(u64 M, u64 K, u64 N);

static int init_ctx(int device, cudaStream_t stream, bool owns, aby3cu_ctx** out) {
    // more hardware work queues than the default 8: a party keeps four streams busy and three parties may share a GPU
    // (read by the driver when the CUDA context is created; no effect if the host application created it already)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 3;
    }
    ABY3CU_REQUIRE(device >= 0 && device < count, "ctx_create: bad device index");
    DeviceGuard g(device);
    cudaDeviceProp prop;
    ABY3CU_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return 3;
    }
    aby3cu_ctx* c = new aby3cu_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (owns) {
        ABY3CU_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    } else {
        c->stream = stream;
    }
    c->owns_stream = owns;
    if (upload_aes_constants()) { delete c; return 1; }
    ABY3CU_CHECK(cudaEventCreate(&c->ev_gemm0));
    ABY3CU_CHECK(cudaEventCreate(&c->ev_gemm1));
    *out = c;
    return 0;
}

}  // namespace aby3cu

using namespace aby3cu;

extern "C" {

int aby3cu_version(void) { return ABY3CU_VERSION; }
const char* aby3cu_last_error(void) { return g_err; }

int aby3cu_device_count(int* count) {
    ABY3CU_REQUIRE(count, "device_count: null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

int aby3cu_ctx_create(int device, aby3cu_ctx** out) {
    ABY3CU_REQUIRE(out, "ctx_create: null argument");
    return init_ctx(device, nullptr, true, out);
}
int aby3cu_ctx_create_on_stream(int device, void* cuda_stream, aby3cu_ctx** out) {
    ABY3CU_REQUIRE(out, "ctx_create_on_stream: null argument");
    return init_ctx(device, (cudaStream_t)cuda_stream, false, out);
}
int aby3cu_ctx_destroy(aby3cu_ctx* ctx) {
    if (!ctx) return 0;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->gemm_ws.ptr) cudaFree(ctx->gemm_ws.ptr);
    if (ctx->progress_stream) { cudaStreamSynchronize(ctx->progress_stream); cudaStreamDestroy(ctx->progress_stream); }
    if (ctx->progress_reset) cudaEventDestroy(ctx->progress_reset);
    if (ctx->progress) cudaFree(ctx->progress);
    if (ctx->ev_gemm0) cudaEventDestroy(ctx->ev_gemm0);
    if (ctx->ev_gemm1) cudaEventDestroy(ctx->ev_gemm1);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}
int aby3cu_ctx_set_corun(aby3cu_ctx* ctx, int on) {
    ABY3CU_REQUIRE(ctx, "ctx_set_corun: null context");
    ctx->corun = on ? 1 : 0;
    return 0;
}
int aby3cu_trace_begin(aby3cu_ctx* ctx) {
    ABY3CU_REQUIRE(ctx, "trace_begin: null context");
    TraceState& t = trace_state();
    if (!t.on) return 0;
    DeviceGuard g(ctx->device);
    std::lock_guard<std::mutex> l(t.mtx);
    for (auto& r : t.recs) cudaEventDestroy(r.end);
    t.recs.clear();
    if (!t.base) ABY3CU_CHECK(cudaEventCreate(&t.base));
    ABY3CU_CHECK(cudaEventRecord(t.base, ctx->stream));
    return 0;
}
int aby3cu_trace_mark(aby3cu_ctx* ctx, const char* static_name) {
    ABY3CU_REQUIRE(ctx && static_name, "trace_mark: null argument");
    DeviceGuard g(ctx->device);
    trace_mark(ctx, static_name);
    return 0;
}
int aby3cu_trace_dump(const char* path) {
    ABY3CU_REQUIRE(path, "trace_dump: null path");
    TraceState& t = trace_state();
    if (!t.on || !t.base) return 0;
    ABY3CU_CHECK(cudaDeviceSynchronize());
    FILE* f = fopen(path, "w");
    ABY3CU_REQUIRE(f, "trace_dump: cannot open the output file");
    std::lock_guard<std::mutex> l(t.mtx);
    fprintf(f, "stream,kernel,end_ms\n");
    for (auto& r : t.recs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.base, r.end) == cudaSuccess) fprintf(f, "%p,%s,%.4f\n", (void*)r.stream, r.name, ms);
        cudaEventDestroy(r.end);
    }
    t.recs.clear();
    fclose(f);
    return 0;
}
int aby3cu_ctx_device(const aby3cu_ctx* ctx) { return ctx ? ctx->device : -1; }
void* aby3cu_ctx_stream(const aby3cu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t aby3cu_launch_count(const aby3cu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int aby3cu_sync(aby3cu_ctx* ctx) {
    ABY3CU_REQUIRE(ctx, "sync: null context");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int aby3cu_malloc(aby3cu_ctx* ctx, void** d_ptr, size_t bytes) {
    ABY3CU_REQUIRE(ctx && d_ptr, "malloc: null argument");
    DeviceGuard g(ctx->device);
    *d_ptr = nullptr;
    if (!bytes) return 0;
    const cudaError_t e = cudaMalloc(d_ptr, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();        // an allocation failure is not sticky: clear it so that the caller may trim and retry
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
int aby3cu_free(aby3cu_ctx* ctx, void* d_ptr) {
    ABY3CU_REQUIRE(ctx, "free: null context");
    if (!d_ptr) return 0;
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaStreamSynchronize(ctx->stream));
    ABY3CU_CHECK(cudaFree(d_ptr));
    return 0;
}
int aby3cu_memset(aby3cu_ctx* ctx, void* d_ptr, int byte, size_t bytes) {
    ABY3CU_REQUIRE(ctx && (d_ptr || !bytes), "memset: null argument");
    if (!bytes) return 0;
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaMemsetAsync(d_ptr, byte, bytes, ctx->stream));
    return 0;
}
int aby3cu_host_alloc(void** h_ptr, size_t bytes) {
    ABY3CU_REQUIRE(h_ptr, "host_alloc: null argument");
    ABY3CU_CHECK(cudaMallocHost(h_ptr, bytes ? bytes : 1));
    return 0;
}
int aby3cu_host_free(void* h_ptr) {
    if (h_ptr) ABY3CU_CHECK(cudaFreeHost(h_ptr));
    return 0;
}
int aby3cu_h2d(aby3cu_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    ABY3CU_REQUIRE(ctx && ((d_dst && h_src) || !bytes), "h2d: null argument");
    if (!bytes) return 0;
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
int aby3cu_d2h(aby3cu_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    ABY3CU_REQUIRE(ctx && ((h_dst && d_src) || !bytes), "d2h: null argument");
    if (!bytes) return 0;
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
// Direct GPU-to-GPU access over NVLink has to be switched on once per ordered device pair; without it a peer copy is
// staged through host memory (an order of magnitude slower).
static void enable_peer(int from, int to) {
    static std::mutex mtx;
    static bool done[64][64];
    if (from == to || from < 0 || to < 0 || from >= 64 || to >= 64) return;
    std::lock_guard<std::mutex> l(mtx);
    if (done[from][to]) return;
    done[from][to] = true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, from, to) != cudaSuccess || !can) { (void)cudaGetLastError(); return; }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(from);
    const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
    if (e != cudaSuccess) (void)cudaGetLastError();             // already enabled (e.g. by NCCL) is fine
    if (prev >= 0) cudaSetDevice(prev);
}

int aby3cu_d2d(aby3cu_ctx* ctx, void* d_dst, int dst_device, const void* d_src, int src_device, size_t bytes) {
    ABY3CU_REQUIRE(ctx && ((d_dst && d_src) || !bytes), "d2d: null argument");
    if (!bytes) return 0;
    if (dst_device != src_device) { enable_peer(dst_device, src_device); enable_peer(src_device, dst_device); }
    DeviceGuard g(ctx->device);
    if (dst_device == src_device)
        ABY3CU_CHECK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    else
        ABY3CU_CHECK(cudaMemcpyPeerAsync(d_dst, dst_device, d_src, src_device, bytes, ctx->stream));
    return 0;
}

int aby3cu_event_create(aby3cu_ctx* ctx, void** event) {
    ABY3CU_REQUIRE(ctx && event, "event_create: null argument");
    DeviceGuard g(ctx->device);
    cudaEvent_t ev;
    ABY3CU_CHECK(cudaEventCreate(&ev));
    *event = ev;
    return 0;
}
int aby3cu_event_create_sync(aby3cu_ctx* ctx, void** event) {
    ABY3CU_REQUIRE(ctx && event, "event_create_sync: null argument");
    DeviceGuard g(ctx->device);
    cudaEvent_t ev;
    ABY3CU_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    *event = ev;
    return 0;
}
int aby3cu_event_destroy(void* event) {
    if (event) ABY3CU_CHECK(cudaEventDestroy((cudaEvent_t)event));
    return 0;
}
int aby3cu_event_record(aby3cu_ctx* ctx, void* event) {
    ABY3CU_REQUIRE(ctx && event, "event_record: null argument");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaEventRecord((cudaEvent_t)event, ctx->stream));
    return 0;
}
int aby3cu_event_wait(aby3cu_ctx* ctx, void* event) {
    ABY3CU_REQUIRE(ctx && event, "event_wait: null argument");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaStreamWaitEvent(ctx->stream, (cudaEvent_t)event, 0));
    return 0;
}
int aby3cu_event_sync(void* event) {
    ABY3CU_REQUIRE(event, "event_sync: null argument");
    ABY3CU_CHECK(cudaEventSynchronize((cudaEvent_t)event));
    return 0;
}
int aby3cu_event_elapsed_ms(void* start, void* stop, float* ms) {
    ABY3CU_REQUIRE(start && stop && ms, "event_elapsed_ms: null argument");
    ABY3CU_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return 0;
}

// ---- CUDA graphs: capture the work enqueued on this context's stream (and on streams that join it through
// events) and replay it with one launch -----------------------------------------------------------------
int aby3cu_capture_begin(aby3cu_ctx* ctx) {
    ABY3CU_REQUIRE(ctx, "capture_begin: null context");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    return 0;
}
int aby3cu_capture_end(aby3cu_ctx* ctx, void** graph_exec) {
    ABY3CU_REQUIRE(ctx && graph_exec, "capture_end: null argument");
    DeviceGuard g(ctx->device);
    cudaGraph_t graph = nullptr;
    ABY3CU_CHECK(cudaStreamEndCapture(ctx->stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return 1; }
    *graph_exec = exec;
    return 0;
}
int aby3cu_graph_launch(aby3cu_ctx* ctx, void* graph_exec, uint64_t kernel_nodes) {
    ABY3CU_REQUIRE(ctx && graph_exec, "graph_launch: null argument");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx->stream));
    ctx->launches += kernel_nodes;          // the kernels inside the graph still run: keep the launch count honest
    return 0;
}
int aby3cu_graph_destroy(void* graph_exec) {
    if (graph_exec) ABY3CU_CHECK(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return 0;
}

int aby3cu_host_keystream(const u8 key[16], u64 byte_off, size_t nbytes, u8* out) {
    ABY3CU_REQUIRE(key && (out || !nbytes), "host_keystream: null argument");
    ABY3CU_REQUIRE(nbytes <= 4096, "host_keystream: only for small key draws (<= 4096 bytes); use aes_ctr_fill");
    AesKey k;
    host_expand_key(key, &k);
    u64 blk = byte_off / 16;
    size_t skip = (size_t)(byte_off % 16);
    while (nbytes) {
        u8 in[16] = {0}, ct[16];
        memcpy(in, &blk, 8);
        host_encrypt_block(k, in, ct);
        const size_t take = (16 - skip) < nbytes ? (16 - skip) : nbytes;
        memcpy(out, ct + skip, take);
        out += take; nbytes -= take; skip = 0; ++blk;
    }
    return 0;
}

int aby3cu_gemm_cross(aby3cu_ctx* ctx, int algo, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                      u64 M, u64 K, u64 N, i64* C, int accumulate) {
    ABY3CU_REQUIRE(ctx, "gemm_cross: null context");
    ABY3CU_REQUIRE(algo >= 0 && algo <= 2, "gemm_cross: bad algo");
    if (M == 0 || N == 0) return 0;
    ABY3CU_REQUIRE(C, "gemm_cross: null C");
    DeviceGuard g(ctx->device);
    if (K == 0) {
        if (!accumulate) ABY3CU_CHECK(cudaMemsetAsync(C, 0, M * N * 8, ctx->stream));
        return 0;
    }
    ABY3CU_REQUIRE(A0 && A1 && B0 && B1, "gemm_cross: null operand");
    if (algo == ABY3CU_GEMM_AUTO) algo = gemm_tc_profitable(M, K, N) ? ABY3CU_GEMM_TCGEN05 : ABY3CU_GEMM_IMAD;
    ctx->last_gemm_algo = algo;
    if (algo == ABY3CU_GEMM_TCGEN05) return gemm_cross_tc(ctx, A0, A1, B0, B1, M, K, N, C, accumulate);
    if (ctx->c_ready) { ABY3CU_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->c_ready, 0)); ctx->c_ready = nullptr; }
    return gemm_cross_imad(ctx, A0, A1, B0, B1, M, K, N, C, accumulate);
}

int aby3cu_gemm_cross_after(aby3cu_ctx* ctx, int algo, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                            u64 M, u64 K, u64 N, i64* C, int accumulate, void* c_ready) {
    ABY3CU_REQUIRE(ctx, "gemm_cross_after: null context");
    ctx->c_ready = (cudaEvent_t)c_ready;
    if (c_ready && (M == 0 || N == 0 || K == 0)) {
        DeviceGuard g(ctx->device);
        ABY3CU_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->c_ready, 0));
        ctx->c_ready = nullptr;
    }
    const int rc = aby3cu_gemm_cross(ctx, algo, A0, A1, B0, B1, M, K, N, C, accumulate);
    // an early exit (bad alignment, workspace allocation failure ...) must not leave the stream un-ordered behind the
    // producer of C: buffers written there are about to be recycled by the caller's error path
    if (ctx->c_ready) {
        DeviceGuard g(ctx->device);
        cudaStreamWaitEvent(ctx->stream, ctx->c_ready, 0);
        ctx->c_ready = nullptr;
    }
    return rc;
}

int aby3cu_gemm_cross_blocks(aby3cu_ctx* ctx, int algo, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                             u64 M, u64 K, u64 N, i64* C, int accumulate, void* c_ready, u64 block_rows, void** block_events, u32 n_blocks) {
    ABY3CU_REQUIRE(ctx, "gemm_cross_blocks: null context");
    ABY3CU_REQUIRE(block_rows && block_rows % 128 == 0 && block_events && (u64)n_blocks * block_rows >= M && (u64)(n_blocks - 1) * block_rows < M,
                   "gemm_cross_blocks: block_rows must be a multiple of 128 and n_blocks = ceil(M / block_rows)");
    ctx->blk_rows = block_rows;
    ctx->blk_events = (cudaEvent_t*)block_events;
    ctx->blk_n = n_blocks;
    ctx->blk_done = 0;
    const int rc = aby3cu_gemm_cross_after(ctx, algo, A0, A1, B0, B1, M, K, N, C, accumulate, c_ready);
    // whatever ran (another algorithm, an early exit): every block event is recorded, so that nobody waits forever
    {
        DeviceGuard g(ctx->device);
        for (; ctx->blk_done < ctx->blk_n; ++ctx->blk_done) cudaEventRecord(ctx->blk_events[ctx->blk_done], ctx->stream);
    }
    ctx->blk_rows = 0; ctx->blk_events = nullptr; ctx->blk_n = 0; ctx->blk_done = 0;
    return rc;
}

int aby3cu_gemm_last_algo(const aby3cu_ctx* ctx) { return ctx ? ctx->last_gemm_algo : 0; }

int aby3cu_gemm_last_main_kernel_ms(aby3cu_ctx* ctx, float* ms) {
    ABY3CU_REQUIRE(ctx && ms, "gemm_last_main_kernel_ms: null argument");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaEventSynchronize(ctx->ev_gemm1));
    ABY3CU_CHECK(cudaEventElapsedTime(ms, ctx->ev_gemm0, ctx->ev_gemm1));
    return 0;
}

}  // extern "C"
