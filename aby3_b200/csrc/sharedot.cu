// sharedot.cu -- kernels of the 3-party "shared OT" and of the bit x arithmetic
// multiplication built on it (aby3/OT/SharedOT.cpp:6-180, Sh3Evaluator.cpp:119-263,
// 418-501).  Sender and helper share an AES key and a running block counter; message
// pair i is masked with AES_key(toBlock(idx0+i)) (low 8 bytes mask m[i][0], high 8 bytes
// m[i][1]); the helper forwards the half selected by the receiver's choice bit.
#include "aes.cuh"

namespace aby3cu {
namespace {

constexpr int kOtThreads = 256;

__device__ __forceinline__ void ot_pad(u32 lane4, const AesKey& k, u64 ctr, u64& lo, u64& hi) {
    u32 o[4];
    aes_encrypt_ctr(lane4, k, ctr, o);
    lo = ((u64)o[1] << 32) | o[0];
    hi = ((u64)o[3] << 32) | o[2];
}
__device__ __forceinline__ u64 stream_elem(u32 lane4, const AesKey& k, u64 e) {
    u64 lo, hi;
    ot_pad(lane4, k, e >> 1, lo, hi);
    return (e & 1) ? hi : lo;
}

__global__ void __launch_bounds__(kOtThreads) k_ot_send(const __grid_constant__ AesKey key, u64 idx0, const u64* __restrict__ msgs,
                                                        u64* __restrict__ out, size_t n) {
    aes_table_init();
    __syncthreads();
    const u32 lane4 = (threadIdx.x & 31) * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 lo, hi;
        ot_pad(lane4, key, idx0 + i, lo, hi);
        const ulonglong2 m = *reinterpret_cast<const ulonglong2*>(msgs + 2 * i);
        *reinterpret_cast<ulonglong2*>(out + 2 * i) = make_ulonglong2(m.x ^ lo, m.y ^ hi);
    }
}

__global__ void __launch_bounds__(kOtThreads) k_ot_help(const __grid_constant__ AesKey key, u64 idx0, const u64* __restrict__ choice,
                                                        u64* __restrict__ out, size_t n) {
    aes_table_init();
    __syncthreads();
    const u32 lane4 = (threadIdx.x & 31) * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 lo, hi;
        ot_pad(lane4, key, idx0 + i, lo, hi);
        out[i] = (choice[i] & 1) ? hi : lo;
    }
}

__global__ void __launch_bounds__(kOtThreads) k_ot_recv(const u64* __restrict__ masked, const u64* __restrict__ help,
                                                        const u64* __restrict__ choice, u64* __restrict__ out, size_t n, int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const ulonglong2 m = *reinterpret_cast<const ulonglong2*>(masked + 2 * i);
        const u64 v = ((choice[i] & 1) ? m.y : m.x) ^ help[i];
        out[i] = accumulate ? out[i] + v : v;
    }
}

// party 0 of asyncMul(si64Matrix, sbMatrix): Sh3Evaluator.cpp:133-160
__global__ void __launch_bounds__(kOtThreads) k_bitmul_p0(const u64* __restrict__ A0, const u64* __restrict__ A1,
                                                          const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                          const __grid_constant__ AesKey kprev, u64 ep, const __grid_constant__ AesKey knext, u64 en,
                                                          u64* __restrict__ C0, u64* __restrict__ C1, u64* __restrict__ msgs, size_t n) {
    aes_table_init();
    __syncthreads();
    const u32 lane4 = (threadIdx.x & 31) * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 z, c1;
        aes_stream_pair(lane4, kprev, ep + 2 * i, z, c1);        // z then c[1] from mPrevCommon
        const u64 c0 = stream_elem(lane4, knext, en + i);          // c[0] from mNextCommon
        const u64 bb0 = (B0[i] ^ B1[i]) & 1;
        const u64 zz = 0 - (c0 + c1) - z;
        const u64 with = A0[i] + A1[i] + zz;
        C0[i] = c0; C1[i] = c1;
        *reinterpret_cast<ulonglong2*>(msgs + 2 * i) = bb0 ? make_ulonglong2(with, zz) : make_ulonglong2(zz, with);
    }
}

// party 2: Sh3Evaluator.cpp:212-240
__global__ void __launch_bounds__(kOtThreads) k_bitmul_p2(const u64* __restrict__ A1, const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                          const __grid_constant__ AesKey knext, u64 en,
                                                          u64* __restrict__ C0, u64* __restrict__ msgs, size_t n) {
    aes_table_init();
    __syncthreads();
    const u32 lane4 = (threadIdx.x & 31) * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 z, c0;
        aes_stream_pair(lane4, knext, en + 2 * i, z, c0);         // z then c[0] from mNextCommon
        const u64 bb1 = (B0[i] ^ B1[i]) & 1;
        const u64 with = A1[i] + z;
        C0[i] = c0;
        *reinterpret_cast<ulonglong2*>(msgs + 2 * i) = bb1 ? make_ulonglong2(with, z) : make_ulonglong2(z, with);
    }
}

// party 0 of asyncMul(i64, sbMatrix): Sh3Evaluator.cpp:430-447
__global__ void __launch_bounds__(kOtThreads) k_bitmul_pub(u64 a, const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                           const __grid_constant__ AesKey kp, const __grid_constant__ AesKey kn, u64 e0,
                                                           u64* __restrict__ msgs, size_t n) {
    aes_table_init();
    __syncthreads();
    const u32 lane4 = (threadIdx.x & 31) * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 zs = stream_elem(lane4, kp, e0 + i) - stream_elem(lane4, kn, e0 + i);    // getShare()
        const u64 bb = (B0[i] ^ B1[i]) & 1;
        *reinterpret_cast<ulonglong2*>(msgs + 2 * i) = bb ? make_ulonglong2(a + zs, zs) : make_ulonglong2(zs, a + zs);
    }
}

// Sh3Converter::bitInjection, choice vectors (Sh3Converter.cpp:244-247, 282-285): bit j of row i of a
// binary share matrix -> word i*bitCount + j (value 0 / 1), the layout the OT kernels index
__global__ void __launch_bounds__(kOtThreads) k_bits_expand(const u64* __restrict__ in, u64 words, u64 bit_count, u64* __restrict__ out, size_t total) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const u64 i = k / bit_count, j = k - i * bit_count;
        out[k] = (in[i * words + (j >> 6)] >> (j & 63)) & 1;
    }
}

// Sh3Converter::bitInjection, party 2 (sender, Sh3Converter.cpp:318-347): d0 / d1 = the next words of the
// nextCommon / prevCommon streams; m[k][c] = -d0[k] - d1[k] + (c ^ b_k), b_k = bit k of in0 ^ in1
__global__ void __launch_bounds__(kOtThreads) k_bitinj_msgs(const u64* __restrict__ in0, const u64* __restrict__ in1, u64 words, u64 bit_count,
                                                            const __grid_constant__ AesKey knext, u64 en, const __grid_constant__ AesKey kprev, u64 ep,
                                                            u64* __restrict__ d0, u64* __restrict__ d1, u64* __restrict__ msgs, size_t total) {
    aes_table_init();
    __syncthreads();
    const u32 lane4 = (threadIdx.x & 31) * 4;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const u64 i = k / bit_count, j = k - i * bit_count;
        const u64 w = i * words + (j >> 6);
        const u64 b = ((in0[w] ^ in1[w]) >> (j & 63)) & 1;
        const u64 x0 = stream_elem(lane4, knext, en + k), x1 = stream_elem(lane4, kprev, ep + k);
        const u64 base = 0 - x0 - x1;
        d0[k] = x0; d1[k] = x1;
        *reinterpret_cast<ulonglong2*>(msgs + 2 * k) = b ? make_ulonglong2(base + 1, base) : make_ulonglong2(base, base + 1);
    }
}

template <class K>
int big_smem(K kernel) {
    ABY3CU_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesTableBytes));
    if (prefer_max_smem(kernel)) return 1;
    return 0;
}
inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace aby3cu

using namespace aby3cu;

extern "C" {

int aby3cu_ot_send(aby3cu_ctx* ctx, const u8 key[16], u64 idx0, const i64* d_msgs, i64* d_out, size_t n) {
    ABY3CU_REQUIRE(ctx && key && ((d_msgs && d_out) || !n), "ot_send: null argument");
    ABY3CU_REQUIRE(al16p(d_msgs) && al16p(d_out), "ot_send: message pairs must be 16-byte aligned");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey k; host_expand_key(key, &k);
    if (big_smem(k_ot_send)) return 1;
    k_ot_send<<<ew_grid(ctx, n, kOtThreads, 3), kOtThreads, kAesTableBytes, ctx->stream>>>(k, idx0, (const u64*)d_msgs, (u64*)d_out, n);
    return post_launch(ctx, "k_ot_send");
}

int aby3cu_ot_help(aby3cu_ctx* ctx, const u8 key[16], u64 idx0, const i64* d_choice, i64* d_out, size_t n) {
    ABY3CU_REQUIRE(ctx && key && ((d_choice && d_out) || !n), "ot_help: null argument");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey k; host_expand_key(key, &k);
    if (big_smem(k_ot_help)) return 1;
    k_ot_help<<<ew_grid(ctx, n, kOtThreads, 3), kOtThreads, kAesTableBytes, ctx->stream>>>(k, idx0, (const u64*)d_choice, (u64*)d_out, n);
    return post_launch(ctx, "k_ot_help");
}

int aby3cu_ot_recv(aby3cu_ctx* ctx, const i64* d_masked, const i64* d_help, const i64* d_choice, i64* d_out, size_t n, int accumulate) {
    ABY3CU_REQUIRE(ctx && ((d_masked && d_help && d_choice && d_out) || !n), "ot_recv: null argument");
    ABY3CU_REQUIRE(al16p(d_masked), "ot_recv: message pairs must be 16-byte aligned");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    k_ot_recv<<<ew_grid(ctx, n, kOtThreads, 8), kOtThreads, 0, ctx->stream>>>((const u64*)d_masked, (const u64*)d_help, (const u64*)d_choice,
                                                                            (u64*)d_out, n, accumulate);
    return post_launch(ctx, "k_ot_recv");
}

int aby3cu_bitmul_msgs_p0(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                          const u8 key_prev_common[16], u64 elem_prev, const u8 key_next_common[16], u64 elem_next,
                          i64* C0, i64* C1, i64* d_msgs, size_t n) {
    ABY3CU_REQUIRE(ctx && key_prev_common && key_next_common && ((A0 && A1 && B0 && B1 && C0 && C1 && d_msgs) || !n), "bitmul_msgs_p0: null argument");
    ABY3CU_REQUIRE(al16p(d_msgs), "bitmul_msgs_p0: message pairs must be 16-byte aligned");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey kp, kn; host_expand_key(key_prev_common, &kp); host_expand_key(key_next_common, &kn);
    if (big_smem(k_bitmul_p0)) return 1;
    k_bitmul_p0<<<ew_grid(ctx, n, kOtThreads, 3), kOtThreads, kAesTableBytes, ctx->stream>>>(
        (const u64*)A0, (const u64*)A1, (const u64*)B0, (const u64*)B1, kp, elem_prev, kn, elem_next, (u64*)C0, (u64*)C1, (u64*)d_msgs, n);
    return post_launch(ctx, "k_bitmul_p0");
}

int aby3cu_bitmul_msgs_p2(aby3cu_ctx* ctx, const i64* A1, const i64* B0, const i64* B1, const u8 key_next_common[16], u64 elem_next,
                          i64* C0, i64* d_msgs, size_t n) {
    ABY3CU_REQUIRE(ctx && key_next_common && ((A1 && B0 && B1 && C0 && d_msgs) || !n), "bitmul_msgs_p2: null argument");
    ABY3CU_REQUIRE(al16p(d_msgs), "bitmul_msgs_p2: message pairs must be 16-byte aligned");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey kn; host_expand_key(key_next_common, &kn);
    if (big_smem(k_bitmul_p2)) return 1;
    k_bitmul_p2<<<ew_grid(ctx, n, kOtThreads, 3), kOtThreads, kAesTableBytes, ctx->stream>>>(
        (const u64*)A1, (const u64*)B0, (const u64*)B1, kn, elem_next, (u64*)C0, (u64*)d_msgs, n);
    return post_launch(ctx, "k_bitmul_p2");
}

int aby3cu_bitmul_pub_msgs(aby3cu_ctx* ctx, i64 a, const i64* B0, const i64* B1, const u8 key_prev[16], const u8 key_next[16], u64 elem0,
                           i64* d_msgs, size_t n) {
    ABY3CU_REQUIRE(ctx && key_prev && key_next && ((B0 && B1 && d_msgs) || !n), "bitmul_pub_msgs: null argument");
    ABY3CU_REQUIRE(al16p(d_msgs), "bitmul_pub_msgs: message pairs must be 16-byte aligned");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey kp, kn; host_expand_key(key_prev, &kp); host_expand_key(key_next, &kn);
    if (big_smem(k_bitmul_pub)) return 1;
    k_bitmul_pub<<<ew_grid(ctx, n, kOtThreads, 3), kOtThreads, kAesTableBytes, ctx->stream>>>((u64)a, (const u64*)B0, (const u64*)B1, kp, kn, elem0,
                                                                                             (u64*)d_msgs, n);
    return post_launch(ctx, "k_bitmul_pub");
}

int aby3cu_bits_expand(aby3cu_ctx* ctx, const i64* d_in, u64 rows, u64 words, u64 bit_count, i64* d_out) {
    ABY3CU_REQUIRE(ctx, "bits_expand: null context");
    ABY3CU_REQUIRE(bit_count <= 64 * words, "bits_expand: bit count exceeds the row width");
    const size_t total = (size_t)rows * bit_count;
    if (!total) return 0;
    ABY3CU_REQUIRE(d_in && d_out, "bits_expand: null argument");
    DeviceGuard g(ctx->device);
    k_bits_expand<<<ew_grid(ctx, total, kOtThreads, 8), kOtThreads, 0, ctx->stream>>>((const u64*)d_in, words, bit_count, (u64*)d_out, total);
    return post_launch(ctx, "k_bits_expand");
}

int aby3cu_bitinj_msgs(aby3cu_ctx* ctx, const i64* d_in0, const i64* d_in1, u64 rows, u64 words, u64 bit_count,
                       const u8 key_next_common[16], u64 elem_next, const u8 key_prev_common[16], u64 elem_prev,
                       i64* d_d0, i64* d_d1, i64* d_msgs) {
    ABY3CU_REQUIRE(ctx && key_next_common && key_prev_common, "bitinj_msgs: null argument");
    ABY3CU_REQUIRE(bit_count <= 64 * words, "bitinj_msgs: bit count exceeds the row width");
    const size_t total = (size_t)rows * bit_count;
    if (!total) return 0;
    ABY3CU_REQUIRE(d_in0 && d_in1 && d_d0 && d_d1 && d_msgs, "bitinj_msgs: null argument");
    ABY3CU_REQUIRE(al16p(d_msgs), "bitinj_msgs: message pairs must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    AesKey kn, kp; host_expand_key(key_next_common, &kn); host_expand_key(key_prev_common, &kp);
    if (big_smem(k_bitinj_msgs)) return 1;
    k_bitinj_msgs<<<ew_grid(ctx, total, kOtThreads, 3), kOtThreads, kAesTableBytes, ctx->stream>>>(
        (const u64*)d_in0, (const u64*)d_in1, words, bit_count, kn, elem_next, kp, elem_prev, (u64*)d_d0, (u64*)d_d1, (u64*)d_msgs, total);
    return post_launch(ctx, "k_bitinj_msgs");
}

}  // extern "C"
