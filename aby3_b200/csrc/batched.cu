// batched.cu -- the small kernels of a latency-bound protocol loop, BATCHED over several independent problems per
// launch (blockIdx.y = problem).  When the three parties of an SGD iteration (aby3-ML/Regression.h:142-171) share a
// GPU, their kernels of one protocol step are the same kernel on three sets of pointers: one launch instead of three
// (or six, counting both share planes / both truncation pairs) shortens the replayed graph from 29 to 10 nodes.
// Arithmetic, keystream offsets and results are those of the single-problem kernels in elementwise.cu / gemm_imad.cu.
#include "aes.cuh"

namespace aby3cu {
namespace {

struct TruncJob {
    AesKey kn, kp;
    u64 en, ep;              // first stream elements at iteration 0
    u64 n;
    unsigned d2;             // shift + 2
    i64 *negr, *rt0, *rt1;
};
struct TruncBatch { TruncJob j[ABY3CU_MAX_BATCH]; };

// getTruncationTuple for every job: NEGR = -(t0 >> 2), RT0 = t0 >> d2, RT1 = t1 >> d2  (Sh3Evaluator.cpp:526-537)
__global__ void __launch_bounds__(kThreads) k_trunc_batch(const __grid_constant__ TruncBatch b, const u64* __restrict__ iter, u64 iter_stride) {
    aes_table_init();
    __syncthreads();
    const TruncJob& J = b.j[blockIdx.y];
    const u64 it = iter ? *iter : 0;
    const u64 en = J.en + it * iter_stride, ep = J.ep + it * iter_stride;
    const u32 Tl = (threadIdx.x & 31) * 4;
    const size_t n = J.n, pairs = (n + 1) / 2;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 2 * p;
        u64 t00, t01, t10, t11;
        aes_stream_pair(Tl, J.kn, en + i, t00, t01);
        aes_stream_pair(Tl, J.kp, ep + i, t10, t11);
        U64x2 negr = {0 - sar(t00, 2), 0 - sar(t01, 2)};
        U64x2 rt0 = {sar(t00, J.d2), sar(t01, J.d2)};
        U64x2 rt1 = {sar(t10, J.d2), sar(t11, J.d2)};
        st2(J.negr, i, n, false, negr);
        st2(J.rt0, i, n, false, rt0);
        st2(J.rt1, i, n, false, rt1);
    }
}

struct PlaneBatch { const i64* x[ABY3CU_MAX_BATCH]; const i64* y[ABY3CU_MAX_BATCH]; i64* out[ABY3CU_MAX_BATCH]; };
template <int OP>
__global__ void __launch_bounds__(kThreads) k_share_op_batch(PlaneBatch b, size_t n) {
    const i64* __restrict__ x = b.x[blockIdx.y];
    const i64* __restrict__ y = b.y[blockIdx.y];
    i64* __restrict__ out = b.out[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 a = (u64)x[i], c = (u64)y[i];
        out[i] = (i64)(OP == ABY3CU_OP_ADD ? a + c : OP == ABY3CU_OP_SUB ? a - c : a ^ c);
    }
}

struct TransposeBatch { const i64* in[ABY3CU_MAX_BATCH]; i64* out[ABY3CU_MAX_BATCH]; };
__global__ void __launch_bounds__(256) k_transpose_batch(TransposeBatch b, u64 rows, u64 cols) {
    __shared__ i64 tile[32][33];
    const i64* __restrict__ in = b.in[blockIdx.y];
    i64* __restrict__ out = b.out[blockIdx.y];
    const u64 tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    for (u64 t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
        const u64 tr = t / tiles_c, tc = t % tiles_c;
        const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
        for (int j = ly; j < 32; j += 8) {
            u64 r = tr * 32 + j, c = tc * 32 + lx;
            if (r < rows && c < cols) tile[j][lx] = in[r * cols + c];
        }
        __syncthreads();
        for (int j = ly; j < 32; j += 8) {
            u64 c = tc * 32 + j, r = tr * 32 + lx;
            if (r < rows && c < cols) out[c * rows + r] = tile[lx][j];
        }
        __syncthreads();
    }
}

struct FinishBatch { const i64* s0[ABY3CU_MAX_BATCH]; const i64* s1[ABY3CU_MAX_BATCH]; const i64* s2[ABY3CU_MAX_BATCH]; i64* C[ABY3CU_MAX_BATCH]; };
// C += (s0 + s1 + s2) >> shift   (Sh3Evaluator.cpp:712-718)
__global__ void __launch_bounds__(kThreads) k_trunc_finish_batch(FinishBatch b, size_t n, unsigned shift) {
    const i64* __restrict__ s0 = b.s0[blockIdx.y];
    const i64* __restrict__ s1 = b.s1[blockIdx.y];
    const i64* __restrict__ s2 = b.s2[blockIdx.y];
    i64* __restrict__ C = b.C[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        C[i] = (i64)((u64)C[i] + sar((u64)s0[i] + (u64)s1[i] + (u64)s2[i], shift));
}

// The cross terms of ALL THREE co-located parties of one GEMV-shaped product (N = 1), each share plane of A read ONCE.
// Party p computes A0_p (B0_p + B1_p) + A1_p B0_p and its A1_p is the previous party's A0 (replicated sharing), so plane
// A0_p serves two parties: p (against B0_p + B1_p) and p + 1 (against B0_{p+1}).  A warp owns a row: three 16-byte loads
// per lane and k-step, six multiply-adds.  Logistic inference at 2^22 x 512 reads 48 GiB instead of 96 GiB.
struct GemvRing { const u64* A0[3]; const u64* B0[3]; const u64* B1[3]; u64* C[3]; };
constexpr u32 kRingKc = 1024;            // contraction chunk: 6 vectors x 1024 x 8 B = 48 KiB of shared memory
__global__ void __launch_bounds__(256) k_gemv_ring(GemvRing g, u64 M, u64 K, u64 k0, u32 kc, int accumulate, int vec) {
    extern __shared__ u64 ring_smem[];
    u64* S[3]; u64* Z[3];
    for (int p = 0; p < 3; ++p) { S[p] = ring_smem + (size_t)(2 * p) * kc; Z[p] = ring_smem + (size_t)(2 * p + 1) * kc; }
    for (u32 i = threadIdx.x; i < kc; i += blockDim.x)
        for (int p = 0; p < 3; ++p) {
            const u64 b0 = g.B0[p][k0 + i], b1 = g.B1[p][k0 + i];
            S[p][i] = b0 + b1;
            Z[p][i] = b0;
        }
    __syncthreads();
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    for (u64 m = warp; m < M; m += nwarps) {
        u64 acc[3] = {0, 0, 0};
        const u64* a[3] = {g.A0[0] + m * K + k0, g.A0[1] + m * K + k0, g.A0[2] + m * K + k0};
        if (vec) {
#pragma unroll 2
            for (u32 k = 2 * lane; k + 1 < kc; k += 64) {
                ulonglong2 x[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) x[p] = __ldcs(reinterpret_cast<const ulonglong2*>(a[p] + k));
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const int q = (p + 1) % 3;
                    acc[p] += x[p].x * S[p][k] + x[p].y * S[p][k + 1];
                    acc[q] += x[p].x * Z[q][k] + x[p].y * Z[q][k + 1];
                }
            }
            if ((kc & 1) && lane == 0) {
                const u32 k = kc - 1;
                for (int p = 0; p < 3; ++p) { const int q = (p + 1) % 3; const u64 x = a[p][k]; acc[p] += x * S[p][k]; acc[q] += x * Z[q][k]; }
            }
        } else {
            for (u32 k = lane; k < kc; k += 32)
                for (int p = 0; p < 3; ++p) { const int q = (p + 1) % 3; const u64 x = a[p][k]; acc[p] += x * S[p][k]; acc[q] += x * Z[q][k]; }
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], off);
        }
        if (lane < 3) {
            const u64 v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
            u64* dst = (lane == 0 ? g.C[0] : (lane == 1 ? g.C[1] : g.C[2])) + m;
            *dst = accumulate ? *dst + v : v;
        }
    }
}

struct GemvBatch {
    const u64* A0[ABY3CU_MAX_BATCH]; const u64* A1[ABY3CU_MAX_BATCH];
    const u64* B0[ABY3CU_MAX_BATCH]; const u64* B1[ABY3CU_MAX_BATCH];
    u64* C[ABY3CU_MAX_BATCH];
};
// C[m] (+)= sum_k A0[m,k] (B0[k] + B1[k]) + A1[m,k] B0[k]   (N = 1; one warp per row, B staged in shared memory in chunks of
// GEMV_KC; a block keeps the partial sums of its rows in registers across chunks)
constexpr u32 GEMV_KC = 2048;
constexpr int GEMV_ROWS_PER_WARP = 4;
__global__ void __launch_bounds__(256) k_gemv_batch(GemvBatch b, u64 M, u64 K, int accumulate) {
    __shared__ u64 sS[GEMV_KC], s0[GEMV_KC];
    const u64* __restrict__ A0 = b.A0[blockIdx.y];
    const u64* __restrict__ A1 = b.A1[blockIdx.y];
    const u64* __restrict__ B0 = b.B0[blockIdx.y];
    const u64* __restrict__ B1 = b.B1[blockIdx.y];
    u64* __restrict__ C = b.C[blockIdx.y];
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    // every warp owns up to GEMV_ROWS_PER_WARP rows per pass over K (rows warp, warp + nwarps, ...)
    for (u64 mbase = 0; mbase < M; mbase += nwarps * GEMV_ROWS_PER_WARP) {
        u64 acc[GEMV_ROWS_PER_WARP];
#pragma unroll
        for (int r = 0; r < GEMV_ROWS_PER_WARP; ++r) acc[r] = 0;
        for (u64 k0 = 0; k0 < K; k0 += GEMV_KC) {
            const u32 kc = (u32)(K - k0 < GEMV_KC ? K - k0 : GEMV_KC);
            __syncthreads();
            for (u32 k = threadIdx.x; k < kc; k += blockDim.x) { const u64 x = B0[k0 + k]; sS[k] = x + B1[k0 + k]; s0[k] = x; }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < GEMV_ROWS_PER_WARP; ++r) {
                const u64 m = mbase + warp + (u64)r * nwarps;
                if (m >= M) continue;
                const u64* a0 = A0 + m * K + k0;
                const u64* a1 = A1 + m * K + k0;
#pragma unroll 4
                for (u32 k = lane; k < kc; k += 32) acc[r] += a0[k] * sS[k] + a1[k] * s0[k];
            }
        }
#pragma unroll
        for (int r = 0; r < GEMV_ROWS_PER_WARP; ++r) {
            u64 v = acc[r];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            const u64 m = mbase + warp + (u64)r * nwarps;
            if (lane == 0 && m < M) C[m] = accumulate ? C[m] + v : v;
        }
    }
}

}  // namespace
}  // namespace aby3cu

using namespace aby3cu;

extern "C" {

int aby3cu_trunc_tuple_batch_at(aby3cu_ctx* ctx, int njobs, const u8* const* keys_next, const u64* elem_next, const u8* const* keys_prev,
                                const u64* elem_prev, const u64* shifts, const u64* counts, i64* const* negr, i64* const* rt0,
                                i64* const* rt1, const u64* d_iter, u64 iter_stride) {
    ABY3CU_REQUIRE(ctx && keys_next && elem_next && keys_prev && elem_prev && shifts && counts && negr && rt0 && rt1, "trunc_tuple_batch_at: null argument");
    ABY3CU_REQUIRE(njobs >= 1 && njobs <= ABY3CU_MAX_BATCH, "trunc_tuple_batch_at: bad job count");
    TruncBatch b;
    u64 maxn = 0;
    for (int j = 0; j < ABY3CU_MAX_BATCH; ++j) {
        const int k = j < njobs ? j : 0;
        ABY3CU_REQUIRE(keys_next[k] && keys_prev[k] && negr[k] && rt0[k] && rt1[k] && shifts[k] + 2 < 64, "trunc_tuple_batch_at: bad job");
        host_expand_key(keys_next[k], &b.j[j].kn);
        host_expand_key(keys_prev[k], &b.j[j].kp);
        b.j[j].en = elem_next[k]; b.j[j].ep = elem_prev[k]; b.j[j].n = counts[k]; b.j[j].d2 = (unsigned)shifts[k] + 2;
        b.j[j].negr = negr[k]; b.j[j].rt0 = rt0[k]; b.j[j].rt1 = rt1[k];
        if (counts[k] > maxn) maxn = counts[k];
    }
    if (!maxn) return 0;
    DeviceGuard g(ctx->device);
    if (enable_big_smem(k_trunc_batch)) return 1;
    const unsigned gx = ew_grid(ctx, (maxn + 1) / 2, kThreads, 1);
    k_trunc_batch<<<dim3(gx, (unsigned)njobs), kThreads, kAesTableBytes, ctx->stream>>>(b, d_iter, iter_stride);
    return post_launch(ctx, "k_trunc_batch");
}

int aby3cu_share_op_batch(aby3cu_ctx* ctx, int op, int nplanes, const i64* const* x, const i64* const* y, i64* const* out, size_t n) {
    ABY3CU_REQUIRE(ctx && x && y && out, "share_op_batch: null argument");
    ABY3CU_REQUIRE(op >= 0 && op <= 2 && nplanes >= 1 && nplanes <= ABY3CU_MAX_BATCH, "share_op_batch: bad argument");
    if (!n) return 0;
    PlaneBatch b;
    for (int j = 0; j < ABY3CU_MAX_BATCH; ++j) {
        const int k = j < nplanes ? j : 0;
        ABY3CU_REQUIRE(x[k] && y[k] && out[k], "share_op_batch: null plane");
        b.x[j] = x[k]; b.y[j] = y[k]; b.out[j] = out[k];
    }
    DeviceGuard g(ctx->device);
    const dim3 grid(ew_grid(ctx, n, kThreads, 2), (unsigned)nplanes);
    if (op == ABY3CU_OP_ADD) k_share_op_batch<ABY3CU_OP_ADD><<<grid, kThreads, 0, ctx->stream>>>(b, n);
    else if (op == ABY3CU_OP_SUB) k_share_op_batch<ABY3CU_OP_SUB><<<grid, kThreads, 0, ctx->stream>>>(b, n);
    else k_share_op_batch<ABY3CU_OP_XOR><<<grid, kThreads, 0, ctx->stream>>>(b, n);
    return post_launch(ctx, "k_share_op_batch");
}

int aby3cu_transpose_i64_batch(aby3cu_ctx* ctx, int nplanes, const i64* const* in, u64 rows, u64 cols, i64* const* out) {
    ABY3CU_REQUIRE(ctx && in && out, "transpose_batch: null argument");
    ABY3CU_REQUIRE(nplanes >= 1 && nplanes <= ABY3CU_MAX_BATCH, "transpose_batch: bad plane count");
    if (!(rows * cols)) return 0;
    TransposeBatch b;
    for (int j = 0; j < ABY3CU_MAX_BATCH; ++j) {
        const int k = j < nplanes ? j : 0;
        ABY3CU_REQUIRE(in[k] && out[k], "transpose_batch: null plane");
        b.in[j] = in[k]; b.out[j] = out[k];
    }
    DeviceGuard g(ctx->device);
    const u64 tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    const unsigned gx = (unsigned)(tiles < (u64)ctx->sm_count * 2 ? tiles : (u64)ctx->sm_count * 2);
    k_transpose_batch<<<dim3(gx, (unsigned)nplanes), 256, 0, ctx->stream>>>(b, rows, cols);
    return post_launch(ctx, "k_transpose_batch");
}

int aby3cu_trunc_finish_batch(aby3cu_ctx* ctx, int njobs, const i64* const* s0, const i64* const* s1, const i64* const* s2, i64* const* C,
                              size_t n, u64 shift) {
    ABY3CU_REQUIRE(ctx && s0 && s1 && s2 && C, "trunc_finish_batch: null argument");
    ABY3CU_REQUIRE(njobs >= 1 && njobs <= ABY3CU_MAX_BATCH && shift < 64, "trunc_finish_batch: bad argument");
    if (!n) return 0;
    FinishBatch b;
    for (int j = 0; j < ABY3CU_MAX_BATCH; ++j) {
        const int k = j < njobs ? j : 0;
        ABY3CU_REQUIRE(s0[k] && s1[k] && s2[k] && C[k], "trunc_finish_batch: null job");
        b.s0[j] = s0[k]; b.s1[j] = s1[k]; b.s2[j] = s2[k]; b.C[j] = C[k];
    }
    DeviceGuard g(ctx->device);
    k_trunc_finish_batch<<<dim3(ew_grid(ctx, n, kThreads, 2), (unsigned)njobs), kThreads, 0, ctx->stream>>>(b, n, (unsigned)shift);
    return post_launch(ctx, "k_trunc_finish_batch");
}

int aby3cu_gemv_cross_batch(aby3cu_ctx* ctx, int njobs, const i64* const* A0, const i64* const* A1, const i64* const* B0,
                            const i64* const* B1, u64 M, u64 K, i64* const* C, int accumulate) {
    ABY3CU_REQUIRE(ctx && A0 && A1 && B0 && B1 && C, "gemv_cross_batch: null argument");
    ABY3CU_REQUIRE(njobs >= 1 && njobs <= ABY3CU_MAX_BATCH, "gemv_cross_batch: bad job count");
    if (!M) return 0;
    ABY3CU_REQUIRE(K >= 1, "gemv_cross_batch: empty contraction");
    GemvBatch b;
    for (int j = 0; j < ABY3CU_MAX_BATCH; ++j) {
        const int k = j < njobs ? j : 0;
        ABY3CU_REQUIRE(A0[k] && A1[k] && B0[k] && B1[k] && C[k], "gemv_cross_batch: null job");
        b.A0[j] = (const u64*)A0[k]; b.A1[j] = (const u64*)A1[k]; b.B0[j] = (const u64*)B0[k]; b.B1[j] = (const u64*)B1[k]; b.C[j] = (u64*)C[k];
    }
    DeviceGuard g(ctx->device);
    const u64 want = (M * 32 + 255) / 256, cap = (u64)ctx->sm_count * 2;
    const unsigned gx = (unsigned)(want < cap ? want : cap);
    k_gemv_batch<<<dim3(gx, (unsigned)njobs), 256, 0, ctx->stream>>>(b, M, K, accumulate);
    return post_launch(ctx, "k_gemv_batch");
}

int aby3cu_gemv_ring(aby3cu_ctx* ctx, const i64* const* A0, const i64* const* B0, const i64* const* B1, u64 M, u64 K, i64* const* C,
                     int accumulate) {
    ABY3CU_REQUIRE(ctx && A0 && B0 && B1 && C, "gemv_ring: null argument");
    if (!M) return 0;
    ABY3CU_REQUIRE(K >= 1, "gemv_ring: empty contraction");
    GemvRing g;
    uintptr_t align = 0;
    for (int p = 0; p < 3; ++p) {
        ABY3CU_REQUIRE(A0[p] && B0[p] && B1[p] && C[p], "gemv_ring: null party");
        g.A0[p] = (const u64*)A0[p]; g.B0[p] = (const u64*)B0[p]; g.B1[p] = (const u64*)B1[p]; g.C[p] = (u64*)C[p];
        align |= reinterpret_cast<uintptr_t>(A0[p]);
    }
    DeviceGuard dg(ctx->device);
    const int vec = (K % 2 == 0) && (align & 15) == 0;
    const u64 want = (M * 32 + 255) / 256, cap = (u64)ctx->sm_count * 4;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    ABY3CU_CHECK(cudaFuncSetAttribute(k_gemv_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kRingKc * sizeof(u64)));
    for (u64 k0 = 0; k0 < K; k0 += kRingKc) {
        const u32 kc = (u32)(K - k0 < kRingKc ? K - k0 : kRingKc);
        k_gemv_ring<<<grid, 256, (size_t)6 * kc * sizeof(u64), ctx->stream>>>(g, M, K, k0, kc, accumulate || k0 > 0, vec && (k0 % 2 == 0));
        if (post_launch(ctx, "k_gemv_ring")) return 1;
    }
    return 0;
}

}  // extern "C"
