// gemm_imad.cu -- CUDA-core (IMAD) version of the matrix cross term
//     C (+)= A0*(B0+B1) + A1*B0   over Z_2^64
// (reference: three Eigen i64 products, aby3/sh3/Sh3Evaluator.cpp:662-665).
// This is the comparison point for the tcgen05 limb GEMM and the path used for
// skinny shapes (GEMV-like, config 3/4) where a tensor-core tile would be empty.
// A 64-bit multiply-accumulate is 1 IMAD.WIDE + 2 IMAD on sm_100.
#include "common.cuh"

namespace aby3cu {
namespace {

constexpr int BM = 128, BN = 64, BK = 16;

// Two passes over K: (A0, B0+B1) then (A1, B0) -- the factored form, 2 products
// instead of the reference's 3.
__global__ void __launch_bounds__(256) k_gemm_imad(const u64* __restrict__ A0, const u64* __restrict__ A1,
                                                   const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                   u64 M, u64 K, u64 N, u64* __restrict__ C, int accumulate) {
    __shared__ __align__(16) u64 As[BK][BM];
    __shared__ __align__(16) u64 Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;       // 16 x 16 threads, 8 rows x 4 cols each
    const u64 tiles_n = (N + BN - 1) / BN, tiles_m = (M + BM - 1) / BM;
    for (u64 tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
        const u64 m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        u64 acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0;
        for (int pass = 0; pass < 2; ++pass) {
            const u64* A = pass ? A1 : A0;
            for (u64 k0 = 0; k0 < K; k0 += BK) {
                // A tile: 128 rows x 16 k; thread loads 8 consecutive k of one row
                {
                    const int r = tid >> 1, kh = (tid & 1) * 8;
                    const u64 gr = m0 + r;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const u64 gk = k0 + kh + j;
                        As[kh + j][r] = (gr < M && gk < K) ? A[gr * K + gk] : 0;
                    }
                }
                // B tile: 16 k x 64 cols; thread loads 4 consecutive cols of one k
                {
                    const int kk = tid >> 4, c = (tid & 15) * 4;
                    const u64 gk = k0 + kk;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const u64 gc = n0 + c + j;
                        u64 v = 0;
                        if (gk < K && gc < N) {
                            v = B0[gk * N + gc];
                            if (pass == 0) v += B1[gk * N + gc];
                        }
                        Bs[kk][c + j] = v;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < BK; ++kk) {
                    u64 a[8], b[4];
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
                }
                __syncthreads();
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u64 gr = m0 + ty * 8 + i;
            if (gr >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const u64 gc = n0 + tx * 4 + j;
                if (gc < N) {
                    u64* dst = C + gr * N + gc;
                    *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
                }
            }
        }
    }
}

// Skinny N (<= 8): one warp per output row, lanes stride over K with coalesced
// 8-byte loads of both A planes; HBM-bound on A (16 B per (m,k)).
template <int NN>
__global__ void __launch_bounds__(256) k_gemm_skinny(const u64* __restrict__ A0, const u64* __restrict__ A1,
                                                     const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                     u64 M, u64 K, u64 N, u64* __restrict__ C, int accumulate) {
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (u64 m = warp; m < M; m += nwarps) {
        u64 acc[NN];
#pragma unroll
        for (int j = 0; j < NN; ++j) acc[j] = 0;
        const u64* a0 = A0 + m * K;
        const u64* a1 = A1 + m * K;
        for (u64 k = lane; k < K; k += 32) {
            const u64 x0 = a0[k], x1 = a1[k];
#pragma unroll
            for (int j = 0; j < NN; ++j) {
                if (j < (int)N) {
                    const u64 b0 = B0[k * N + j], b1 = B1[k * N + j];
                    acc[j] += x0 * (b0 + b1) + x1 * b0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NN; ++j) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < NN; ++j)
                if (j < (int)N) {
                    u64* dst = C + m * N + j;
                    *dst = accumulate ? *dst + acc[j] : acc[j];
                }
        }
    }
}

}  // namespace

int gemm_cross_imad(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                    u64 M, u64 K, u64 N, i64* C, int accumulate) {
    if (N <= 8) {
        const u64 want = (M * 32 + 255) / 256;
        const u64 cap = (u64)ctx->sm_count * 8;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
#define ABY3CU_SKINNY(NN) k_gemm_skinny<NN><<<grid, 256, 0, ctx->stream>>>((const u64*)A0, (const u64*)A1, (const u64*)B0, (const u64*)B1, M, K, N, (u64*)C, accumulate)
        if (N == 1) ABY3CU_SKINNY(1); else if (N == 2) ABY3CU_SKINNY(2); else if (N <= 4) ABY3CU_SKINNY(4); else ABY3CU_SKINNY(8);
#undef ABY3CU_SKINNY
        return post_launch(ctx, "k_gemm_skinny");
    }
    const u64 tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const u64 cap = (u64)ctx->sm_count * 2;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    ABY3CU_CHECK(cudaEventRecord(ctx->ev_gemm0, ctx->stream));
    k_gemm_imad<<<grid, 256, 0, ctx->stream>>>((const u64*)A0, (const u64*)A1, (const u64*)B0, (const u64*)B1, M, K, N, (u64*)C, accumulate);
    if (post_launch(ctx, "k_gemm_imad")) return 1;
    ABY3CU_CHECK(cudaEventRecord(ctx->ev_gemm1, ctx->stream));
    return 0;
}

}  // namespace aby3cu
