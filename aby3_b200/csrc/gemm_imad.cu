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

// Skinny N (<= 8): GEMV-like, HBM-bound on A (16 B per (m,k): both share planes are streamed once).
// The K-chunk of B' = B0 + B1 and of B0 is staged in shared memory once per block; one warp per output row,
// every lane streams 16-byte pieces of both A planes (8 independent loads in flight per lane), warp-shuffle reduction.
constexpr int SKINNY_KC = 2048;          // (k, n) pairs of B held in shared memory: 2 x 16 KiB
template <int NN>
__global__ void __launch_bounds__(256) k_gemm_skinny(const u64* __restrict__ A0, const u64* __restrict__ A1,
                                                     const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                     u64 M, u64 K, u64 N, u64 k0, u32 kc, u64* __restrict__ C, int accumulate, int vec) {
    extern __shared__ __align__(16) u64 skinny_smem[];
    u64* sS = skinny_smem;                 // [kc][NN]  B0 + B1
    u64* s0 = skinny_smem + (size_t)kc * NN;   // [kc][NN]  B0
    for (u32 i = threadIdx.x; i < kc * NN; i += blockDim.x) {
        const u32 k = i / NN, j = i - k * NN;
        u64 b0 = 0, b1 = 0;
        if (j < N) { b0 = B0[(k0 + k) * N + j]; b1 = B1[(k0 + k) * N + j]; }
        sS[i] = b0 + b1;
        s0[i] = b0;
    }
    __syncthreads();
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    for (u64 m = warp; m < M; m += nwarps) {
        u64 acc[NN];
#pragma unroll
        for (int j = 0; j < NN; ++j) acc[j] = 0;
        const u64* a0 = A0 + m * K + k0;
        const u64* a1 = A1 + m * K + k0;
        if (vec) {
            // K, k0 even and 16-byte aligned planes: rows start on 16-byte boundaries
#pragma unroll 4
            for (u32 k = 2 * lane; k + 1 < kc; k += 64) {
                const ulonglong2 x0 = __ldcs(reinterpret_cast<const ulonglong2*>(a0 + k));
                const ulonglong2 x1 = __ldcs(reinterpret_cast<const ulonglong2*>(a1 + k));
#pragma unroll
                for (int j = 0; j < NN; ++j)
                    acc[j] += x0.x * sS[k * NN + j] + x1.x * s0[k * NN + j] + x0.y * sS[(k + 1) * NN + j] + x1.y * s0[(k + 1) * NN + j];
            }
            if ((kc & 1) && lane == 0) {
                const u32 k = kc - 1;
#pragma unroll
                for (int j = 0; j < NN; ++j) acc[j] += a0[k] * sS[k * NN + j] + a1[k] * s0[k * NN + j];
            }
        } else {
#pragma unroll 4
            for (u32 k = lane; k < kc; k += 32) {
                const u64 x0 = a0[k], x1 = a1[k];
#pragma unroll
                for (int j = 0; j < NN; ++j) acc[j] += x0 * sS[k * NN + j] + x1 * s0[k * NN + j];
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
        const int NN = N == 1 ? 1 : (N == 2 ? 2 : (N <= 4 ? 4 : 8));
        const u32 kc_max = SKINNY_KC / NN;
        const u64 want = (M * 32 + 255) / 256;
        const u64 cap = (u64)ctx->sm_count * 6;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        const int vec = (K % 2 == 0) && ((reinterpret_cast<uintptr_t>(A0) | reinterpret_cast<uintptr_t>(A1)) & 15) == 0;
        for (u64 k0 = 0; k0 < K; k0 += kc_max) {
            const u32 kc = (u32)(K - k0 < kc_max ? K - k0 : kc_max);
            const size_t smem = (size_t)2 * kc * NN * sizeof(u64);
            const int acc = accumulate || k0 > 0;
#define ABY3CU_SKINNY(N_) k_gemm_skinny<N_><<<grid, 256, smem, ctx->stream>>>((const u64*)A0, (const u64*)A1, (const u64*)B0, (const u64*)B1, M, K, N, k0, kc, (u64*)C, acc, vec)
            if (NN == 1) ABY3CU_SKINNY(1); else if (NN == 2) ABY3CU_SKINNY(2); else if (NN == 4) ABY3CU_SKINNY(4); else ABY3CU_SKINNY(8);
#undef ABY3CU_SKINNY
            if (post_launch(ctx, "k_gemm_skinny")) return 1;
        }
        return 0;
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
