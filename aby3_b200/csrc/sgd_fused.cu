// sgd_fused.cu -- SGD_Linear (aby3-ML/Regression.h:112-184) for three parties that share ONE GPU, the WHOLE training
// run as ONE persistent kernel.
//
// An iteration at B = 128, F = 1024 touches 6 MiB of X and does a few hundred thousand 64-bit multiplies: as a chain of
// kernels (even replayed as a CUDA graph, ml/SgdGraph.h: 7 dependent nodes) it is bound by launch latency, ~7 us per
// node.  Here one co-resident grid (<= one CTA per SM) walks the iterations itself; the only global synchronisation is
// two grid barriers per iteration:
//
//   phase A  one CTA per batch row b, all three parties: the row of X is read straight from the training matrix (no
//            extractBatch copy),  V1_p[b] = XX_p[b,:] x w_p - r_p[b]   (asyncMul(XX, w, shift D): Sh3Evaluator.cpp:651-700)
//            and  E_p[b] = RTrunc_p[b] - YY_p[b]   (Regression.h:160: error -= YY, applied before the opened value is
//            added: addition commutes mod 2^64).  The truncation pair comes from the common keystreams in registers.
//   -- grid barrier --
//   phase B  every CTA rebuilds the opened product for itself:  E_p[p][b] += (V1_0 + V1_1 + V1_2)[b] >> D  for p in {0,1}
//            (Sh3Evaluator.cpp:703-724), 9 B words from L2 into shared memory.
//   phase C  one CTA per slab of 8 features, all three parties:  V2_p[f] = XX_p[:,f]^T x E_p - r'_p[f]  (the transposed
//            batch is never materialised: column slabs of the gathered rows are 64-byte pieces, L2 resident since phase
//            A); the three parties' V2 of a feature meet in the same CTA, so the second open-and-truncate and
//            w_p -= update_p  (Regression.h:166-171) finish locally.
//   -- grid barrier --
//
// Arithmetic is wrapping 64-bit add / mul only, so the order of the partial sums is immaterial: shares of w and the
// keystream offsets are bit-identical to the facade loop, to the graph replay and to the oracle
// (tests/test_gpu_sh3.py::test_fused_sgd_matches_facade_and_oracle).
#include "aes.cuh"

namespace aby3cu {
namespace {

struct SgdKeys { AesKey kn[3], kp[3]; };

struct SgdParams {
    const u64* X[3][2];
    const u64* Y[3][2];
    u64* w[3][2];
    const u64* idx;          // iters x B row indices
    u64 B, F, iters;
    unsigned d1, d2;         // shift of the first product (D) and of the second (D + log2(B / lr))
    u64 en[3], ep[3];        // first element of the next / prev common keystream at iteration 0
    u64* V1;                 // [3][B]      XX w - r
    u64* E;                  // [3][2][B]   RTrunc - YY
    unsigned long long* bar; // grid barrier counter, zero at launch
};

constexpr int kSgdThreads = 256;
constexpr int kSlab = 8;     // features per phase-C task: 64-byte pieces of a row

__device__ __forceinline__ u64 ks_elem(u32 lane4, const AesKey& key, u64 e) {
    u32 o[4];
    aes_encrypt_ctr(lane4, key, e >> 1, o);
    return (e & 1) ? (((u64)o[3] << 32) | o[2]) : (((u64)o[1] << 32) | o[0]);
}

__device__ __forceinline__ ulonglong2 ld_nc2(const u64* p) { return __ldg(reinterpret_cast<const ulonglong2*>(p)); }
__device__ __forceinline__ ulonglong2 ld_cg2(const u64* p) { return __ldcg(reinterpret_cast<const ulonglong2*>(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// All CTAs of the (co-resident, cooperatively launched) grid.  Monotonic counter: barrier k completes at k * gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long& passed) {
    __syncthreads();
    if (threadIdx.x == 0) {
        passed += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1ull);
        unsigned long long seen;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
            if (++spins == 0x40000000u) __trap();        // never hang the device: a lost CTA is a bug, not a wait
        } while (seen < passed);
    }
    __syncthreads();
}

__device__ __forceinline__ u64 warp_sum(u64 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kSgdThreads, 1) k_sgd_linear(const __grid_constant__ SgdParams P, const __grid_constant__ SgdKeys K) {
    // shared memory: the AES table, then per-iteration staging
    u64* sm = reinterpret_cast<u64*>(aby3_smem + kAesTableWords);
    u64* sEs = sm;                       // [3][B]  e0 + e1 of the final error shares
    u64* sE0 = sEs + 3 * P.B;            // [3][B]  e0
    u64* sRow = sE0 + 3 * P.B;           // [B]     row offsets (elements) of the batch
    u64* sRed = sRow + P.B;              // [8 warps][24]
    u64* sKs = sRed + 8 * 24;            // [2][24] keystream words of the slab's second truncation pair
    aes_table_init();
    __syncthreads();

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Tl = lane * 4;
    const u64 B = P.B, F = P.F, S = B + F;
    const u64 slabs = (F + kSlab - 1) / kSlab;
    unsigned long long passed = 0;

    for (u64 it = 0; it < P.iters; ++it) {
        const u64* idx = P.idx + it * B;
        // ------------------------------------------------------------------ phase A
        for (u64 b = blockIdx.x; b < B; b += gridDim.x) {
            const u64 row = __ldg(idx + b), off = row * F;
            u64 acc[3] = {0, 0, 0};
            // first chunk: loads in flight while six threads draw the truncation pair
            ulonglong2 x0[3], x1[3], w0[3], w1[3];
            u64 f = 2 * (u64)tid;
            const bool have = f < F;
            if (have) {
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    x0[p] = ld_nc2(P.X[p][0] + off + f); x1[p] = ld_nc2(P.X[p][1] + off + f);
                    w0[p] = ld_cg2(P.w[p][0] + f);       w1[p] = ld_cg2(P.w[p][1] + f);
                }
            }
            u64 ks = 0, yy = 0;
            if (warp == 7 && lane < 6) {                       // lane = 2 * party + stream (0: next -> t0, 1: prev -> t1)
                const int p = lane >> 1, s = lane & 1;
                yy = __ldg(P.Y[p][s] + row);
                ks = s ? ks_elem(Tl, K.kp[p], P.ep[p] + it * S + b) : ks_elem(Tl, K.kn[p], P.en[p] + it * S + b);
            }
            if (have) {
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    acc[p] += x0[p].x * (w0[p].x + w1[p].x) + x1[p].x * w0[p].x + x0[p].y * (w0[p].y + w1[p].y) + x1[p].y * w0[p].y;
            }
            for (f += 2 * kSgdThreads; f < F; f += 2 * kSgdThreads) {
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const ulonglong2 a0 = ld_nc2(P.X[p][0] + off + f), a1 = ld_nc2(P.X[p][1] + off + f);
                    const ulonglong2 b0 = ld_cg2(P.w[p][0] + f), b1 = ld_cg2(P.w[p][1] + f);
                    acc[p] += a0.x * (b0.x + b1.x) + a1.x * b0.x + a0.y * (b0.y + b1.y) + a1.y * b0.y;
                }
            }
#pragma unroll
            for (int p = 0; p < 3; ++p) acc[p] = warp_sum(acc[p]);
            if (lane == 0) { sRed[warp * 24 + 0] = acc[0]; sRed[warp * 24 + 1] = acc[1]; sRed[warp * 24 + 2] = acc[2]; }
            __syncthreads();
            if (warp == 7 && lane < 6) {
                const int p = lane >> 1, s = lane & 1;
                P.E[(2 * p + s) * B + b] = sar(ks, P.d1 + 2) - yy;                    // RTrunc - YY
                if (s == 0) {
                    u64 v = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) v += sRed[w * 24 + p];
                    P.V1[p * B + b] = v - sar(ks, 2);                                 // XX w - r
                }
            }
            __syncthreads();
        }
        grid_barrier(P.bar, passed);

        // ------------------------------------------------------------------ phase B
        for (u64 b = tid; b < B; b += kSgdThreads) {
            const u64 s = __ldcg(P.V1 + b) + __ldcg(P.V1 + B + b) + __ldcg(P.V1 + 2 * B + b);
            const u64 o = sar(s, P.d1);
            u64 e[3][2];
#pragma unroll
            for (int p = 0; p < 3; ++p) { e[p][0] = __ldcg(P.E + (2 * p) * B + b); e[p][1] = __ldcg(P.E + (2 * p + 1) * B + b); }
            e[0][0] += o;                                       // party 0 adds into its share 0, party 1 into its share 1
            e[1][1] += o;
#pragma unroll
            for (int p = 0; p < 3; ++p) { sEs[p * B + b] = e[p][0] + e[p][1]; sE0[p * B + b] = e[p][0]; }
            sRow[b] = __ldg(idx + b) * F;
        }
        __syncthreads();
        // the rows of the next batch on their way into L2 while this one finishes
        if (it + 1 < P.iters) {
            for (u64 b = blockIdx.x; b < B; b += gridDim.x) {
                const u64 off = __ldg(idx + B + b) * F;
                for (u64 l = tid; l < 6 * ((F + 15) / 16); l += kSgdThreads) {
                    const u64 pl = l / ((F + 15) / 16), c = l % ((F + 15) / 16);
                    prefetch_l2(P.X[pl >> 1][pl & 1] + off + 16 * c);
                }
            }
        }

        // ------------------------------------------------------------------ phase C
        for (u64 slab = blockIdx.x; slab < slabs; slab += gridDim.x) {
            const u64 f0 = slab * kSlab;
            const u32 q = tid & 3;
            const u64 fq = f0 + 2 * q;
            u64 acc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
            // second truncation pair of the slab's features: 48 keystream words, drawn while the loads are in flight
            ulonglong2 a0[3], a1[3];
            u64 b = tid >> 2;
            const bool have = fq < F && b < B;
            if (have) {
                const u64 off = sRow[b] + fq;
#pragma unroll
                for (int p = 0; p < 3; ++p) { a0[p] = ld_nc2(P.X[p][0] + off); a1[p] = ld_nc2(P.X[p][1] + off); }
            }
            if (tid >= 64 && tid < 64 + 48) {                   // warps 2 and 3: t = stream * 24 + party * 8 + feature
                const u32 t = tid - 64, s = t / 24, p = (t % 24) / 8, j = t % 8;
                const u64 e = it * S + B + f0 + j;
                sKs[t] = (f0 + j < F) ? (s ? ks_elem(Tl, K.kp[p], P.ep[p] + e) : ks_elem(Tl, K.kn[p], P.en[p] + e)) : 0;
            }
            if (have) {
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const u64 es = sEs[p * B + b], e0 = sE0[p * B + b];
                    acc[p][0] += a0[p].x * es + a1[p].x * e0;
                    acc[p][1] += a0[p].y * es + a1[p].y * e0;
                }
            }
            if (fq < F) {
                for (b += kSgdThreads / 4; b < B; b += kSgdThreads / 4) {
                    const u64 off = sRow[b] + fq;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const ulonglong2 c0 = ld_nc2(P.X[p][0] + off), c1 = ld_nc2(P.X[p][1] + off);
                        const u64 es = sEs[p * B + b], e0 = sE0[p * B + b];
                        acc[p][0] += c0.x * es + c1.x * e0;
                        acc[p][1] += c0.y * es + c1.y * e0;
                    }
                }
            }
            // lanes with the same q hold partial sums of the same two features
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    u64 v = acc[p][h];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (lane < 4) sRed[warp * 24 + p * 8 + 2 * lane + h] = v;
                }
            __syncthreads();
            if (tid < 24) {                                     // t = party * 8 + feature, all in warp 0
                const u32 p = tid / 8, j = tid % 8;
                const u64 f = f0 + j;
                u64 v = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += sRed[w * 24 + tid];
                const u64 t0 = sKs[tid], t1 = sKs[24 + tid];
                v -= sar(t0, 2);                                // XX^T e - r'
                const u64 s = __shfl_sync(0x00ffffffu, v, j) + __shfl_sync(0x00ffffffu, v, 8 + j) + __shfl_sync(0x00ffffffu, v, 16 + j);
                const u64 o = sar(s, P.d2);
                u64 u0 = sar(t0, P.d2 + 2), u1 = sar(t1, P.d2 + 2);
                if (p == 0) u0 += o;
                if (p == 1) u1 += o;
                if (f < F) {
                    P.w[p][0][f] = __ldcg(P.w[p][0] + f) - u0;                          // w -= update
                    P.w[p][1][f] = __ldcg(P.w[p][1] + f) - u1;
                }
            }
            __syncthreads();
        }
        grid_barrier(P.bar, passed);
    }
}

}  // namespace
}  // namespace aby3cu

using namespace aby3cu;

extern "C" {

size_t aby3cu_sgd_linear_colocated_work_bytes(uint64_t B) { return (size_t)(9 * B + 8) * 8; }

int aby3cu_sgd_linear_colocated(aby3cu_ctx* ctx, const int64_t* const* d_X, const int64_t* const* d_Y, int64_t* const* d_w,
                                const uint64_t* d_batch_idx, uint64_t F, uint64_t B, uint64_t iters, uint64_t shift1,
                                uint64_t shift2, const uint8_t* const* key_next, const uint64_t* elem_next,
                                const uint8_t* const* key_prev, const uint64_t* elem_prev, void* d_work) {
    ABY3CU_REQUIRE(ctx && d_X && d_Y && d_w && d_batch_idx && key_next && elem_next && key_prev && elem_prev && d_work,
                   "sgd_linear_colocated: null argument");
    if (!iters) return 0;
    ABY3CU_REQUIRE(B >= 1 && F >= 2 && F % 2 == 0, "sgd_linear_colocated: F must be even (16-byte row pieces)");
    ABY3CU_REQUIRE(shift1 + 2 < 64 && shift2 + 2 < 64, "sgd_linear_colocated: shift too large");
    const size_t smem = (size_t)kAesTableBytes + (7 * B + 8 * 24 + 48) * 8;
    ABY3CU_REQUIRE(smem <= 200 * 1024, "sgd_linear_colocated: batch too large for the shared-memory staging");
    DeviceGuard g(ctx->device);
    SgdParams P;
    SgdKeys K;
    for (int p = 0; p < 3; ++p) {
        for (int s = 0; s < 2; ++s) {
            P.X[p][s] = (const u64*)d_X[2 * p + s];
            P.Y[p][s] = (const u64*)d_Y[2 * p + s];
            P.w[p][s] = (u64*)d_w[2 * p + s];
            ABY3CU_REQUIRE(P.X[p][s] && P.Y[p][s] && P.w[p][s], "sgd_linear_colocated: null matrix");
            ABY3CU_REQUIRE(al16(P.X[p][s]) && al16(P.w[p][s]), "sgd_linear_colocated: X and w must be 16-byte aligned");
        }
        host_expand_key(key_next[p], &K.kn[p]);
        host_expand_key(key_prev[p], &K.kp[p]);
        P.en[p] = elem_next[p];
        P.ep[p] = elem_prev[p];
    }
    P.idx = d_batch_idx; P.B = B; P.F = F; P.iters = iters; P.d1 = (unsigned)shift1; P.d2 = (unsigned)shift2;
    P.V1 = (u64*)d_work;
    P.E = P.V1 + 3 * B;
    P.bar = (unsigned long long*)(P.E + 6 * B);
    ABY3CU_CHECK(cudaMemsetAsync(P.bar, 0, 64, ctx->stream));
    ABY3CU_CHECK(cudaFuncSetAttribute(k_sgd_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const u64 slabs = (F + kSlab - 1) / kSlab;
    u64 want = B > slabs ? B : slabs;
    unsigned grid = (unsigned)(want < (u64)ctx->sm_count ? want : (u64)ctx->sm_count);
    int per_sm = 0;
    ABY3CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sgd_linear, kSgdThreads, smem));
    ABY3CU_REQUIRE(per_sm >= 1, "sgd_linear_colocated: the kernel does not fit an SM");
    void* args[] = {(void*)&P, (void*)&K};
    // cooperative launch: the driver guarantees (or refuses) co-residency of the whole grid -- the grid barrier needs it
    ABY3CU_CHECK(cudaLaunchCooperativeKernel((const void*)k_sgd_linear, dim3(grid), dim3(kSgdThreads), args, smem, ctx->stream));
    return post_launch(ctx, "k_sgd_linear");
}

}  // extern "C"
