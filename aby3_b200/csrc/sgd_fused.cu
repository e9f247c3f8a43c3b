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
    long long* dbg;          // optional: SM clock of CTA 0 at six points of the first 64 iterations (tools/sgd_probe.py)
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
// bytes: a multiple of 16, p 16-byte aligned
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Barrier over all CTAs of the (co-resident, cooperatively launched) grid, split in two so that work which does not
// depend on the other CTAs -- the next loads, the next keystream words -- is issued between arrive and wait.
// Monotonic counter: barrier k completes at k * gridDim.x.  Call both halves from every thread.
__device__ __forceinline__ void grid_arrive(unsigned long long* ctr, long long* stamp = nullptr) {
    __syncthreads();                                         // this CTA's writes are done ...
    if (stamp) stamp[0] = clock64();
    if (threadIdx.x == 0)                                    // ... and published by one release (cumulative over the CTA barrier)
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ull) : "memory");
    if (stamp) stamp[1] = clock64();
}
__device__ __forceinline__ void grid_wait(unsigned long long* ctr, unsigned long long& passed) {
    passed += gridDim.x;
    if (threadIdx.x == 0) {
        unsigned long long seen;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
            if (++spins == 0x40000000u) __trap();            // never hang the device: a lost CTA is a bug, not a wait
        } while (seen < passed);
    }
    __syncthreads();
}

__device__ __forceinline__ u64 warp_sum(u64 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- phase A pieces: one batch row, three parties --------------------------------------------------------------
__device__ __forceinline__ void a_load(const SgdParams& P, u64 off, u64 f, ulonglong2 (&x0)[3], ulonglong2 (&x1)[3]) {
    if (f < P.F) {
#pragma unroll
        for (int p = 0; p < 3; ++p) { x0[p] = ld_nc2(P.X[p][0] + off + f); x1[p] = ld_nc2(P.X[p][1] + off + f); }
    }
}
__device__ __forceinline__ void a_mac(const SgdParams& P, u64 f, const ulonglong2 (&x0)[3], const ulonglong2 (&x1)[3], u64 (&acc)[3]) {
    if (f < P.F) {
        ulonglong2 w0[3], w1[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) { w0[p] = ld_cg2(P.w[p][0] + f); w1[p] = ld_cg2(P.w[p][1] + f); }
#pragma unroll
        for (int p = 0; p < 3; ++p)
            acc[p] += x0[p].x * (w0[p].x + w1[p].x) + x1[p].x * w0[p].x + x0[p].y * (w0[p].y + w1[p].y) + x1[p].y * w0[p].y;
    }
}
// first truncation pair of row b (warp 7, lane = 2 * party + stream; 0: next -> t0, 1: prev -> t1) and the label
__device__ __forceinline__ void a_pair(const SgdParams& P, const SgdKeys& K, u64 it, u64 b, u64 row, u32 warp, u32 lane, u64& ks, u64& yy) {
    if (warp == 7 && lane < 6) {
        const int p = lane >> 1, s = lane & 1;
        const u64 e = it * (P.B + P.F) + b;
        yy = __ldg(P.Y[p][s] + row);
        ks = s ? ks_elem(lane * 4, K.kp[p], P.ep[p] + e) : ks_elem(lane * 4, K.kn[p], P.en[p] + e);
    }
}
__device__ __forceinline__ void a_finish(const SgdParams& P, u64 b, u32 warp, u32 lane, u64 ks, u64 yy, u64 (&acc)[3], u64* sRed) {
#pragma unroll
    for (int p = 0; p < 3; ++p) acc[p] = warp_sum(acc[p]);
    if (lane == 0) { sRed[warp * 24 + 0] = acc[0]; sRed[warp * 24 + 1] = acc[1]; sRed[warp * 24 + 2] = acc[2]; }
    __syncthreads();
    if (warp == 7 && lane < 6) {
        const int p = lane >> 1, s = lane & 1;
        P.E[(2 * p + s) * P.B + b] = sar(ks, P.d1 + 2) - yy;                          // RTrunc - YY
        if (s == 0) {
            u64 v = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sRed[w * 24 + p];
            P.V1[p * P.B + b] = v - sar(ks, 2);                                       // XX w - r
        }
    }
    __syncthreads();
}

// ---- phase C pieces: one slab of 8 features, three parties -----------------------------------------------------
__device__ __forceinline__ void c_load(const SgdParams& P, u64 off, bool ok, ulonglong2 (&c0)[3], ulonglong2 (&c1)[3]) {
    if (ok) {
#pragma unroll
        for (int p = 0; p < 3; ++p) { c0[p] = ld_nc2(P.X[p][0] + off); c1[p] = ld_nc2(P.X[p][1] + off); }
    }
}
__device__ __forceinline__ void c_mac(const u64* sEs, const u64* sE0, u64 B, u64 b, bool ok, const ulonglong2 (&c0)[3],
                                      const ulonglong2 (&c1)[3], u64 (&acc)[3][2]) {
    if (ok) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const u64 es = sEs[p * B + b], e0 = sE0[p * B + b];
            acc[p][0] += c0[p].x * es + c1[p].x * e0;
            acc[p][1] += c0[p].y * es + c1[p].y * e0;
        }
    }
}
// second truncation pair of the slab: 48 keystream words (threads 64..111: t = stream * 24 + party * 8 + feature)
template <class PT>
__device__ __forceinline__ void c_pairs(const PT& P, const SgdKeys& K, u64 it, u64 f0, u32 tid, u64* sKs) {
    if (tid >= 64 && tid < 64 + 48) {
        const u32 t = tid - 64, s = t / 24, p = (t % 24) / 8, j = t % 8;
        const u64 e = it * (P.B + P.F) + P.B + f0 + j;
        const u32 Tl = (tid & 31) * 4;
        sKs[t] = (f0 + j < P.F) ? (s ? ks_elem(Tl, K.kp[p], P.ep[p] + e) : ks_elem(Tl, K.kn[p], P.en[p] + e)) : 0;
    }
}
// reduce the partial sums, open-and-truncate, w -= update.  wo0 / wo1: the old w words (threads < 24).
__device__ __forceinline__ void c_finish(const SgdParams& P, u64 f0, u32 tid, u32 warp, u32 lane, u64 (&acc)[3][2], u64* sRed,
                                         const u64* sKs, u64 wo0, u64 wo1) {
    // lanes with the same (lane & 3) hold partial sums of the same two features
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
    if (tid < 24) {                                             // t = party * 8 + feature, all in warp 0
        const u32 p = tid / 8, j = tid % 8;
        const u64 f = f0 + j;
        u64 v = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += sRed[w * 24 + tid];
        const u64 t0 = sKs[tid], t1 = sKs[24 + tid];
        v -= sar(t0, 2);                                        // XX^T e - r'
        const u64 s = __shfl_sync(0x00ffffffu, v, j) + __shfl_sync(0x00ffffffu, v, 8 + j) + __shfl_sync(0x00ffffffu, v, 16 + j);
        const u64 o = sar(s, P.d2);
        u64 u0 = sar(t0, P.d2 + 2), u1 = sar(t1, P.d2 + 2);
        if (p == 0) u0 += o;                                    // party 0 adds into its share 0, party 1 into its share 1
        if (p == 1) u1 += o;
        if (f < P.F) {
            P.w[p][0][f] = wo0 - u0;                            // w -= update
            P.w[p][1][f] = wo1 - u1;
        }
    }
    __syncthreads();
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

    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 B = P.B, F = P.F;
    const u64 slabs = (F + kSlab - 1) / kSlab;
    const u64 bMine = blockIdx.x, slabMine = blockIdx.x;         // the tasks whose loads are issued ahead of the barriers
    const u64 fA0 = 2 * (u64)tid, fA1 = fA0 + 2 * kSgdThreads;   // phase A: two 16-byte pieces per thread held in registers
    const u64 fq = slabMine * kSlab + 2 * (tid & 3);             // phase C: this thread's two features of the slab
    const u64 bC0 = tid >> 2, bC1 = bC0 + kSgdThreads / 4;       //          and its two batch rows held in registers
    unsigned long long passed = 0;

    // registers carried across the barriers
    ulonglong2 ax0[2][3], ax1[2][3];     // phase A: this CTA's row of X (first 1024 features)
    u64 ks = 0, yy = 0;                  //          first truncation pair word and label (warp 7, lanes < 6)
    if (bMine < B && P.iters) {
        const u64 row = __ldg(P.idx + bMine), off = row * F;
        a_load(P, off, fA0, ax0[0], ax1[0]);
        a_load(P, off, fA1, ax0[1], ax1[1]);
        a_pair(P, K, 0, bMine, row, warp, lane, ks, yy);
    }

#define SGD_STAMP(k) do { if (P.dbg && blockIdx.x == 0 && tid == 0 && it < 64) P.dbg[it * 16 + (k)] = clock64(); } while (0)
    for (u64 it = 0; it < P.iters; ++it) {
        const u64* idx = P.idx + it * B;
        SGD_STAMP(0);
        // ------------------------------------------------------------------ phase A: V1 = XX w - r, E = RTrunc - YY
        if (bMine < B) {
            const u64 off = __ldg(idx + bMine) * F;
            u64 acc[3] = {0, 0, 0};
            a_mac(P, fA0, ax0[0], ax1[0], acc);
            a_mac(P, fA1, ax0[1], ax1[1], acc);
            for (u64 f = fA1 + 2 * kSgdThreads; f < F; f += 2 * kSgdThreads) {
                ulonglong2 x0[3], x1[3];
                a_load(P, off, f, x0, x1);
                a_mac(P, f, x0, x1, acc);
            }
            a_finish(P, bMine, warp, lane, ks, yy, acc, sRed);
        }
        for (u64 b = bMine + gridDim.x; b < B; b += gridDim.x) {
            const u64 row = __ldg(idx + b), off = row * F;
            u64 acc[3] = {0, 0, 0}, ks2 = 0, yy2 = 0;
            a_pair(P, K, it, b, row, warp, lane, ks2, yy2);
            for (u64 f = fA0; f < F; f += 2 * kSgdThreads) {
                ulonglong2 x0[3], x1[3];
                a_load(P, off, f, x0, x1);
                a_mac(P, f, x0, x1, acc);
            }
            a_finish(P, b, warp, lane, ks2, yy2, acc, sRed);
        }
        SGD_STAMP(1);
        grid_arrive(P.bar, (P.dbg && blockIdx.x == 0 && tid == 0 && it < 64) ? P.dbg + it * 16 + 8 : nullptr);
        // between arrive and wait: everything of phase C that does not depend on the other CTAs
        ulonglong2 cx0[2][3], cx1[2][3];
        u64 wo0 = 0, wo1 = 0;
        const bool mineC = slabMine < slabs && fq < F;
        if (slabMine < slabs) {
            const bool ok0 = mineC && bC0 < B, ok1 = mineC && bC1 < B;
            const u64 r0 = ok0 ? __ldg(idx + bC0) * F : 0, r1 = ok1 ? __ldg(idx + bC1) * F : 0;
            SGD_STAMP(12);
            c_load(P, r0 + fq, ok0, cx0[0], cx1[0]);
            c_load(P, r1 + fq, ok1, cx0[1], cx1[1]);
            c_pairs(P, K, it, slabMine * kSlab, tid, sKs);
            if (tid < 24 && slabMine * kSlab + tid % 8 < F) {
                wo0 = __ldcg(P.w[tid / 8][0] + slabMine * kSlab + tid % 8);
                wo1 = __ldcg(P.w[tid / 8][1] + slabMine * kSlab + tid % 8);
            }
        }
        SGD_STAMP(2);
        grid_wait(P.bar, passed);
        SGD_STAMP(3);

        // ------------------------------------------------------------------ phase B: open the first product
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
        SGD_STAMP(4);

        // ------------------------------------------------------------------ phase C: V2 = XX^T e - r', open, w -= update
        if (slabMine < slabs) {
            u64 acc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
            c_mac(sEs, sE0, B, bC0, mineC && bC0 < B, cx0[0], cx1[0], acc);
            c_mac(sEs, sE0, B, bC1, mineC && bC1 < B, cx0[1], cx1[1], acc);
            if (mineC) {
                for (u64 b = bC1 + kSgdThreads / 4; b < B; b += kSgdThreads / 4) {
                    ulonglong2 c0[3], c1[3];
                    c_load(P, sRow[b] + fq, true, c0, c1);
                    c_mac(sEs, sE0, B, b, true, c0, c1, acc);
                }
            }
            c_finish(P, slabMine * kSlab, tid, warp, lane, acc, sRed, sKs, wo0, wo1);
        }
        for (u64 slab = slabMine + gridDim.x; slab < slabs; slab += gridDim.x) {
            const u64 f0 = slab * kSlab, fq2 = f0 + 2 * (tid & 3);
            u64 acc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
            c_pairs(P, K, it, f0, tid, sKs);
            u64 v0 = 0, v1 = 0;
            if (tid < 24 && f0 + tid % 8 < F) { v0 = __ldcg(P.w[tid / 8][0] + f0 + tid % 8); v1 = __ldcg(P.w[tid / 8][1] + f0 + tid % 8); }
            if (fq2 < F) {
                for (u64 b = bC0; b < B; b += kSgdThreads / 4) {
                    ulonglong2 c0[3], c1[3];
                    c_load(P, sRow[b] + fq2, true, c0, c1);
                    c_mac(sEs, sE0, B, b, true, c0, c1, acc);
                }
            }
            c_finish(P, f0, tid, warp, lane, acc, sRed, sKs, v0, v1);
        }
        SGD_STAMP(5);
        grid_arrive(P.bar, (P.dbg && blockIdx.x == 0 && tid == 0 && it < 64) ? P.dbg + it * 16 + 10 : nullptr);
        // between arrive and wait: the next iteration's row of X, truncation pair word and label
        if (it + 1 < P.iters && bMine < B) {
            const u64 row = __ldg(idx + B + bMine), off = row * F;
            a_load(P, off, fA0, ax0[0], ax1[0]);
            a_load(P, off, fA1, ax0[1], ax1[1]);
            a_pair(P, K, it + 1, bMine, row, warp, lane, ks, yy);
        }
        SGD_STAMP(6);
        grid_wait(P.bar, passed);
        SGD_STAMP(7);
    }
#undef SGD_STAMP
}


// =====================================================================================================================
// The feature-resident variant: ONE grid barrier per iteration, w never leaves its CTA.
//
// CTA c owns the slab of 8 features [8c, 8c + 8): it keeps those entries of w (three parties, both planes) in shared
// memory for the whole run and holds the slab's column pieces of the batch (B x 64 bytes per plane, half a piece per thread) in
// registers.  Both products of an iteration contract over that same data:
//   * first product:   the CTA's partial  sum_{f in slab} XX_p[b,f] w_p[f], summed over the three parties (only the OPENED
//     value V1_0 + V1_1 + V1_2 is ever used, Sh3Evaluator.cpp:712-718), goes into  S[b]  with one RED per row; the CTA
//     that "owns" row b (b mod grid) adds  -(r_0 + r_1 + r_2)[b]  and publishes  E_p[b] = RTrunc_p[b] - YY_p[b];
//   * -- the grid barrier --  every thread then reads S[b] and E[.][b] of its own two rows;
//   * second product:  XX_p[:,f]^T e_p  for the slab's features over all rows: complete inside the CTA, so are the second
//     open-and-truncate and  w -= update.
// The batch pieces of iteration it + 1 are loaded (from L2: the owners prefetch whole rows of iteration it + 2 one
// iteration ahead) while iteration it waits at its barrier; keystream words are drawn between arrive and wait.
// S is triple-buffered and E double-buffered by iteration so that a fast CTA never overwrites what a slow one still reads.
// Shapes: B <= 128 (one row per pair of threads), F % 4 == 0, F / 8 <= SMs; everything else goes to k_sgd_linear above.
// =====================================================================================================================
struct SlabParams {
    const u64* X[3][2];
    const u64* Y[3][2];
    u64* w[3][2];
    const u64* idx;
    u64 B, F, iters;
    unsigned d1, d2;
    u64 en[3], ep[3];
    unsigned long long* S;   // [3][B]     opened first product, accumulated by RED; zero at launch
    u64* E;                  // [2][6][B]  RTrunc - YY, [2 * party + plane]
    unsigned long long* bar;
    long long* dbg;
    unsigned flags;          // 1: prefetch the index lines, 2: bulk-prefetch the rows two iterations ahead
};

__device__ __forceinline__ void red_add(unsigned long long* p, u64 v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct alignas(32) U64x4 { u64 v[4]; };

__global__ void __launch_bounds__(kSgdThreads, 1) k_sgd_linear_slab(const __grid_constant__ SlabParams P, const __grid_constant__ SgdKeys K) {
    u64* sm = reinterpret_cast<u64*>(aby3_smem + kAesTableWords);
    u64* sW = sm;                        // [6][8]   this CTA's entries of w, [2 * party + plane][feature]
    u64* sRed = sW + 48;                 // [8 warps][24]
    u64* sKs = sRed + 8 * 24;            // [2][24]  second truncation pair of the slab
    u64* sE = sKs + 48;                  // [6][128] e0 + e1 and e0 of the opened error, per party
    u64* sX = sE + 6 * 128;              // [2][6][128][8] the slab's pieces of the batch: this iteration's and the next one's (cp.async)
    aes_table_init();

    // thread = (batch row, half of the slab): four features of one row, three parties, both planes, in registers
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = tid & 1;
    const u64 B = P.B, F = P.F, S = B + F;
    const u64 f0 = (u64)blockIdx.x * kSlab, fq = f0 + 4 * q;
    const u64 bRow = tid >> 1;
    const bool rOk = bRow < B, live = rOk && fq < F;        // F % 4 == 0: a piece is whole or absent
    // owner duty: threads 128..151 = 4 slots x (party, stream); slot m handles the rows blockIdx.x + (m + 4 j) * gridDim.x
    const bool ownerT = tid >= 128 && tid < 152;
    const u32 oSlot = ownerT ? (tid - 128) / 6 : 0, oP = ownerT ? ((tid - 128) % 6) >> 1 : 0, oS = ownerT ? (tid - 128) & 1 : 0;
    const u64 oRow0 = (u64)blockIdx.x + (u64)oSlot * gridDim.x;
    unsigned long long passed = 0;

    if (tid < 48 && f0 + tid % 8 < F) sW[tid] = P.w[tid / 16][(tid / 8) & 1][f0 + tid % 8];
    else if (tid < 48) sW[tid] = 0;

    auto rowIdx = [&](u64 it) -> u64 { return __ldg(P.idx + it * B + bRow); };       // multiplied by F where it is used
    u64 r1 = 0, r2 = 0;                                  // this thread's row one and two iterations ahead
    // The batch pieces travel global -> shared by cp.async (LDGSTS): no registers are held while they are in flight and the
    // issuing thread does not queue behind them, so they are issued a whole barrier ahead.  A thread only ever reads the
    // 6 x 32 bytes it copied itself: cp.async.wait_group is all the synchronisation they need.
    auto pieces = [&](u64 rowIndex, u32 buf) {
        const u64 off = rowIndex * F + fq;
#pragma unroll
        for (int pl = 0; pl < 6; ++pl) {
            const u64* src = P.X[pl >> 1][pl & 1] + off;
            const u32 dst = (u32)__cvta_generic_to_shared(sX + ((buf * 6 + pl) * 128 + bRow) * 8 + 4 * q);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 2) : "memory");
        }
    };
    U64x4 xa[3], xb[3];                                  // planes 0 / 1 of the current iteration
    if (rOk && P.iters) {
        if (live) pieces(rowIdx(0), 0);
        if (P.iters > 1) r1 = rowIdx(1);
        if (P.iters > 2) r2 = rowIdx(2);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // first truncation pair word and label of the owned row (first pass), one iteration ahead
    u64 ks = 0, yy = 0;
    auto pair1 = [&](u64 it, u64 b, u64& ks_, u64& yy_) {
        const u64 e = it * S + b;
        yy_ = __ldg(P.Y[oP][oS] + __ldg(P.idx + it * B + b));
        ks_ = oS ? ks_elem(lane * 4, K.kp[oP], P.ep[oP] + e) : ks_elem(lane * 4, K.kn[oP], P.en[oP] + e);
    };
    __syncthreads();
    if (ownerT && oRow0 < B && P.iters) pair1(0, oRow0, ks, yy);

#define SLAB_STAMP(k) do { if (P.dbg && blockIdx.x == 0 && tid == 0 && it < 64) P.dbg[it * 16 + (k)] = clock64(); } while (0)
    for (u64 it = 0; it < P.iters; ++it) {
        unsigned long long* Sit = P.S + (it % 3) * B;
        u64* Eit = P.E + (it & 1) * 6 * B;
        SLAB_STAMP(0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (live) {
                xa[p] = *reinterpret_cast<const U64x4*>(sX + (((it & 1) * 6 + 2 * p) * 128 + bRow) * 8 + 4 * q);
                xb[p] = *reinterpret_cast<const U64x4*>(sX + (((it & 1) * 6 + 2 * p + 1) * 128 + bRow) * 8 + 4 * q);
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) xa[p].v[h] = xb[p].v[h] = 0;
            }
        }
        // ---- first product: partial over the slab, summed over the parties, one RED per row
        {
            u64 v = 0;
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const u64 w0 = sW[(2 * p) * 8 + 4 * q + h], ws = w0 + sW[(2 * p + 1) * 8 + 4 * q + h];
                    v += xa[p].v[h] * ws + xb[p].v[h] * w0;
                }
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (q == 0 && rOk) red_add(Sit + bRow, v);
        }
        // ---- owner: -(r_0 + r_1 + r_2) into S, E = RTrunc - YY
        if (ownerT) {
            u64 k1 = ks, y1 = yy;
            for (u64 b = oRow0; b < B; b += 4ull * gridDim.x) {
                if (b != oRow0) pair1(it, b, k1, y1);
                Eit[(2 * oP + oS) * B + b] = sar(k1, P.d1 + 2) - y1;
                if (oS == 0) red_add(Sit + b, 0 - sar(k1, 2));
            }
        }
        SLAB_STAMP(2);
        grid_arrive(P.bar);
        SLAB_STAMP(3);
        // ---- between arrive and wait: keystream words (second pair of this slab, first pair of the owned row, next iteration)
        c_pairs(P, K, it, f0, tid, sKs);
        if (ownerT && oRow0 < B && it + 1 < P.iters) pair1(it + 1, oRow0, ks, yy);
        if (it + 1 < P.iters && live) pieces(r1, (u32)((it + 1) & 1));          // the next iteration's batch pieces
        asm volatile("cp.async.commit_group;" ::: "memory");
        SLAB_STAMP(4);
        grid_wait(P.bar, passed);
        SLAB_STAMP(5);

        // ---- open the first product: thread b < B reads row b's S and E once (coalesced, L2), the CTA shares them.
        u64 sv = 0, ev[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        if (tid < B) {
            sv = __ldcg(Sit + tid);
#pragma unroll
            for (int p = 0; p < 3; ++p) { ev[p][0] = __ldcg(Eit + (2 * p) * B + tid); ev[p][1] = __ldcg(Eit + (2 * p + 1) * B + tid); }
        }
        SLAB_STAMP(8);
        // ---- the row index three ahead, whole rows two ahead into L2
        if ((P.flags & 1) && it + 8 < P.iters && tid * 16 < B) prefetch_l2(P.idx + (it + 8) * B + tid * 16);     // the index lines themselves
        u64 r3 = 0;
        if (it + 3 < P.iters && rOk) r3 = rowIdx(it + 3);
        // whole rows of iteration it + 2 on their way into L2: the TMA engine takes a row per instruction (8 KiB at F = 1024).
        // The owner CTA of a row issues them; the two threads that hold its index share the six planes.
        if ((P.flags & 2) && it + 2 < P.iters && rOk && bRow % gridDim.x == blockIdx.x) {
            for (u32 pl = q; pl < 6; pl += 2) bulk_prefetch_l2(P.X[pl >> 1][pl & 1] + r2 * F, (u32)(F * 8));
        }
        SLAB_STAMP(9);
        if (tid < B) {
            const u64 o = sar(sv, P.d1);
            ev[0][0] += o;                                  // party 0 adds into its share 0, party 1 into its share 1
            ev[1][1] += o;
#pragma unroll
            for (int p = 0; p < 3; ++p) { sE[(2 * p) * 128 + tid] = ev[p][0] + ev[p][1]; sE[(2 * p + 1) * 128 + tid] = ev[p][0]; }
        }
        __syncthreads();
        SLAB_STAMP(6);
        // ---- second product, open, w -= update: all inside the CTA
        {
            u64 acc[3][4];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const u64 es = rOk ? sE[(2 * p) * 128 + bRow] : 0, e0 = rOk ? sE[(2 * p + 1) * 128 + bRow] : 0;
#pragma unroll
                for (int h = 0; h < 4; ++h) acc[p][h] = xa[p].v[h] * es + xb[p].v[h] * e0;
            }
            // lanes with the same q hold partial sums of the same four features
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    u64 v = acc[p][h];
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (lane < 2) sRed[warp * 24 + p * 8 + 4 * lane + h] = v;
                }
            __syncthreads();
            if (tid < 24) {                                     // t = party * 8 + feature, all in warp 0
                const u32 p = tid / 8, j = tid % 8;
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
                sW[(2 * p) * 8 + j] -= u0;                      // w -= update
                sW[(2 * p + 1) * 8 + j] -= u1;
            }
        }
        // S of iteration it + 2 starts from zero (its last readers were released by this iteration's barrier)
        if (tid < 128) for (u64 b = (u64)blockIdx.x + (u64)tid * gridDim.x; b < B; b += 128ull * gridDim.x) P.S[((it + 2) % 3) * B + b] = 0;
        r1 = r2; r2 = r3;
        __syncthreads();
        SLAB_STAMP(7);
    }
#undef SLAB_STAMP
    if (tid < 48 && f0 + tid % 8 < F) P.w[tid / 16][(tid / 8) & 1][f0 + tid % 8] = sW[tid];
}

}  // namespace
}  // namespace aby3cu

using namespace aby3cu;

extern "C" {

size_t aby3cu_sgd_linear_colocated_work_bytes(uint64_t B) { return (size_t)(15 * B + 8 + 64 * 16) * 8; }

int aby3cu_sgd_linear_colocated(aby3cu_ctx* ctx, const int64_t* const* d_X, const int64_t* const* d_Y, int64_t* const* d_w,
                                const uint64_t* d_batch_idx, uint64_t F, uint64_t B, uint64_t iters, uint64_t shift1,
                                uint64_t shift2, const uint8_t* const* key_next, const uint64_t* elem_next,
                                const uint8_t* const* key_prev, const uint64_t* elem_prev, void* d_work) {
    ABY3CU_REQUIRE(ctx && d_X && d_Y && d_w && d_batch_idx && key_next && elem_next && key_prev && elem_prev && d_work,
                   "sgd_linear_colocated: null argument");
    if (!iters) return 0;
    ABY3CU_REQUIRE(B >= 1 && F >= 2 && F % 2 == 0, "sgd_linear_colocated: F must be even (16-byte row pieces)");
    ABY3CU_REQUIRE(shift1 + 2 < 64 && shift2 + 2 < 64, "sgd_linear_colocated: shift too large");
    DeviceGuard g(ctx->device);
    const u64 slabs = (F + kSlab - 1) / kSlab;
    // feature-resident kernel (one barrier per iteration) where the shape allows it; ABY3CU_SGD_VARIANT=rows forces the other
    static const bool force_rows = [] { const char* e = getenv("ABY3CU_SGD_VARIANT"); return e && e[0] == 'r'; }();
    bool slab = !force_rows && B <= kSgdThreads / 2 && F % 4 == 0 && slabs <= (u64)ctx->sm_count;
    for (int i = 0; i < 6; ++i) slab = slab && (reinterpret_cast<uintptr_t>(d_X[i]) & 31) == 0;          // 256-bit loads
    const size_t smem = slab ? (size_t)kAesTableBytes + (48 + 8 * 24 + 48 + 6 * 128 + 2 * 6 * 128 * 8) * 8 : (size_t)kAesTableBytes + (7 * B + 8 * 24 + 48) * 8;
    ABY3CU_REQUIRE(smem <= 200 * 1024, "sgd_linear_colocated: batch too large for the shared-memory staging");
    SgdParams P;
    SlabParams Q;
    SgdKeys K;
    for (int p = 0; p < 3; ++p) {
        for (int s = 0; s < 2; ++s) {
            Q.X[p][s] = P.X[p][s] = (const u64*)d_X[2 * p + s];
            Q.Y[p][s] = P.Y[p][s] = (const u64*)d_Y[2 * p + s];
            Q.w[p][s] = P.w[p][s] = (u64*)d_w[2 * p + s];
            ABY3CU_REQUIRE(P.X[p][s] && P.Y[p][s] && P.w[p][s], "sgd_linear_colocated: null matrix");
            ABY3CU_REQUIRE(al16(P.X[p][s]) && al16(P.w[p][s]), "sgd_linear_colocated: X and w must be 16-byte aligned");
        }
        host_expand_key(key_next[p], &K.kn[p]);
        host_expand_key(key_prev[p], &K.kp[p]);
        Q.en[p] = P.en[p] = elem_next[p];
        Q.ep[p] = P.ep[p] = elem_prev[p];
    }
    Q.idx = P.idx = d_batch_idx; Q.B = P.B = B; Q.F = P.F = F; Q.iters = P.iters = iters;
    Q.d1 = P.d1 = (unsigned)shift1; Q.d2 = P.d2 = (unsigned)shift2;
    // work: [15 B words of staging][8 barrier words][64 x 16 stamps]
    P.V1 = (u64*)d_work;
    P.E = P.V1 + 3 * B;
    Q.S = (unsigned long long*)d_work;
    Q.E = (u64*)d_work + 3 * B;
    Q.bar = P.bar = (unsigned long long*)((u64*)d_work + 15 * B);
    // ABY3CU_SGD_STAMPS=1: CTA 0 leaves its SM clock at a few points of the first 64 iterations behind the barrier word
    static const bool stamps = [] { const char* e = getenv("ABY3CU_SGD_STAMPS"); return e && e[0] == '1'; }();
    Q.dbg = P.dbg = stamps ? (long long*)(P.bar + 8) : nullptr;
    static const unsigned flags = [] { const char* e = getenv("ABY3CU_SGD_FLAGS"); return e ? (unsigned)atoi(e) : 3u; }();
    Q.flags = flags;
    ABY3CU_CHECK(cudaMemsetAsync(d_work, 0, aby3cu_sgd_linear_colocated_work_bytes(B), ctx->stream));
    const void* kernel = slab ? (const void*)k_sgd_linear_slab : (const void*)k_sgd_linear;
    ABY3CU_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid;
    if (slab) {
        grid = (unsigned)slabs;
    } else {
        const u64 want = B > slabs ? B : slabs;
        grid = (unsigned)(want < (u64)ctx->sm_count ? want : (u64)ctx->sm_count);
    }
    int per_sm = 0;
    ABY3CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSgdThreads, smem));
    ABY3CU_REQUIRE(per_sm >= 1, "sgd_linear_colocated: the kernel does not fit an SM");
    void* args_rows[] = {(void*)&P, (void*)&K};
    void* args_slab[] = {(void*)&Q, (void*)&K};
    // cooperative launch: the driver guarantees (or refuses) co-residency of the whole grid -- the grid barrier needs it
    ABY3CU_CHECK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kSgdThreads), slab ? args_slab : args_rows, smem, ctx->stream));
    return post_launch(ctx, slab ? "k_sgd_linear_slab" : "k_sgd_linear");
}

}  // extern "C"
