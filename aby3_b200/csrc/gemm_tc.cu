// gemm_tc.cu -- the int64 (mod 2^64) cross-term GEMM on the 5th-generation tensor
// cores:  C (+)= A0*(B0+B1) + A1*B0   (reference: three Eigen i64 products at
// aby3/sh3/Sh3Evaluator.cpp:662-665).
//
// Arithmetic.  Write every 64-bit operand in base 256: a = sum_i a_i 2^(8i),
// b = sum_j b_j 2^(8j).  Modulo 2^64 only the 36 limb pairs with i+j < 8 matter:
//     C = sum_{s<8} 2^(8s) S_s ,   S_s = sum_{i+j=s} sum_k a_i[m,k] b_j[k,n].
// S_s is a u8 x u8 -> s32 contraction: tcgen05.mma kind::i8 with unsigned
// operands.  Accumulators wrap mod 2^32 (saturate bit off).  For s >= 4 only
// S_s mod 2^(64-8s) <= 2^32 is needed, so wrapping is harmless for any K; for
// s <= 3 the exact value is needed, and S_s <= (s+1) * Ke * 255^2 < 2^32 holds
// while Ke <= 16512, i.e. K <= 8256 per launch (Ke = 2K, both products of the
// factored cross term run as ONE contraction [A0|A1] x [B0+B1 ; B0]).  Larger K
// is split on the host and accumulated through C.
//
// Mapping.  One CTA per SM, persistent over 128 x 64 output tiles.  The eight
// accumulators D_0..D_7 of a tile live side by side in TMEM (8 x 64 = 512
// columns, all of it).  Because D_(i+j) sits 64 columns after D_(i+j-1) and the
// B limb planes of a tile are stacked along the MMA's N dimension, ONE
// instruction  D[64i ...] += A_i x [B_0|B_1|...|B_(7-i)]^T  (N = 64(8-i), cut
// into pieces of <= 256) updates every accumulator that limb i of A feeds:
// 12 MMA instructions per 32-deep K step instead of 36.
//
// Data movement.  A pre-pass splits the int64 operands into limb planes stored
// directly in the no-swizzle K-major core-matrix layout the MMA descriptors
// want, one contiguous 32 KiB (A) / 16 KiB (B) chunk per (tile, k-block), so a
// pipeline stage is filled by two bulk TMA copies (cp.async.bulk, UBLKCP) that
// complete on an mbarrier.  Warp 0 = TMA producer, warp 1 = MMA issuer (one
// elected thread), warps 2..5 = epilogue (tcgen05.ld -> recombine limbs mod 2^64
// -> C (+)=).
#include <stdlib.h>

#include "common.cuh"

namespace aby3cu {
namespace tc {

constexpr int TM = 128;            // tile rows  (= MMA M, TMEM lanes)
constexpr int TN = 64;             // tile cols  (per limb accumulator)
constexpr int TK = 32;             // k per stage (= MMA K for 8-bit operands)
#ifndef ABY3CU_GEMM_STAGES
#define ABY3CU_GEMM_STAGES 3
#endif
constexpr int STAGES = ABY3CU_GEMM_STAGES;
constexpr u32 A_CHUNK = 8 * TM * TK;      // 32768 B
constexpr u32 B_CHUNK = 8 * TN * TK;      // 16384 B
constexpr u32 STAGE_BYTES = A_CHUNK + B_CHUNK;
constexpr u32 SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + barriers, tmem slot, alignment slack
constexpr int THREADS = 192;
#ifndef ABY3CU_GEMM_BALANCED
#define ABY3CU_GEMM_BALANCED 1
#endif
constexpr bool kBalancedSplit = ABY3CU_GEMM_BALANCED;
constexpr u64 K_MAX = 8192;        // per launch, see the exactness bound above
constexpr u32 RASTER_M = 8;        // tile rows per rasterisation group

// ------------------------------------------------------------------ PTX wrappers --
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    u32 done = 0;
    for (u32 spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// the same copy with an L2 eviction-priority hint (createpolicy) for the line it leaves behind in L2
__device__ __forceinline__ void bulk_g2s_hint(u32 dst, const void* src, u32 bytes, u32 bar, u64 policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ u64 l2_policy_evict_last() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_first() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u32 bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_i8(u32 d_tmem, u64 a_desc, u64 b_desc, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(u32 taddr, u32 v[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle (canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units):
// LBO = byte distance between the two 16-byte K halves, SBO = between 8-row groups.
__device__ __forceinline__ u64 smem_desc(u32 addr, u32 lbo, u32 sbo) {
    return (u64)((addr & 0x3FFFF) >> 4) | ((u64)(lbo >> 4) << 16) | ((u64)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::i8, D = s32, A and B unsigned 8-bit, both K-major, M = 128
__device__ __forceinline__ constexpr u32 make_idesc(u32 n) {
    return (2u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------ limb packing --
// byte `limb` of four consecutive 64-bit values, packed into one word: three PRMT instead of shifts / masks / ors
__device__ __forceinline__ u32 limb4(const u64* v, int limb) {
    const int b = limb & 3;
    const u32 x0 = limb < 4 ? (u32)v[0] : (u32)(v[0] >> 32), x1 = limb < 4 ? (u32)v[1] : (u32)(v[1] >> 32);
    const u32 x2 = limb < 4 ? (u32)v[2] : (u32)(v[2] >> 32), x3 = limb < 4 ? (u32)v[3] : (u32)(v[3] >> 32);
    const u32 sel = (u32)b | ((u32)(4 + b) << 4);                  // byte b of the first, byte b of the second operand
    const u32 t01 = __byte_perm(x0, x1, sel), t23 = __byte_perm(x2, x3, sel);
    return __byte_perm(t01, t23, 0x5410);
}
// A chunk (mt, kb): [limb 8][kc 2][row-group 16][row 8][16 B]; one thread = one
// row x 2 x 16 k: reads 2 x 128 contiguous bytes, writes 2 x 8 x 16 bytes.
__global__ void __launch_bounds__(128) k_pack_a(const u64* __restrict__ A0, const u64* __restrict__ A1,
                                                u64 row0, u64 rows, u64 M, u64 K, u64 k0, u64 kblocks_half, u8* __restrict__ out, int vec) {
    // 128-thread CTAs (one warp per register file, 48 registers): admitted next to a running GEMM CTA of another party,
    // where this HBM-bound pass costs the tensor pipe nothing
    const u64 kb = blockIdx.x, mt = blockIdx.y;          // kb over both halves
    const int r = threadIdx.x;
    const u64* A = (kb < kblocks_half) ? A0 : A1;
    const u64 m = row0 + mt * TM + r;
#pragma unroll 1
    for (int kc = 0; kc < 2; ++kc) {
        const u64 kbase = k0 + (kb % kblocks_half) * TK + kc * 16;
        u64 v[16];
        if (vec && m < row0 + rows && m < M && kbase + 16 <= K) {
            // the thread's 128 bytes as four 256-bit loads (a warp touches 32 rows: 128 line wavefronts instead of 512)
            const u64* src = A + m * K + kbase;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[j]), "=l"(v[j + 1]), "=l"(v[j + 2]), "=l"(v[j + 3]) : "l"(src + j));
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const u64 k = kbase + j;
                v[j] = (m < row0 + rows && m < M && k < K) ? A[m * K + k] : 0;
            }
        }
        u8* chunk = out + (mt * (2 * kblocks_half) + kb) * A_CHUNK + kc * 2048 + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            u32 w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = limb4(v + 4 * q, i);
            *reinterpret_cast<uint4*>(chunk + i * 4096) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// B chunk (nt, kb): [kc 2][limb 8][row-group 8][row 8][16 B], rows = output columns n.
// First half of K carries B0+B1, second half B0.  One thread = one n x 16 k.
__global__ void __launch_bounds__(128) k_pack_b(const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                u64 K, u64 N, u64 k0, u64 kblocks_half, u8* __restrict__ out) {
    const u64 kb = blockIdx.x, nt = blockIdx.y;
    const int nl = threadIdx.x & 63, kc = threadIdx.x >> 6;
    const bool first = kb < kblocks_half;
    const u64 kbase = k0 + (kb % kblocks_half) * TK + kc * 16;
    const u64 n = nt * TN + nl;
    u64 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const u64 k = kbase + j;
        u64 x = 0;
        if (n < N && k < K) {
            x = B0[k * N + n];
            if (first) x += B1[k * N + n];
        }
        v[j] = x;
    }
    u8* chunk = out + (nt * (2 * kblocks_half) + kb) * B_CHUNK + kc * 8192 + (nl >> 3) * 128 + (nl & 7) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u32 w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = limb4(v + 4 * q, i);
        *reinterpret_cast<uint4*>(chunk + i * 1024) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ------------------------------------------------------------------ the GEMM -----
struct Params {
    const u8* pa;       // packed A: [mtiles][kblocks][A_CHUNK]
    const u8* pb;       // packed B: [ntiles][kblocks][B_CHUNK]
    u64* C;             // row-major, leading dimension N
    u64 row0;           // first row of C covered by tile row 0
    u64 rows_end;       // one past the last valid row (global)
    u64 N;
    u32 mtiles, ntiles, kblocks;
    int accumulate;
    int l2hint;         // 0 none; 1: A panels evict_last, B panels evict_first; 2: A panels evict_last only (ABY3CU_GEMM_L2HINT)
    u32* progress;      // may be NULL: progress[g] += 1 per epilogue warp per finished tile of raster group g
};

__global__ void __launch_bounds__(THREADS, 1) k_gemm_tc(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) u8 smem[];
    const u32 sbase = (smem_u32(smem) + 1023u) & ~1023u;          // stage buffers, 1 KiB aligned
    u8* sgen = smem + (sbase - smem_u32(smem));
    const u32 bar_base = sbase + STAGES * STAGE_BYTES;             // full[S], empty[S], tmem_full, tmem_empty
    u32* tmem_slot = reinterpret_cast<u32*>(sgen + STAGES * STAGE_BYTES + 128);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const u32 tmem_full = bar_base + 8u * (2 * STAGES), tmem_empty = bar_base + 8u * (2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem = *tmem_slot;

    const u32 ntiles_total = p.mtiles * p.ntiles;
    // Tile order: groups of RASTER_M tile rows, column-major inside a group, so that the ~148 tiles in flight share
    // 8 A panels and ~19 B panels through L2 instead of streaming all of B for every wave.
    auto tile_coords = [&](u32 t, u32& mt, u32& nt) {
        const u32 group = RASTER_M * p.ntiles, g = t / group, r = t - g * group;
        const u32 gm = (p.mtiles - g * RASTER_M < RASTER_M) ? (p.mtiles - g * RASTER_M) : RASTER_M;
        mt = g * RASTER_M + r % gm;
        nt = r / gm;
    };

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            u32 stage = 0, phase = 0;
            // A raster group's 8 A panels (64 MiB at K = 4096) are re-read by every tile column of the group and fit L2; the
            // B panels stream through once per group: tell L2 which of the two to keep
            const u64 pol_a = l2_policy_evict_last(), pol_b = l2_policy_evict_first();
            for (u32 t = blockIdx.x; t < ntiles_total; t += gridDim.x) {
                u32 mt, nt;
                tile_coords(t, mt, nt);
                const u8* a = p.pa + (u64)mt * p.kblocks * A_CHUNK;
                const u8* b = p.pb + (u64)nt * p.kblocks * B_CHUNK;
                for (u32 kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    const u32 dst = sbase + stage * STAGE_BYTES;
                    if (p.l2hint) bulk_g2s_hint(dst, a + (u64)kb * A_CHUNK, A_CHUNK, full_bar(stage), pol_a);
                    else bulk_g2s(dst, a + (u64)kb * A_CHUNK, A_CHUNK, full_bar(stage));
                    if (p.l2hint == 1) bulk_g2s_hint(dst + A_CHUNK, b + (u64)kb * B_CHUNK, B_CHUNK, full_bar(stage), pol_b);
                    else bulk_g2s(dst + A_CHUNK, b + (u64)kb * B_CHUNK, B_CHUNK, full_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            u32 stage = 0, phase = 0, tile_iter = 0;
            for (u32 t = blockIdx.x; t < ntiles_total; t += gridDim.x, ++tile_iter) {
                // the epilogue must have drained the previous tile's accumulators
                mbar_wait(tmem_empty, (tile_iter & 1) ^ 1);
                tc_fence_after();
                for (u32 kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const u32 sa = sbase + stage * STAGE_BYTES, sb = sa + A_CHUNK;
                    const u64 bd0 = smem_desc(sb, 8192, 128);            // B rows 0..
                    const u64 bd1 = smem_desc(sb + 4096, 8192, 128);     // B rows 256..
                    const u32 first = kb != 0;                            // kb 0 overwrites all 512 columns via limb 0
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const u64 ad = smem_desc(sa + i * 4096, 2048, 128);
                        const int n = TN * (8 - i);
                        const u32 acc = (i == 0) ? first : 1u;
                        if (n > 256) {
                            // two equal halves (256, 224, 192, 160 wide) instead of 256 + a narrow remainder
                            const int h = (kBalancedSplit ? n / 2 : 256);
                            mma_i8(tmem + 64 * i, ad, bd0, make_idesc(h), acc);
                            mma_i8(tmem + 64 * i + h, ad, kBalancedSplit ? smem_desc(sb + h * 16, 8192, 128) : bd1, make_idesc(n - h), acc);
                        } else {
                            mma_i8(tmem + 64 * i, ad, bd0, make_idesc(n), acc);
                        }
                    }
                    tc_commit(empty_bar(stage));                           // frees the stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tmem_full);                                      // accumulators complete
            }
        }
    } else {
        // ================================ epilogue ====================================
        const u32 quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31
        const u32 row_in_tile = quarter * 32 + lane;
        u32 tile_iter = 0;
        for (u32 t = blockIdx.x; t < ntiles_total; t += gridDim.x, ++tile_iter) {
            u32 mt, nt;
            tile_coords(t, mt, nt);
            const u64 grow = p.row0 + (u64)mt * TM + row_in_tile;
            const u64 gcol0 = (u64)nt * TN;
            const bool rowok = grow < p.rows_end;
            const bool fast = gcol0 + TN <= p.N && ((p.N & 1) == 0);
            u64* dst = p.C + grow * p.N + gcol0;
            // The tile's C values (z or -r pre-loaded by the caller, or an earlier K chunk) are fetched into
            // registers WHILE the MMAs of this tile still run: one thread owns one row, so these loads are 64-byte
            // pieces 8*N bytes apart -- latency-bound, and otherwise exposed after the accumulators complete.
            u64 cpre[TN];
            if (p.accumulate && rowok) {
                if (fast) {
#pragma unroll
                    for (int j = 0; j < TN; j += 2) {
                        const ulonglong2 o = __ldcs(reinterpret_cast<const ulonglong2*>(dst + j));
                        cpre[j] = o.x; cpre[j + 1] = o.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < TN; ++j) cpre[j] = (gcol0 + j < p.N) ? dst[j] : 0;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) cpre[j] = 0;
            }
            mbar_wait(tmem_full, tile_iter & 1);
            tc_fence_after();
            const u32 tlane = tmem + ((quarter * 32u) << 16);
#pragma unroll
            for (int c0 = 0; c0 < TN; c0 += 8) {
                u32 d[8][8];
#pragma unroll
                for (int s = 0; s < 8; ++s) tmem_ld8(tlane + 64 * s + c0, d[s]);
                tmem_ld_wait();
                u64 out[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    u64 acc = cpre[c0 + j];
#pragma unroll
                    for (int s = 0; s < 8; ++s) acc += (u64)d[s][j] << (8 * s);
                    out[j] = acc;
                }
                if (rowok) {
                    if (fast) {
#pragma unroll
                        for (int j = 0; j < 8; j += 2)
                            __stcs(reinterpret_cast<ulonglong2*>(dst + c0 + j), make_ulonglong2(out[j], out[j + 1]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (gcol0 + c0 + j < p.N) dst[c0 + j] = out[j];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tmem_empty);
            if (p.progress) {
                // this warp's 32 rows of the tile are in memory: count them towards the tile's raster group
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(p.progress + t / (RASTER_M * p.ntiles), 1u);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace tc

bool gemm_tc_profitable(u64 M, u64 K, u64 N) {
    // a 128x64 tile must be reasonably full and the contraction long enough to
    // amortise the limb-split pre-pass; skinny shapes go to the CUDA-core kernels
    return N >= 32 && M >= 64 && K >= 64 && M * N * K >= (1ull << 21);
}

// cuStreamWaitValue32 through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned int, unsigned int);
static StreamWaitValue32Fn stream_wait_value32() {
    static StreamWaitValue32Fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        (void)cudaGetLastError();
        return (StreamWaitValue32Fn)p;
    }();
    return fn;
}
constexpr u32 kProgressSlots = 1024;

static int ensure_ws(aby3cu_ctx* ctx, size_t bytes) {
    if (ctx->gemm_ws.bytes >= bytes) return 0;
    {   // growing means a host synchronisation and a driver allocation: not inside a stream capture
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(ctx->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
            set_error("gemm_cross(tcgen05): the limb workspace would have to grow during a stream capture; run the shape once before capturing");
            return 2;
        }
    }
    if (ctx->gemm_ws.ptr) {
        ABY3CU_CHECK(cudaStreamSynchronize(ctx->stream));
        ABY3CU_CHECK(cudaFree(ctx->gemm_ws.ptr));
        ctx->gemm_ws.ptr = nullptr; ctx->gemm_ws.bytes = 0;
    }
    ABY3CU_CHECK(cudaMalloc(&ctx->gemm_ws.ptr, bytes));
    ctx->gemm_ws.bytes = bytes;
    return 0;
}

int gemm_cross_tc(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                  u64 M, u64 K, u64 N, i64* C, int accumulate) {
    using namespace tc;
    ABY3CU_REQUIRE((reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_cross(tcgen05): C must be 16-byte aligned");
    // function attributes are per device: set once per device, not on every call
    static std::mutex attr_mtx;
    static bool attr_done[64] = {};
    {
        std::lock_guard<std::mutex> g(attr_mtx);
        if (ctx->device < 0 || ctx->device >= 64 || !attr_done[ctx->device]) {
            ABY3CU_CHECK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            if (prefer_max_smem(k_gemm_tc)) return 1;
            // the limb pre-pass of ANOTHER party is meant to run beside a GEMM CTA: same carve-out, or the SM has to drain first
            if (prefer_max_smem(k_pack_a) || prefer_max_smem(k_pack_b)) return 1;
            if (ctx->device >= 0 && ctx->device < 64) attr_done[ctx->device] = true;
        }
    }
    const u64 ntiles = (N + TN - 1) / TN;
    // bound the limb-plane workspace: B panel for one K chunk + A panel for one row block
    // (ABY3CU_WS_LIMIT_MB shrinks it so that tests can exercise the row-block loop)
    size_t WS_A_LIMIT = 1ull << 30;
    if (const char* e = getenv("ABY3CU_WS_LIMIT_MB")) {
        const long mb = atol(e);
        if (mb > 0) WS_A_LIMIT = (size_t)mb << 20;
    }
    // 256-bit loads of A need 32-byte aligned rows
    const int vec_a = ((reinterpret_cast<uintptr_t>(A0) | reinterpret_cast<uintptr_t>(A1)) & 31) == 0 && K % 4 == 0;
    for (u64 k0 = 0; k0 < K; k0 += K_MAX) {
        const u64 kc = (K - k0 < K_MAX) ? (K - k0) : K_MAX;
        const u64 kbh = (kc + TK - 1) / TK, kblocks = 2 * kbh;
        const size_t b_bytes = (size_t)ntiles * kblocks * B_CHUNK;
        u64 rows_per_block = (WS_A_LIMIT / ((size_t)kblocks * A_CHUNK)) * TM;
        // aby3cu_gemm_cross_blocks: the caller wants C in row blocks of blk_rows (a multiple of the tile height) with an
        // event after each; the workspace bound may only make the launches smaller, as long as they divide a block
        const bool blocks = ctx->blk_rows && ctx->blk_events && ctx->blk_rows % TM == 0;
        // ONE launch when the whole product fits the workspace and the blocks are whole raster groups: the kernel counts
        // finished tiles per group and a helper stream fires the block events (no wave quantisation per block);
        // otherwise one launch per block
        const u64 group_rows = (u64)RASTER_M * TM;
        const bool progress = blocks && K <= K_MAX && rows_per_block >= M && ctx->blk_rows % group_rows == 0 &&
                              (M + group_rows - 1) / group_rows <= kProgressSlots && stream_wait_value32() != nullptr &&
                              !getenv("ABY3CU_NO_PROGRESS");
        if (blocks && !progress && rows_per_block > ctx->blk_rows) rows_per_block = ctx->blk_rows;
        if (blocks && !progress && ctx->blk_rows % rows_per_block) rows_per_block = TM;
        if (rows_per_block < TM) rows_per_block = TM;
        if (progress) {
            if (!ctx->progress) {
                ABY3CU_CHECK(cudaMalloc(&ctx->progress, kProgressSlots * sizeof(u32)));
                ABY3CU_CHECK(cudaStreamCreateWithFlags(&ctx->progress_stream, cudaStreamNonBlocking));
                ABY3CU_CHECK(cudaEventCreateWithFlags(&ctx->progress_reset, cudaEventDisableTiming));
            }
            // the helper stream must have fired the previous product's events before the counters are cleared
            ABY3CU_CHECK(cudaEventRecord(ctx->progress_reset, ctx->progress_stream));
            ABY3CU_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->progress_reset, 0));
            ABY3CU_CHECK(cudaMemsetAsync(ctx->progress, 0, kProgressSlots * sizeof(u32), ctx->stream));
            ABY3CU_CHECK(cudaEventRecord(ctx->progress_reset, ctx->stream));
            ABY3CU_CHECK(cudaStreamWaitEvent(ctx->progress_stream, ctx->progress_reset, 0));
            const u64 mt_all = (M + TM - 1) / TM, gpb = ctx->blk_rows / group_rows, ngroups = (mt_all + RASTER_M - 1) / RASTER_M;
            for (u32 b = 0; b < ctx->blk_n; ++b) {
                for (u64 g = b * gpb; g < (b + 1) * gpb && g < ngroups; ++g) {
                    const u64 gm = (mt_all - g * RASTER_M < RASTER_M) ? (mt_all - g * RASTER_M) : RASTER_M;
                    const u32 target = (u32)(gm * ntiles * 4);            // four epilogue warps per tile
                    if (stream_wait_value32()(ctx->progress_stream, (unsigned long long)(uintptr_t)(ctx->progress + g), target, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) {
                        set_error("cuStreamWaitValue32 failed");
                        return 1;
                    }
                }
                ABY3CU_CHECK(cudaEventRecord(ctx->blk_events[b], ctx->progress_stream));
            }
            ctx->blk_done = ctx->blk_n;          // all events are armed
        }
        if (rows_per_block > M) rows_per_block = ((M + TM - 1) / TM) * TM;
        const size_t a_bytes = (size_t)(rows_per_block / TM) * kblocks * A_CHUNK;
        if (ensure_ws(ctx, a_bytes + b_bytes)) return 1;
        u8* pa = (u8*)ctx->gemm_ws.ptr;
        u8* pb = pa + a_bytes;
        k_pack_b<<<dim3((unsigned)kblocks, (unsigned)ntiles), 128, 0, ctx->stream>>>((const u64*)B0, (const u64*)B1, K, N, k0, kbh, pb);
        if (post_launch(ctx, "k_pack_b")) return 1;
        const int acc_this = accumulate || k0 > 0;
        for (u64 r0 = 0; r0 < M; r0 += rows_per_block) {
            const u64 rows = (M - r0 < rows_per_block) ? (M - r0) : rows_per_block;
            const u64 mtiles = (rows + TM - 1) / TM;
            k_pack_a<<<dim3((unsigned)kblocks, (unsigned)mtiles), 128, 0, ctx->stream>>>((const u64*)A0, (const u64*)A1, r0, rows, M, K, k0, kbh, pa, vec_a);
            if (post_launch(ctx, "k_pack_a")) return 1;
            Params p;
            p.pa = pa; p.pb = pb; p.C = (u64*)C; p.row0 = r0; p.rows_end = r0 + rows; p.N = N;
            p.mtiles = (u32)mtiles; p.ntiles = (u32)ntiles; p.kblocks = (u32)kblocks; p.accumulate = acc_this;
            p.progress = progress ? ctx->progress : nullptr;
            static const int l2hint = [] { const char* e = getenv("ABY3CU_GEMM_L2HINT"); return e ? atoi(e) : 0; }();
            p.l2hint = l2hint;
            const u64 tiles = mtiles * ntiles;
            // ABY3CU_GEMM_SMS=n: the persistent grid leaves SMs free for the other parties' keystream / pre-pass kernels (the GEMM
            // is bound by the board's power limit, not by its SM count: DESIGN 3.1)
            static const u64 gemm_sms = [] { const char* e = getenv("ABY3CU_GEMM_SMS"); const long v = e ? atol(e) : 0; return (u64)(v > 0 ? v : 0); }();
            const u64 sms = (gemm_sms && gemm_sms < (u64)ctx->sm_count) ? gemm_sms : (u64)ctx->sm_count;
            const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
            if (ctx->c_ready) { ABY3CU_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->c_ready, 0)); ctx->c_ready = nullptr; }   // after the limb pre-pass
            if (k0 == 0 && r0 == 0) ABY3CU_CHECK(cudaEventRecord(ctx->ev_gemm0, ctx->stream));
            trace_mark(ctx, "(k_gemm_tc ready)");
            k_gemm_tc<<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(p);
            if (post_launch(ctx, "k_gemm_tc")) return 1;
            ABY3CU_CHECK(cudaEventRecord(ctx->ev_gemm1, ctx->stream));
            // the last K chunk makes the rows final: signal every block that is complete now
            if (blocks && !progress && k0 + K_MAX >= K) {
                const u64 done_rows = r0 + rows;
                while (ctx->blk_done < ctx->blk_n && ((u64)(ctx->blk_done + 1) * ctx->blk_rows <= done_rows || done_rows >= M)) {
                    ABY3CU_CHECK(cudaEventRecord(ctx->blk_events[ctx->blk_done], ctx->stream));
                    ++ctx->blk_done;
                }
            }
        }
    }
    return 0;
}

}  // namespace aby3cu
