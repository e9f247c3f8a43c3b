// aes.cuh -- AES-128 for the device-side keystreams.
//
// The reference gets its randomness from cryptoTools' AES-NI code
// (oc::AES::ecbEncCounterMode, called at aby3/sh3/Sh3ShareGen.h:52-53 and
// Sh3BinaryEvaluator.cpp:1420-1421).  B200 has no AES instruction, so the device
// version is the classic T-table formulation, arranged for the SM:
//   * one combined table in shared memory, 256 entries x 64 words: words 0..31
//     of an entry hold Te0[x] replicated once per lane, words 32..63 hold
//     Te1[x] = rotl8(Te0[x]).  Lane l only ever touches bank l, so every LDS is
//     conflict-free whatever the data;
//   * Te2/Te3 are rotl16 of Te0/Te1, and rotation is linear over xor, so each
//     output column costs 4 LDS + 1 rotate instead of 4 LDS + 3 rotates;
//   * the whole table offset lane*4 + (byte << 8) comes out of one PRMT and the
//     table base is an LDS immediate;
//   * round keys arrive as __grid_constant__ kernel parameters (constant bank).
// State words are little-endian columns: w_c = b[4c] | b[4c+1]<<8 | ...
#pragma once
#include "common.cuh"

namespace aby3cu {

struct AesKey {
    u32 rk[44];   // rk[4*r + c], little-endian column words
};

constexpr int kAesTableWords = 256 * 64;              // 64 KiB
constexpr int kAesTableBytes = kAesTableWords * 4;

// host: S-box from its definition, key schedule (FIPS-197 5.2)
void host_sbox(u8 sbox[256]);
void host_expand_key(const u8 key[16], AesKey* out);
void host_encrypt_block(const AesKey& k, const u8 in[16], u8 out[16]);
// uploads Te0 to the current device's constant memory (once per device)
int upload_aes_constants();

#ifdef __CUDACC__

// Te0[x] = (2S, S, S, 3S) as little-endian bytes; uploaded by upload_aes_constants().
// The library is built as ONE translation unit (aby3cu_all.cu), so a plain
// definition here is the single definition.
__constant__ u32 c_Te0[256];

// Dynamic shared memory of every AES-using kernel starts with the 64 KiB table.
// Addressing it through this symbol (not through a pointer variable) lets ptxas
// fold the table base into the LDS immediate: one PRMT builds the whole offset
// lane*4 + (byte << 8), and the lookup is a single LDS [R + imm].
extern __shared__ __align__(16) u32 aby3_smem[];

// Fill the shared-memory table.  Call from every thread of the CTA, then
// __syncthreads().
__device__ __forceinline__ void aes_table_init() {
    for (int i = threadIdx.x; i < kAesTableWords; i += blockDim.x) {
        const u32 t0 = c_Te0[i >> 6];
        aby3_smem[i] = (i & 32) ? __byte_perm(t0, 0, 0x2103) : t0;   // rotl8
    }
}

// lane4 = (threadIdx.x & 31) * 4 (< 128, so its bytes 1..3 are zero).
// Table byte offset of entry x.byte_K for this lane: lane4 | (byte << 8).
template <int K>
__device__ __forceinline__ u32 tix(u32 x, u32 lane4) {
    return __byte_perm(x, lane4, 0x6504 | (K << 4));
}
// Te0 lookup (TBL = 0) or Te1 lookup (TBL = 1, 128 bytes further)
template <int TBL>
__device__ __forceinline__ u32 tlu(u32 off) {
    return *reinterpret_cast<const u32*>(reinterpret_cast<const char*>(aby3_smem) + off + TBL * 128);
}

// Encrypts the counter block toBlock(ctr) = (ctr as 8 LE bytes, 8 zero bytes).
// out[0..3] are the four little-endian words of the ciphertext (bytes 0..15).
__device__ __forceinline__ void aes_encrypt_ctr(u32 lane4, const AesKey& key, u64 ctr, u32 out[4]) {
    u32 s0 = (u32)ctr ^ key.rk[0];
    u32 s1 = (u32)(ctr >> 32) ^ key.rk[1];
    u32 s2 = key.rk[2];
    u32 s3 = key.rk[3];
#pragma unroll
    for (int r = 1; r < 10; ++r) {
        u32 a0 = tlu<0>(tix<0>(s0, lane4)), b0 = tlu<1>(tix<1>(s1, lane4));
        u32 c0 = tlu<0>(tix<2>(s2, lane4)), d0 = tlu<1>(tix<3>(s3, lane4));
        u32 a1 = tlu<0>(tix<0>(s1, lane4)), b1 = tlu<1>(tix<1>(s2, lane4));
        u32 c1 = tlu<0>(tix<2>(s3, lane4)), d1 = tlu<1>(tix<3>(s0, lane4));
        u32 a2 = tlu<0>(tix<0>(s2, lane4)), b2 = tlu<1>(tix<1>(s3, lane4));
        u32 c2 = tlu<0>(tix<2>(s0, lane4)), d2 = tlu<1>(tix<3>(s1, lane4));
        u32 a3 = tlu<0>(tix<0>(s3, lane4)), b3 = tlu<1>(tix<1>(s0, lane4));
        u32 c3 = tlu<0>(tix<2>(s1, lane4)), d3 = tlu<1>(tix<3>(s2, lane4));
        s0 = a0 ^ b0 ^ __byte_perm(c0 ^ d0, 0, 0x1032) ^ key.rk[4 * r + 0];
        s1 = a1 ^ b1 ^ __byte_perm(c1 ^ d1, 0, 0x1032) ^ key.rk[4 * r + 1];
        s2 = a2 ^ b2 ^ __byte_perm(c2 ^ d2, 0, 0x1032) ^ key.rk[4 * r + 2];
        s3 = a3 ^ b3 ^ __byte_perm(c3 ^ d3, 0, 0x1032) ^ key.rk[4 * r + 3];
    }
    // last round: SubBytes + ShiftRows only.  Te0[x] = [2S,S,S,3S], Te1[x] = [3S,2S,S,S]
    // (bytes 0..3), so S sits in byte 1 of Te0 and in bytes 2,3 of Te1.
#define ABY3CU_LAST(o, x0, x1, x2, x3, kk)                                          \
    {                                                                               \
        u32 u = tlu<0>(tix<0>(x0, lane4)), v = tlu<0>(tix<1>(x1, lane4));           \
        u32 w = tlu<1>(tix<2>(x2, lane4)), y = tlu<1>(tix<3>(x3, lane4));           \
        u32 lo = __byte_perm(u, v, 0x0051);                                         \
        u32 hi = __byte_perm(w, y, 0x7200);                                         \
        o = __byte_perm(lo, hi, 0x7610) ^ key.rk[kk];                               \
    }
    ABY3CU_LAST(out[0], s0, s1, s2, s3, 40)
    ABY3CU_LAST(out[1], s1, s2, s3, s0, 41)
    ABY3CU_LAST(out[2], s2, s3, s0, s1, 42)
    ABY3CU_LAST(out[3], s3, s0, s1, s2, 43)
#undef ABY3CU_LAST
}

// keystream elements e and e+1 (little-endian u64 at keystream bytes 8e.. and 8e+8..)
__device__ __forceinline__ void aes_stream_pair(u32 lane4, const AesKey& key, u64 e, u64& v0, u64& v1) {
    u32 o[4];
    aes_encrypt_ctr(lane4, key, e >> 1, o);
    if ((e & 1) == 0) {
        v0 = ((u64)o[1] << 32) | o[0];
        v1 = ((u64)o[3] << 32) | o[2];
    } else {
        u32 p[4];
        aes_encrypt_ctr(lane4, key, (e >> 1) + 1, p);
        v0 = ((u64)o[3] << 32) | o[2];
        v1 = ((u64)p[1] << 32) | p[0];
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// The WIDE form: four tables and counter-mode round folding.
//
// Bit slicing does not pay on this machine: LOP3 / PRMT issue on the alu pipe at 16 lanes per SM sub-partition per clock
// (64 per SM), a bit-sliced AES needs ~750 logic operations per block, i.e. ~12 SM cycles per block against the 5.5 of
// the table form above -- which is itself balanced between the load pipe (160 four-byte lookups per block through a
// 128 B/clk shared-memory port = 5.0 cycles) and the alu pipe (312 operations = 4.9 cycles).  So the way forward is fewer
// lookups and fewer logic operations per block:
//   * all four tables Te0..Te3 (128 KiB, lane-replicated as above): an output column is 4 lookups and two 3-input xors,
//     no rotate and no extra xor (240 instead of 312 alu operations per block);
//   * CTR folding: the counter block is toBlock(ctr) = 8 little-endian counter bytes + 8 zero bytes, so within a run of 256
//     consecutive counters only byte 0 of the state differs.  After round 1 only column 0 depends on it (one lookup), after
//     round 2 each column depends on it through ONE byte of that column (four lookups); everything else is a per-run
//     constant (27 lookups, amortised over the run).  Rounds 1-2 cost 5 lookups instead of 32: 133 + 27/run per block.
// A lane keeps the constants of the run its current counter lies in (AesRun) and refreshes them when the counter leaves it.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kAesWideTableWords = 256 * 128;             // 128 KiB
constexpr int kAesWideTableBytes = kAesWideTableWords * 4;

// region 0 (64 KiB): entry x = [32 x Te0[x] | 32 x Te1[x]]; region 1: [32 x Te2[x] | 32 x Te3[x]].  Te_k = rotl(Te0, 8k).
__device__ __forceinline__ void aes_wide_table_init() {
    for (int i = threadIdx.x; i < kAesWideTableWords; i += blockDim.x) {
        const u32 t0 = c_Te0[(i >> 6) & 255];
        const int tbl = ((i >> 14) << 1) | ((i >> 5) & 1);
        const u32 sel = tbl == 0 ? 0x3210u : tbl == 1 ? 0x2103u : tbl == 2 ? 0x1032u : 0x0321u;
        aby3_smem[i] = __byte_perm(t0, 0, sel);
    }
}
template <int TBL>
__device__ __forceinline__ u32 tlw(u32 off) {
    return *reinterpret_cast<const u32*>(reinterpret_cast<const char*>(aby3_smem) + off + (TBL >> 1) * 65536 + (TBL & 1) * 128);
}
// T_TBL[byte K of x]
template <int TBL, int K>
__device__ __forceinline__ u32 tw(u32 x, u32 lane4) { return tlw<TBL>(tix<K>(x, lane4)); }

__device__ __forceinline__ u32 xor3(u32 a, u32 b, u32 c) {
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

struct AesRun {
    u64 hi = ~0ull;        // ctr >> 8 of the run the constants below belong to
    u32 k0;                // rk[0] ^ (low counter word with byte 0 cleared): column 0 before round 1 is k0 ^ byte0
    u32 P0, Q0, Q1, Q2, Q3;
};

__device__ __forceinline__ void aes_run_setup(u32 lane4, const AesKey& key, u64 ctr, AesRun& r) {
    r.hi = ctr >> 8;
    const u32 s0 = ((u32)ctr & ~0xFFu) ^ key.rk[0];       // byte 0 of the counter enters per block
    const u32 s1 = (u32)(ctr >> 32) ^ key.rk[1];
    const u32 s2 = key.rk[2], s3 = key.rk[3];
    r.k0 = s0;
    // round 1 without the byte-0 term of column 0
    r.P0 = xor3(tw<1, 1>(s1, lane4), tw<2, 2>(s2, lane4), tw<3, 3>(s3, lane4)) ^ key.rk[4];
    const u32 P1 = xor3(tw<0, 0>(s1, lane4), tw<1, 1>(s2, lane4), tw<2, 2>(s3, lane4)) ^ tw<3, 3>(s0, lane4) ^ key.rk[5];
    const u32 P2 = xor3(tw<0, 0>(s2, lane4), tw<1, 1>(s3, lane4), tw<2, 2>(s0, lane4)) ^ tw<3, 3>(s1, lane4) ^ key.rk[6];
    const u32 P3 = xor3(tw<0, 0>(s3, lane4), tw<1, 1>(s0, lane4), tw<2, 2>(s1, lane4)) ^ tw<3, 3>(s2, lane4) ^ key.rk[7];
    // round 2 without the terms of column 0
    r.Q0 = xor3(tw<1, 1>(P1, lane4), tw<2, 2>(P2, lane4), tw<3, 3>(P3, lane4)) ^ key.rk[8];
    r.Q1 = xor3(tw<0, 0>(P1, lane4), tw<1, 1>(P2, lane4), tw<2, 2>(P3, lane4)) ^ key.rk[9];
    r.Q2 = xor3(tw<0, 0>(P2, lane4), tw<1, 1>(P3, lane4), tw<3, 3>(P1, lane4)) ^ key.rk[10];
    r.Q3 = xor3(tw<0, 0>(P3, lane4), tw<2, 2>(P1, lane4), tw<3, 3>(P2, lane4)) ^ key.rk[11];
}

// the same function as aes_encrypt_ctr, through the wide tables and the run constants
__device__ __forceinline__ void aes_wide_encrypt_ctr(u32 lane4, const AesKey& key, AesRun& run, u64 ctr, u32 out[4]) {
    if ((ctr >> 8) != run.hi) aes_run_setup(lane4, key, ctr, run);
    // round 1: column 0 = Te0[byte 0] ^ P0; columns 1..3 are run constants (folded into Q)
    const u32 w0 = tw<0, 0>(run.k0 ^ ((u32)ctr & 0xFFu), lane4) ^ run.P0;
    // round 2: each column sees column 0 through one byte
    u32 s0 = tw<0, 0>(w0, lane4) ^ run.Q0;
    u32 s1 = tw<3, 3>(w0, lane4) ^ run.Q1;
    u32 s2 = tw<2, 2>(w0, lane4) ^ run.Q2;
    u32 s3 = tw<1, 1>(w0, lane4) ^ run.Q3;
#pragma unroll
    for (int r = 3; r < 10; ++r) {
        const u32 a0 = tw<0, 0>(s0, lane4), b0 = tw<1, 1>(s1, lane4), c0 = tw<2, 2>(s2, lane4), d0 = tw<3, 3>(s3, lane4);
        const u32 a1 = tw<0, 0>(s1, lane4), b1 = tw<1, 1>(s2, lane4), c1 = tw<2, 2>(s3, lane4), d1 = tw<3, 3>(s0, lane4);
        const u32 a2 = tw<0, 0>(s2, lane4), b2 = tw<1, 1>(s3, lane4), c2 = tw<2, 2>(s0, lane4), d2 = tw<3, 3>(s1, lane4);
        const u32 a3 = tw<0, 0>(s3, lane4), b3 = tw<1, 1>(s0, lane4), c3 = tw<2, 2>(s1, lane4), d3 = tw<3, 3>(s2, lane4);
        s0 = xor3(xor3(a0, b0, c0), d0, key.rk[4 * r + 0]);
        s1 = xor3(xor3(a1, b1, c1), d1, key.rk[4 * r + 1]);
        s2 = xor3(xor3(a2, b2, c2), d2, key.rk[4 * r + 2]);
        s3 = xor3(xor3(a3, b3, c3), d3, key.rk[4 * r + 3]);
    }
    // last round: S sits in byte 1 of Te0 and in bytes 2, 3 of Te1 (region 0 keeps the layout of the narrow form)
#define ABY3CU_LASTW(o, x0, x1, x2, x3, kk)                                             \
    {                                                                                   \
        u32 u = tw<0, 0>(x0, lane4), v = tw<0, 1>(x1, lane4);                           \
        u32 w = tw<1, 2>(x2, lane4), y = tw<1, 3>(x3, lane4);                           \
        u32 lo = __byte_perm(u, v, 0x0051);                                             \
        u32 hi = __byte_perm(w, y, 0x7200);                                             \
        o = __byte_perm(lo, hi, 0x7610) ^ key.rk[kk];                                   \
    }
    ABY3CU_LASTW(out[0], s0, s1, s2, s3, 40)
    ABY3CU_LASTW(out[1], s1, s2, s3, s0, 41)
    ABY3CU_LASTW(out[2], s2, s3, s0, s1, 42)
    ABY3CU_LASTW(out[3], s3, s0, s1, s2, 43)
#undef ABY3CU_LASTW
}

// the same function through the four wide tables WITHOUT the run constants: for kernels that keep several keystreams per lane
// and have no registers left for their AesRun states (240 alu operations, 160 lookups per block)
__device__ __forceinline__ void aes_wide_encrypt_plain(u32 lane4, const AesKey& key, u64 ctr, u32 out[4]) {
    u32 s0 = (u32)ctr ^ key.rk[0];
    u32 s1 = (u32)(ctr >> 32) ^ key.rk[1];
    u32 s2 = key.rk[2];
    u32 s3 = key.rk[3];
#pragma unroll
    for (int r = 1; r < 10; ++r) {
        const u32 a0 = tw<0, 0>(s0, lane4), b0 = tw<1, 1>(s1, lane4), c0 = tw<2, 2>(s2, lane4), d0 = tw<3, 3>(s3, lane4);
        const u32 a1 = tw<0, 0>(s1, lane4), b1 = tw<1, 1>(s2, lane4), c1 = tw<2, 2>(s3, lane4), d1 = tw<3, 3>(s0, lane4);
        const u32 a2 = tw<0, 0>(s2, lane4), b2 = tw<1, 1>(s3, lane4), c2 = tw<2, 2>(s0, lane4), d2 = tw<3, 3>(s1, lane4);
        const u32 a3 = tw<0, 0>(s3, lane4), b3 = tw<1, 1>(s0, lane4), c3 = tw<2, 2>(s1, lane4), d3 = tw<3, 3>(s2, lane4);
        s0 = xor3(xor3(a0, b0, c0), d0, key.rk[4 * r + 0]);
        s1 = xor3(xor3(a1, b1, c1), d1, key.rk[4 * r + 1]);
        s2 = xor3(xor3(a2, b2, c2), d2, key.rk[4 * r + 2]);
        s3 = xor3(xor3(a3, b3, c3), d3, key.rk[4 * r + 3]);
    }
#define ABY3CU_LASTW(o, x0, x1, x2, x3, kk)                                             \
    {                                                                                   \
        u32 u = tw<0, 0>(x0, lane4), v = tw<0, 1>(x1, lane4);                           \
        u32 w = tw<1, 2>(x2, lane4), y = tw<1, 3>(x3, lane4);                           \
        u32 lo = __byte_perm(u, v, 0x0051);                                             \
        u32 hi = __byte_perm(w, y, 0x7200);                                             \
        o = __byte_perm(lo, hi, 0x7610) ^ key.rk[kk];                                   \
    }
    ABY3CU_LASTW(out[0], s0, s1, s2, s3, 40)
    ABY3CU_LASTW(out[1], s1, s2, s3, s0, 41)
    ABY3CU_LASTW(out[2], s2, s3, s0, s1, 42)
    ABY3CU_LASTW(out[3], s3, s0, s1, s2, 43)
#undef ABY3CU_LASTW
}

// One interface for both forms.  WIDE needs an even first element (a whole block per pair).
template <bool WIDE>
struct AesStream {
    AesRun run;
    __device__ __forceinline__ void pair(u32 lane4, const AesKey& key, u64 e, u64& v0, u64& v1) {
        if (WIDE) {
            u32 o[4];
            aes_wide_encrypt_ctr(lane4, key, run, e >> 1, o);
            v0 = ((u64)o[1] << 32) | o[0];
            v1 = ((u64)o[3] << 32) | o[2];
        } else {
            aes_stream_pair(lane4, key, e, v0, v1);
        }
    }
    __device__ __forceinline__ void block(u32 lane4, const AesKey& key, u64 ctr, u32 o[4]) {
        if (WIDE) aes_wide_encrypt_ctr(lane4, key, run, ctr, o);
        else aes_encrypt_ctr(lane4, key, ctr, o);
    }
};
template <bool WIDE>
__device__ __forceinline__ void aes_tables_init() {
    if (WIDE) aes_wide_table_init(); else aes_table_init();
}

// Work distribution of the keystream kernels.  Narrow: grid-stride over `items`.  Wide: a warp takes runs of 256
// consecutive items (8 steps of 32 lanes), so that a lane's counter stays inside one 256-counter run for 8 steps.
template <bool WIDE, class Body>
__device__ __forceinline__ void aes_for_each(size_t items, Body body) {
    if (WIDE) {
        const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
        const u32 lane = threadIdx.x & 31;
        for (size_t run = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; run * 256 < items; run += warps) {
#pragma unroll 1
            for (int t = 0; t < 8; ++t) {
                const size_t i = run * 256 + t * 32 + lane;
                if (i < items) body(i);
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (size_t)gridDim.x * blockDim.x) body(i);
    }
}

#endif  // __CUDACC__

}  // namespace aby3cu
