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

#endif  // __CUDACC__

}  // namespace aby3cu
