// binary.cu -- kernels behind Sh3BinaryEvaluator: bit-matrix transposes between
// the row-major sbMatrix layout and the bit-sliced wire memory, the per-level
// gate interpreter with in-register AES-CTR zero shares, and the send-buffer
// pack / receive scatter.  All HBM-bound bitwise work on 128-bit words.
#include <stdlib.h>

#include "aes.cuh"

namespace aby3cu {
namespace {

// ---------------------------------------------------------------------------------
// Bit-matrix transpose (oc::transpose semantics, LSB first).  A CTA stages a
// TR x (32*TCW)-bit tile in shared memory, warps transpose 32x32-bit blocks
// with 32 ballots, and the transposed tile is written back in full words so
// both the global reads and the global writes are coalesced.
// ---------------------------------------------------------------------------------
template <int TR, int TCW>
__global__ void __launch_bounds__(256) k_bit_transpose(const u8* __restrict__ in, const u32* __restrict__ row_index,
                                                       u64 rows, u64 cols, u64 in_stride,
                                                       u8* __restrict__ out, u64 out_stride,
                                                       const u8* __restrict__ invert) {
    constexpr int TRW = TR / 32;
    __shared__ u32 sIn[TR][TCW + 1];
    __shared__ u32 sOut[TCW * 32][TRW + 1];
    const u64 tiles_r = (rows + TR - 1) / TR, tiles_c = (cols + 32 * TCW - 1) / (32 * TCW);
    const u64 in_row_words = (cols + 31) / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (u64 t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
        const u64 r0 = (t / tiles_c) * TR, cw0 = (t % tiles_c) * TCW;
        for (int idx = threadIdx.x; idx < TR * TCW; idx += blockDim.x) {
            const int r = idx / TCW, cw = idx % TCW;
            const u64 gr = r0 + r, gw = cw0 + cw;
            u32 w = 0;
            if (gr < rows && gw < in_row_words) {
                const u64 src_row = row_index ? (u64)row_index[gr] : gr;
                w = *reinterpret_cast<const u32*>(in + src_row * in_stride + gw * 4);
                if (invert && invert[gr]) w = ~w;
            }
            sIn[r][cw] = w;
        }
        __syncthreads();
        for (int b = warp; b < TRW * TCW; b += nwarps) {
            const int rb = b / TCW, cw = b % TCW;
            const u32 w = sIn[rb * 32 + lane][cw];
            u32 mine = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const u32 bal = __ballot_sync(0xffffffffu, (w >> j) & 1u);
                if (lane == j) mine = bal;
            }
            sOut[cw * 32 + lane][rb] = mine;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < TCW * 32 * TRW; idx += blockDim.x) {
            const int orow = idx / TRW, ow = idx % TRW;
            const u64 gc = cw0 * 32 + orow;          // output row = input column
            const u64 gr = r0 + (u64)ow * 32;        // first input row covered by this word
            if (gc < cols && gr < rows)
                *reinterpret_cast<u32*>(out + gc * out_stride + (gr / 32) * 4) = sOut[orow][ow];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Fast paths for the two shapes the binary engine actually uses: rows of 64-bit
// words <-> bit-sliced rows.  A thread owns 8 instances: it holds an 8 x 64 bit
// matrix in registers, transposes it as eight 8x8 bit blocks (three masked
// exchange steps each), and the bytes are staged through shared memory so that
// both global sides move whole 128-byte lines.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ u64 transpose8x8(u64 x) {
    u64 t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x ^= t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x ^= t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x ^= t ^ (t << 28);
    return x;
}
// byte K of each of four 32-bit words -> one word
template <int K>
__device__ __forceinline__ u32 gather_byte(u32 a, u32 b, u32 c, u32 d) {
    const u32 ab = __byte_perm(a, b, K | ((4 + K) << 4));
    const u32 cd = __byte_perm(c, d, K | ((4 + K) << 4));
    return __byte_perm(ab, cd, 0x5410);
}
template <int K>
__device__ __forceinline__ u64 block_of(const u64 r[8]) {
    // byte i of the result = byte K of row i
    constexpr int W = K >> 2, B = K & 3;
    const u32 lo = gather_byte<B>((u32)(r[0] >> (32 * W)), (u32)(r[1] >> (32 * W)), (u32)(r[2] >> (32 * W)), (u32)(r[3] >> (32 * W)));
    const u32 hi = gather_byte<B>((u32)(r[4] >> (32 * W)), (u32)(r[5] >> (32 * W)), (u32)(r[6] >> (32 * W)), (u32)(r[7] >> (32 * W)));
    return ((u64)hi << 32) | lo;
}

// 32 x 32 bit-matrix transpose in registers, LSB-first: out[b] bit i = in[i] bit b.  Five masked exchange stages,
// 16 register pairs each (SHF + LOP3 + LOP3 + SHF + LOP3): 400 ops per 128 bytes.
__device__ __forceinline__ void transpose32(u32 A[32]) {
    constexpr u32 M[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int j = 16 >> s;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k & j) continue;
            const u32 t = ((A[k] >> j) ^ A[k | j]) & M[s];
            A[k] ^= t << j;
            A[k | j] ^= t;
        }
    }
}

// Fast path, 32 instances per thread: the thread's 32 x 64 bit block is two 32 x 32 register transposes, and every
// bit-row leaves as one 32-bit word -- a warp writes 128 contiguous bytes of a row, no shared-memory staging.
// in: `rows` instances x W words (in_stride = 8 W); out row (64 w + b) = bit b of word w of every instance.
__global__ void __launch_bounds__(128) k_bits_to_sliced32(const u64* __restrict__ in, u64 rows, u64 cols, u64 W,
                                                          u8* __restrict__ out, u64 out_stride, int vec) {
    const u64 groups = (rows + 31) / 32;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < groups * W; g += (u64)gridDim.x * blockDim.x) {
        const u64 w = g / groups, t = g - w * groups, i0 = 32 * t;
        u32 lo[32], hi[32];
        if (vec && i0 + 32 <= rows) {                   // W == 1, 16-byte aligned: 16 x LDG.128 of 256 contiguous bytes
            const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + i0);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const ulonglong2 v = p[i];
                lo[2 * i] = (u32)v.x; hi[2 * i] = (u32)(v.x >> 32);
                lo[2 * i + 1] = (u32)v.y; hi[2 * i + 1] = (u32)(v.y >> 32);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const u64 v = (i0 + i < rows) ? in[(i0 + i) * W + w] : 0;
                lo[i] = (u32)v; hi[i] = (u32)(v >> 32);
            }
        }
        transpose32(lo);
        transpose32(hi);
        u8* base = out + 4 * t;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const u64 r0 = 64 * w + b, r1 = r0 + 32;
            if (r0 < cols) *reinterpret_cast<u32*>(base + r0 * out_stride) = lo[b];
            if (r1 < cols) *reinterpret_cast<u32*>(base + r1 * out_stride) = hi[b];
        }
    }
}

constexpr int kFastInst = 2048;      // instances per CTA tile (256 threads x 8)

// in: `rows` instances x W words of 64 bits (in_stride = 8*W bytes); out row (64*w + b) holds
// bit b of word w of every instance.  Only the first `cols` bit-rows are written.
__global__ void __launch_bounds__(256) k_bits_to_sliced(const u64* __restrict__ in, u64 rows, u64 cols, u64 W,
                                                        u8* __restrict__ out, u64 out_stride) {
    __shared__ __align__(16) u8 sOut[64][kFastInst / 8 + 16];
    const u64 tiles_r = (rows + kFastInst - 1) / kFastInst;
    for (u64 t = blockIdx.x; t < tiles_r * W; t += gridDim.x) {
        const u64 r0 = (t / W) * kFastInst, w = t % W;
        const u64 i0 = r0 + (u64)threadIdx.x * 8;
        u64 r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (i0 + i < rows) ? in[(i0 + i) * W + w] : 0;
#define ABY3CU_BLK(K)                                                                   \
        {                                                                               \
            const u64 x = transpose8x8(block_of<K>(r));                                 \
            _Pragma("unroll") for (int j = 0; j < 8; ++j) sOut[8 * K + j][threadIdx.x] = (u8)(x >> (8 * j)); \
        }
        ABY3CU_BLK(0) ABY3CU_BLK(1) ABY3CU_BLK(2) ABY3CU_BLK(3) ABY3CU_BLK(4) ABY3CU_BLK(5) ABY3CU_BLK(6) ABY3CU_BLK(7)
#undef ABY3CU_BLK
        __syncthreads();
        // 64 bit-rows x 256 bytes, written as 16-byte pieces
        const u64 valid_bytes = (rows - r0 + 7) / 8 < (u64)(kFastInst / 8) ? (rows - r0 + 7) / 8 : (u64)(kFastInst / 8);
        for (int idx = threadIdx.x; idx < 64 * (kFastInst / 8 / 16); idx += blockDim.x) {
            const int b = idx / (kFastInst / 8 / 16), piece = idx % (kFastInst / 8 / 16);
            const u64 bitrow = 64 * w + b;
            if (bitrow >= cols || (u64)piece * 16 >= valid_bytes) continue;
            u8* dst = out + bitrow * out_stride + r0 / 8 + piece * 16;
            const uint4 v = *reinterpret_cast<const uint4*>(&sOut[b][piece * 16]);
            // the caller guarantees out_stride covers the 16-byte piece (dispatch in launch_transpose)
            *reinterpret_cast<uint4*>(dst) = v;
        }
        __syncthreads();
    }
}

// inverse: `bits` bit-rows (gathered through row_index, optionally complemented) of `width`
// instances -> width x W words; bits beyond `bits` in the last word are zero.
__global__ void __launch_bounds__(256) k_sliced_to_bits(const u8* __restrict__ in, const u32* __restrict__ row_index, u64 bits,
                                                        u64 width, u64 in_stride, u64* __restrict__ out, u64 W,
                                                        const u8* __restrict__ invert) {
    __shared__ __align__(16) u8 sIn[64][kFastInst / 8 + 16];
    const u64 tiles_c = (width + kFastInst - 1) / kFastInst;
    for (u64 t = blockIdx.x; t < tiles_c * W; t += gridDim.x) {
        const u64 c0 = (t / W) * kFastInst, w = t % W;
        for (int idx = threadIdx.x; idx < 64 * (kFastInst / 8 / 16); idx += blockDim.x) {
            const int b = idx / (kFastInst / 8 / 16), piece = idx % (kFastInst / 8 / 16);
            const u64 bitrow = 64 * w + b;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (bitrow < bits && c0 / 8 + (u64)piece * 16 < in_stride) {
                const u64 src_row = row_index ? (u64)row_index[bitrow] : bitrow;
                v = *reinterpret_cast<const uint4*>(in + src_row * in_stride + c0 / 8 + piece * 16);
                if (invert && invert[bitrow]) { v.x = ~v.x; v.y = ~v.y; v.z = ~v.z; v.w = ~v.w; }
            }
            *reinterpret_cast<uint4*>(&sIn[b][piece * 16]) = v;
        }
        __syncthreads();
        const u64 i0 = c0 + (u64)threadIdx.x * 8;
        u64 blk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            u64 x = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) x |= (u64)sIn[8 * k + j][threadIdx.x] << (8 * j);
            blk[k] = transpose8x8(x);        // byte i = bits 8k..8k+7 of instance i
        }
        u64 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i] |= ((blk[k] >> (8 * i)) & 0xFFull) << (8 * k);
        }
        if (W == 1) {
            // the tile's 2048 output words are contiguous: stage them through the (now free) shared tile so that a warp
            // stores 256 contiguous bytes per instruction instead of 32 eight-byte pieces 64 bytes apart
            __syncthreads();
            u64* so = reinterpret_cast<u64*>(&sIn[0][0]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const u32 n = 8 * threadIdx.x + i; so[n + (n >> 5)] = v[i]; }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const u32 n = j * 256 + threadIdx.x;
                if (c0 + n < width) out[c0 + n] = so[n + (n >> 5)];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i0 + i < width) out[(i0 + i) * W + w] = v[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// One AND-depth level.  A thread owns one 16-byte column chunk of every wire row
// and walks the level's gate list in order, so chains of linear gates inside a
// level see their own earlier writes (same thread, same addresses).
// ---------------------------------------------------------------------------------
struct alignas(16) W2 { u64 a, b; };

__device__ __forceinline__ W2 ldw(const u64* mem, u64 row, u64 rw, u64 c) { return *reinterpret_cast<const W2*>(mem + row * rw + 2 * c); }
__device__ __forceinline__ void stw(u64* mem, u64 row, u64 rw, u64 c, W2 v) { *reinterpret_cast<W2*>(mem + row * rw + 2 * c) = v; }

// ONE: plane 0 only, linear gates only -- the second plane of a wire is the previous party's first plane (replicated sharing),
// so co-located parties read it there instead of recomputing it (Sh3BinaryEvaluator "shared planes")
template <bool AES, bool ONE = false>
__global__ void __launch_bounds__(256) k_bin_level(const uint4* __restrict__ gates, u32 n_gates, u64* mem0, u64* mem1, u64 rw,
                                                   const __grid_constant__ AesKey kp, const __grid_constant__ AesKey kn, u64 and0) {
    if (AES) { aes_table_init(); __syncthreads(); }
    const u32 Tl = (threadIdx.x & 31) * 4;
    const u64 chunks = rw / 2;
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (u64)gridDim.x * blockDim.x) {
        u64 and_idx = and0;
        for (u32 g = 0; g < n_gates; ++g) {
            const uint4 G = gates[g];
            const u32 type = G.w;
            if (ONE) {
                const W2 a = ldw(mem0, G.x, rw, c);
                W2 o = a;
                if (type != 10) {
                    const W2 b = ldw(mem0, G.y, rw, c);
                    o = {a.a ^ b.a, a.b ^ b.b};
                    if (type == 9) o = {~o.a, ~o.b};
                }
                stw(mem0, G.z, rw, c, o);
                continue;
            }
            const W2 a0 = ldw(mem0, G.x, rw, c), a1 = ldw(mem1, G.x, rw, c);
            W2 b0 = {0, 0}, b1 = {0, 0};
            if (type != 10) { b0 = ldw(mem0, G.y, rw, c); b1 = ldw(mem1, G.y, rw, c); }
            W2 o0, o1;
            bool linear = true;
            switch (type) {
            case 6:  o0 = {a0.a ^ b0.a, a0.b ^ b0.b}; o1 = {a1.a ^ b1.a, a1.b ^ b1.b}; break;          // Xor
            case 9:  o0 = {~(a0.a ^ b0.a), ~(a0.b ^ b0.b)}; o1 = {~(a1.a ^ b1.a), ~(a1.b ^ b1.b)}; break; // Nxor
            case 10: o0 = a0; o1 = a1; break;                                                          // copy
            default: {
                linear = false;
                W2 x0 = a0, x1 = a1, y0 = b0, y1 = b1;
                if (type == 1) { x0 = {~a0.a, ~a0.b}; x1 = {~a1.a, ~a1.b}; y0 = {~b0.a, ~b0.b}; y1 = {~b1.a, ~b1.b}; }  // Nor
                else if (type == 4) { x0 = {~a0.a, ~a0.b}; x1 = {~a1.a, ~a1.b}; }                                        // na_And
                o0.a = (x0.a & y0.a) ^ (x0.a & y1.a) ^ (x1.a & y0.a);
                o0.b = (x0.b & y0.b) ^ (x0.b & y1.b) ^ (x1.b & y0.b);
                if (type == 14) { o0.a ^= a0.a ^ b0.a; o0.b ^= a0.b ^ b0.b; }                                            // Or
                if (AES) {
                    u32 p[4], q[4];
                    const u64 ctr = and_idx * chunks + c;
                    aes_encrypt_ctr(Tl, kp, ctr, p);
                    aes_encrypt_ctr(Tl, kn, ctr, q);
                    o0.a ^= (((u64)(p[1] ^ q[1])) << 32) | (u64)(p[0] ^ q[0]);
                    o0.b ^= (((u64)(p[3] ^ q[3])) << 32) | (u64)(p[2] ^ q[2]);
                }
                ++and_idx;
                o1 = {0, 0};
            }
            }
            stw(mem0, G.z, rw, c, o0);
            if (linear) stw(mem1, G.z, rw, c, o1);
        }
    }
}

// The linear gates of one level on plane 0 only (shared planes), with memory-level parallelism: the in-order interpreter above
// is one dependent chain of "load descriptor -> load operands -> store" per gate (15-27 us for the ~20 gates of a level, whatever
// the row size).  The host cuts the list into batches of up to 8 mutually independent gates (no gate of a batch reads or writes a
// wire another gate of the batch writes; `first[g]` != 0 starts a batch): all operand rows of a batch are loaded before any of its
// outputs is stored, so up to 16 row loads per thread are in flight.  Same values as the in-order walk.
constexpr int kLinBatch = 8;
__global__ void __launch_bounds__(256) k_bin_linear_plane0(const uint4* __restrict__ gates, const u8* __restrict__ first, u32 n_gates,
                                                           u64* mem0, u64 rw) {
    const u64 chunks = rw / 2;
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (u64)gridDim.x * blockDim.x) {
        u32 g = 0;
        while (g < n_gates) {
            u32 e = g + 1;
            while (e < n_gates && e - g < kLinBatch && !first[e]) ++e;
            const u32 cnt = e - g;
            uint4 G[kLinBatch];
            W2 a[kLinBatch], b[kLinBatch];
#pragma unroll
            for (int k = 0; k < kLinBatch; ++k)
                if ((u32)k < cnt) {
                    G[k] = __ldg(gates + g + k);
                    a[k] = ldw(mem0, G[k].x, rw, c);
                    b[k] = (G[k].w != 10) ? ldw(mem0, G[k].y, rw, c) : W2{0, 0};
                }
#pragma unroll
            for (int k = 0; k < kLinBatch; ++k)
                if ((u32)k < cnt) {
                    W2 o = a[k];
                    if (G[k].w != 10) {
                        o = {a[k].a ^ b[k].a, a[k].b ^ b[k].b};
                        if (G[k].w == 9) o = {~o.a, ~o.b};
                    }
                    stw(mem0, G[k].z, rw, c, o);
                }
            g = e;
        }
    }
}

// All gates of the list are NONLINEAR and mutually independent (the nonlinear gates of one
// AND-depth level, whose linear producers have already run): gate x column parallel, gate g uses
// the zero-share block range of nonlinear gate and0 + g.  Fills the machine even when the number
// of instances is small compared with the number of gates.
template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? 512 : 256, 1) k_bin_and_layer(const uint4* __restrict__ gates, u32 n_gates, u64* mem0, const u64* mem1, u64 rw,
                                                                       const __grid_constant__ AesKey kp, const __grid_constant__ AesKey kn, u64 and0) {
    aes_tables_init<WIDE>();
    __syncthreads();
    const u32 Tl = (threadIdx.x & 31) * 4;
    const u64 chunks = rw / 2;
    struct Ops { W2 a0, a1, b0, b1; };
    auto load_ops = [&](const uint4 G, size_t c) {
        Ops o;
        o.a0 = ldw(mem0, G.x, rw, c); o.a1 = ldw(mem1, G.x, rw, c);
        o.b0 = ldw(mem0, G.y, rw, c); o.b1 = ldw(mem1, G.y, rw, c);
        return o;
    };
    auto gate_ops = [&](const uint4 G, u32 g, size_t c, const Ops& in, AesStream<WIDE>& sp, AesStream<WIDE>& sn) {
        const u32 type = G.w;
        const W2 a0 = in.a0, a1 = in.a1, b0 = in.b0, b1 = in.b1;
        W2 x0 = a0, x1 = a1, y0 = b0, y1 = b1;
        if (type == 1) { x0 = {~a0.a, ~a0.b}; x1 = {~a1.a, ~a1.b}; y0 = {~b0.a, ~b0.b}; y1 = {~b1.a, ~b1.b}; }
        else if (type == 4) { x0 = {~a0.a, ~a0.b}; x1 = {~a1.a, ~a1.b}; }
        W2 o0;
        o0.a = (x0.a & y0.a) ^ (x0.a & y1.a) ^ (x1.a & y0.a);
        o0.b = (x0.b & y0.b) ^ (x0.b & y1.b) ^ (x1.b & y0.b);
        if (type == 14) { o0.a ^= a0.a ^ b0.a; o0.b ^= a0.b ^ b0.b; }
        u32 p[4], q[4];
        const u64 ctr = (and0 + g) * chunks + c;
        sp.block(Tl, kp, ctr, p);
        sn.block(Tl, kn, ctr, q);
        o0.a ^= (((u64)(p[1] ^ q[1])) << 32) | (u64)(p[0] ^ q[0]);
        o0.b ^= (((u64)(p[3] ^ q[3])) << 32) | (u64)(p[2] ^ q[2]);
        stw(mem0, G.z, rw, c, o0);
    };
    auto gate_chunk = [&](const uint4 G, u32 g, size_t c, AesStream<WIDE>& sp, AesStream<WIDE>& sn) { gate_ops(G, g, c, load_ops(G, c), sp, sn); };
    if (WIDE) {
        // ONE wave of one CTA per SM over the flat list of (gate, run of 256 chunks): a 2-D grid of gx x ceil(SMs / gx) CTAs
        // is a few CTAs more than there are SMs, and with one 128 KiB CTA per SM those few ran as a second wave (the
        // 26-gate levels of the comparison circuit took 0.122 ms instead of 0.06)
        const u64 rpg = (chunks + 255) / 256, total = rpg * n_gates;
        const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
        const u32 lane = threadIdx.x & 31;
        AesStream<WIDE> sp, sn;
        for (u64 r = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total; r += warps) {
            const u32 g = (u32)(r / rpg);
            const u64 run = r - (u64)g * rpg;
            const uint4 G = gates[g];
            // the operands of step t + 1 are in flight while step t draws its two keystream blocks (4 warps per scheduler
            // do not hide a DRAM round trip by themselves: long-scoreboard was the top stall)
            const u64 cbase = run * 256 + lane;
            Ops cur = {};
            if (cbase < chunks) cur = load_ops(G, cbase);
#pragma unroll 1
            for (int t = 0; t < 8; ++t) {
                const u64 c = cbase + t * 32;
                Ops nxt = {};
                if (t + 1 < 8 && c + 32 < chunks) nxt = load_ops(G, c + 32);
                if (c < chunks) gate_ops(G, g, c, cur, sp, sn);
                cur = nxt;
            }
        }
    } else {
        for (u32 g = blockIdx.y; g < n_gates; g += gridDim.y) {
            const uint4 G = gates[g];
            AesStream<WIDE> sp, sn;
            aes_for_each<WIDE>(chunks, [&](size_t c) { gate_chunk(G, g, c, sp, sn); });
        }
    }
}

// ---------------------------------------------------------------------------------
// One-level BITWISE circuit (int_int_bitwiseAnd / bitwiseOr: gate g = bit g of every instance) evaluated on the
// ROW-MAJOR share words, without moving the operands into the bit-sliced wire memory and back (SURVEY section 7,
// "Transposes around shallow circuits").  The zero share of the reference is defined per (gate, instance column)
// -- gate g draws the keystream blocks [(and0 + g) * chunks, +chunks), bit j of that range masks instance j
// (Sh3BinaryEvaluator.cpp:1406-1442) -- so only z is produced bit-sliced and transposed, in registers:
// a warp owns a tile of 128 instances; lane l encrypts block `tile` of gates l and l + 32 under both keys (4 AES
// blocks: 2 x 128 bits of z), eight 32 x 32 bit transposes across the warp turn them into the 64-bit z word of
// instances 32q + l (q = 0..3), and the AND / OR formula runs on the row-major words.  Same shares, bit for bit,
// as k_bin_and_layer on the transposed operands.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ u32 warp_transpose32(u32 x, u32 lane) {
    // lane l holds row l; on return lane l holds column l (bit r = row r's bit l), LSB first
    u32 m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const u32 y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
        const bool lower = (lane & j) == 0;
        const u32 sh = lower ? (y << j) : (y >> j);
        const u32 mm = lower ? m : ~m;
        x = (x & mm) | (sh & ~mm);
        m ^= m << (j >> 1);
    }
    return x;
}

template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? 512 : 256, 1) k_bitwise_rowmajor(const u64* __restrict__ a0, const u64* __restrict__ a1,
                                                          const u64* __restrict__ b0, const u64* __restrict__ b1,
                                                          u64* __restrict__ out0, u64* __restrict__ out_copy, u64 n, u32 bits, u64 chunks,
                                                          const __grid_constant__ AesKey kp, const __grid_constant__ AesKey kn,
                                                          u64 and0, u32 type) {
    aes_tables_init<WIDE>();
    __syncthreads();
    const u32 lane = threadIdx.x & 31, Tl = lane * 4;
    const u64 tiles = (n + 127) / 128;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u64 keep = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
    // a warp takes runs of kRun CONSECUTIVE tiles: lane l's counters (and0 + l) * chunks + t are then consecutive too,
    // which is what the wide AES form folds its first two rounds over
    constexpr u64 kRun = WIDE ? 32 : 1;
    AesStream<WIDE> lp, ln, hp, hn;
    for (u64 run = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); run * kRun < tiles; run += warps) {
#pragma unroll 1
        for (u64 tt = 0; tt < kRun; ++tt) {
            const u64 t = run * kRun + tt;
            if (t >= tiles) break;
            // the four instances' operand words first: their DRAM latency hides under ~2000 keystream instructions (they
            // used to be loaded after the transposes and waited for -- long-scoreboard was the top stall, issue 54 % active)
            u64 x0[4], x1[4], y0[4], y1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const u64 j = t * 128 + 32 * q + lane;
                if (j < n) { x0[q] = __ldcs(a0 + j); x1[q] = __ldcs(a1 + j); y0[q] = __ldcs(b0 + j); y1[q] = __ldcs(b1 + j); }
                else { x0[q] = x1[q] = y0[q] = y1[q] = 0; }
            }
            u32 zl[4] = {0, 0, 0, 0}, zh[4] = {0, 0, 0, 0};
            if (lane < bits) {
                u32 p[4], q[4];
                const u64 ctr = (and0 + lane) * chunks + t;
                lp.block(Tl, kp, ctr, p);
                ln.block(Tl, kn, ctr, q);
#pragma unroll
                for (int i = 0; i < 4; ++i) zl[i] = p[i] ^ q[i];
            }
            if (lane + 32 < bits) {
                u32 p[4], q[4];
                const u64 ctr = (and0 + lane + 32) * chunks + t;
                hp.block(Tl, kp, ctr, p);
                hn.block(Tl, kn, ctr, q);
#pragma unroll
                for (int i = 0; i < 4; ++i) zh[i] = p[i] ^ q[i];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const u32 lo = warp_transpose32(zl[q], lane);
                const u32 hi = bits > 32 ? warp_transpose32(zh[q], lane) : 0u;
                const u64 j = t * 128 + 32 * q + lane;
                if (j < n) {
                    u64 o = (x0[q] & y0[q]) ^ (x0[q] & y1[q]) ^ (x1[q] & y0[q]);
                    if (type == 14) o ^= x0[q] ^ y0[q];                             // Or (:912-981)
                    o = (o ^ (((u64)hi << 32) | lo)) & keep;
                    out0[j] = o;
                    if (out_copy) out_copy[j] = o;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// aby3-Basic's compare-exchange selection (BoolBasic.cpp:275-312) after its comparison: with m = the comparison bit
// widened to a 0 / -1 mask (share by share), the reference evaluates int_int_bitwiseAnd(64) TWICE over the stacked
// 2n-row operands  [m; m] & [A; B]  and  NOT[m; m] & [A; B]  (two engine runs with their own zero-share keys, a
// reshare each) and xors halves of the two results:  min = t1[0:n] ^ t2[n:2n],  max = t1[n:2n] ^ t2[0:n].
// Here both evaluations and the xors are ONE pass over the row-major words per party: lane l draws the keystream blocks
// of gates l and l + 32 of BOTH evaluations for tile t of the 2n-wide instance range (8 AES blocks), the 32 x 32
// transposes turn them into the z words of instances 128 t + 32 q + l, and the instance's two products go to the min /
// max word of element (j mod n) by xor-reduction (element i receives one contribution from instance i and one from
// instance n + i, which live in different tiles when n is not a multiple of 128).  min0 / max0 must be zero on entry.
// Plane 0 of the results only; plane 1 is the previous party's plane 0 (xor is linear), i.e. ONE reshare instead of two.
// Same share words as the two bitwise_rowmajor runs + share_op xors.
// ---------------------------------------------------------------------------------
template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? 512 : 256, WIDE ? 1 : 2) k_maxmin_rowmajor(const u64* __restrict__ c0, const u64* __restrict__ c1,
                                                            const u64* __restrict__ A0, const u64* __restrict__ A1,
                                                            const u64* __restrict__ B0, const u64* __restrict__ B1,
                                                            u64* min0, u64* max0, u64 n, u64 chunks,
                                                            const __grid_constant__ AesKey kp1, const __grid_constant__ AesKey kn1,
                                                            const __grid_constant__ AesKey kp2, const __grid_constant__ AesKey kn2,
                                                            u32 not_plane) {
    // WIDE: the four-table AES (no rotates: 240 instead of 312 alu operations per block; the kernel is alu-bound, ncu: alu
    // 79 %, lsu 70 %) without the run constants -- eight keystreams per lane leave no registers for eight AesRun states
    aes_tables_init<WIDE>();
    __syncthreads();
    const u32 lane = threadIdx.x & 31, Tl = lane * 4;
    auto enc = [&](const AesKey& k, u64 ctr, u32* o) {
        if (WIDE) aes_wide_encrypt_plain(Tl, k, ctr, o); else aes_encrypt_ctr(Tl, k, ctr, o);
    };
    const u64 n2 = 2 * n, tiles = (n2 + 127) / 128;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 t = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5); t < tiles; t += warps) {
        u64 m0[4], m1[4], y0[4], y1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u64 j = t * 128 + 32 * q + lane;
            if (j < n2) {
                const bool first = j < n;
                const u64 i = first ? j : j - n;
                m0[q] = 0 - __ldg(c0 + i); m1[q] = 0 - __ldg(c1 + i);          // a share of the bit is 0 or 1: its mask share 0 or -1 (:279-284)
                y0[q] = __ldcs((first ? A0 : B0) + i); y1[q] = __ldcs((first ? A1 : B1) + i);
            } else { m0[q] = m1[q] = y0[q] = y1[q] = 0; }
        }
        u32 z1l[4], z1h[4], z2l[4], z2h[4];
        {
            u32 p[4], q[4];
            const u64 cl = (u64)lane * chunks + t, ch = (u64)(lane + 32) * chunks + t;
            enc(kp1, cl, p); enc(kn1, cl, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) z1l[i] = p[i] ^ q[i];
            enc(kp1, ch, p); enc(kn1, ch, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) z1h[i] = p[i] ^ q[i];
            enc(kp2, cl, p); enc(kn2, cl, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) z2l[i] = p[i] ^ q[i];
            enc(kp2, ch, p); enc(kn2, ch, q);
#pragma unroll
            for (int i = 0; i < 4; ++i) z2h[i] = p[i] ^ q[i];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u64 z1 = ((u64)warp_transpose32(z1h[q], lane) << 32) | warp_transpose32(z1l[q], lane);
            const u64 z2 = ((u64)warp_transpose32(z2h[q], lane) << 32) | warp_transpose32(z2l[q], lane);
            const u64 j = t * 128 + 32 * q + lane;
            if (j < n2) {
                const u64 nx0 = not_plane == 1 ? ~m0[q] : m0[q], nx1 = not_plane == 2 ? ~m1[q] : m1[q];   // the complemented share x_1 (:315-343)
                const u64 t1 = (m0[q] & y0[q]) ^ (m0[q] & y1[q]) ^ (m1[q] & y0[q]) ^ z1;
                const u64 t2 = (nx0 & y0[q]) ^ (nx0 & y1[q]) ^ (nx1 & y0[q]) ^ z2;
                const bool first = j < n;
                const u64 i = first ? j : j - n;
                atomicXor(reinterpret_cast<unsigned long long*>(min0 + i), (unsigned long long)(first ? t1 : t2));
                atomicXor(reinterpret_cast<unsigned long long*>(max0 + i), (unsigned long long)(first ? t2 : t1));
            }
        }
    }
}

// rows <-> contiguous message; VEC = bytes moved per thread step (16, 8 or 1).  PACK may
// complement the rows flagged in `invert` (inverted output wires, getOutput :1252-1258).
template <int VEC, bool PACK>
__global__ void __launch_bounds__(256) k_bin_rows(u8* mem, u64 row_bytes, const u32* __restrict__ locs, u32 n_locs, u64 nbytes, u8* buf,
                                                  const u8* __restrict__ invert) {
    // blockIdx.y walks the rows, blockIdx.x / threadIdx.x the row's VEC-byte pieces (no 64-bit divisions)
    const u64 per = nbytes / VEC;
    for (u32 j = blockIdx.y; j < n_locs; j += gridDim.y) {
        u8* mrow = mem + (u64)locs[j] * row_bytes;
        u8* brow = buf + (u64)j * nbytes;
        const bool inv = PACK && invert && invert[j];
        for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < per; c += (u64)gridDim.x * blockDim.x) {
            u8* m = mrow + c * VEC;
            u8* b = brow + c * VEC;
            if (VEC == 16) {
                if (PACK) {
                    uint4 v = *reinterpret_cast<const uint4*>(m);
                    if (inv) { v.x = ~v.x; v.y = ~v.y; v.z = ~v.z; v.w = ~v.w; }
                    *reinterpret_cast<uint4*>(b) = v;
                } else *reinterpret_cast<uint4*>(m) = *reinterpret_cast<const uint4*>(b);
            } else if (VEC == 8) {
                if (PACK) { u64 v = *reinterpret_cast<const u64*>(m); *reinterpret_cast<u64*>(b) = inv ? ~v : v; }
                else *reinterpret_cast<u64*>(m) = *reinterpret_cast<const u64*>(b);
            } else {
                if (PACK) *b = inv ? (u8)~*m : *m; else *m = *b;
            }
        }
    }
}

// One wire row -> one 64-bit word per instance holding that bit (the output of a comparison circuit: getOutput of a 1-bit
// bundle, Sh3BinaryEvaluator.cpp:1285-1404).  A thread per instance, 8-byte coalesced stores; the general 8 x 8 staged kernel
// ran this shape at 0.24 of the HBM write rate.
__global__ void __launch_bounds__(256) k_row_to_bit_words(const u8* __restrict__ mem, const u32* __restrict__ row_index, u64 row_bytes, u64 width,
                                                          u64* __restrict__ out, const u8* __restrict__ invert) {
    const u64* row = reinterpret_cast<const u64*>(mem + (u64)(row_index ? row_index[0] : 0u) * row_bytes);
    const u64 flip = (invert && invert[0]) ? 1 : 0;
    for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < width; c += (u64)gridDim.x * blockDim.x)
        __stcs(out + c, ((row[c >> 6] >> (c & 63)) & 1) ^ flip);
}

int launch_transpose(aby3cu_ctx* ctx, const void* in, const u32* row_index, u64 rows, u64 cols, u64 in_stride,
                     void* out, u64 out_stride, const u8* invert) {
    ABY3CU_REQUIRE(ctx && ((in && out) || !(rows * cols)), "bit_transpose: null argument");
    if (!(rows * cols)) return 0;
    ABY3CU_REQUIRE(in_stride % 4 == 0 && out_stride % 4 == 0, "bit_transpose: strides must be multiples of 4");
    ABY3CU_REQUIRE(((cols + 31) / 32) * 4 <= in_stride, "bit_transpose: in_stride too small for cols");
    ABY3CU_REQUIRE(((rows + 31) / 32) * 4 <= out_stride, "bit_transpose: out_stride too small for rows");
    ABY3CU_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
                   "bit_transpose: pointers must be 4-byte aligned");
    DeviceGuard g(ctx->device);
    const u64 cap = (u64)ctx->sm_count * 4;
    const auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    // fast path 1: instances x 64-bit words -> bit-sliced rows (setInput)
    if (!row_index && !invert && in_stride % 8 == 0 && in_stride <= 128 && rows > cols && cols <= in_stride * 8 &&
        out_stride % 16 == 0 && out_stride >= ((((rows + 7) / 8) + 15) & ~15ull) && al(in, 8) && al(out, 16)) {
        const u64 W = in_stride / 8;
        if (!getenv("ABY3CU_TRANSPOSE8")) {
            const u64 threads = ((rows + 31) / 32) * W, blocks = (threads + 127) / 128, capb = (u64)ctx->sm_count * 8;
            k_bits_to_sliced32<<<(unsigned)(blocks < capb ? blocks : capb), 128, 0, ctx->stream>>>((const u64*)in, rows, cols, W, (u8*)out, out_stride,
                                                                                                    W == 1 && al(in, 16));
            return post_launch(ctx, "k_bits_to_sliced32");
        }
        const u64 tiles = ((rows + kFastInst - 1) / kFastInst) * W;
        const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
        k_bits_to_sliced<<<grid, 256, 0, ctx->stream>>>((const u64*)in, rows, cols, W, (u8*)out, out_stride);
        return post_launch(ctx, "k_bits_to_sliced");
    }
    // fast path 3: ONE bit-sliced row -> one word per instance
    if (rows == 1 && out_stride == 8 && cols > 64 && in_stride % 8 == 0 && al(in, 8) && al(out, 8)) {
        const u64 blocks = (cols + 255) / 256, capb = (u64)ctx->sm_count * 8;
        k_row_to_bit_words<<<(unsigned)(blocks < capb ? blocks : capb), 256, 0, ctx->stream>>>((const u8*)in, row_index, in_stride, cols, (u64*)out, invert);
        return post_launch(ctx, "k_row_to_bit_words");
    }
    // fast path 2: bit-sliced rows -> instances x 64-bit words (getOutput)
    if (out_stride % 8 == 0 && out_stride <= 128 && cols > rows && rows <= out_stride * 8 && in_stride % 16 == 0 &&
        in_stride >= ((((cols + 7) / 8) + 15) & ~15ull) && al(in, 16) && al(out, 8)) {
        // (the 32-instance register transpose does not pay off in this direction: 64 four-byte row reads per thread ran
        // at half the rate of the shared-memory staged 8 x 8 kernel)
        const u64 W = out_stride / 8;
        const u64 tiles = ((cols + kFastInst - 1) / kFastInst) * W;
        const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
        k_sliced_to_bits<<<grid, 256, 0, ctx->stream>>>((const u8*)in, row_index, rows, cols, in_stride, (u64*)out, W, invert);
        return post_launch(ctx, "k_sliced_to_bits");
    }
#define ABY3CU_BT(TR, TCW)                                                                                        \
    do {                                                                                                          \
        const u64 tiles = ((rows + TR - 1) / TR) * ((cols + 32 * TCW - 1) / (32 * TCW));                          \
        const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);                                              \
        k_bit_transpose<TR, TCW><<<grid, 256, 0, ctx->stream>>>((const u8*)in, row_index, rows, cols, in_stride,  \
                                                                (u8*)out, out_stride, invert);                    \
    } while (0)
    if (cols <= 64) ABY3CU_BT(1024, 2);
    else if (rows <= 64) ABY3CU_BT(64, 32);
    else ABY3CU_BT(256, 8);
#undef ABY3CU_BT
    return post_launch(ctx, "k_bit_transpose");
}

}  // namespace
}  // namespace aby3cu

using namespace aby3cu;

extern "C" {

uint64_t aby3cu_bin_row_bytes(uint64_t width) { return 256 * ((width + 2047) / 2048); }

int aby3cu_bit_transpose(aby3cu_ctx* ctx, const void* d_in, u64 rows, u64 cols, u64 in_stride, void* d_out,
                         u64 out_stride, const u8* d_invert_rows) {
    return launch_transpose(ctx, d_in, nullptr, rows, cols, in_stride, d_out, out_stride, d_invert_rows);
}

int aby3cu_bit_transpose_gather(aby3cu_ctx* ctx, const void* d_in, const u32* d_row_index, u64 rows, u64 cols,
                                u64 in_stride, void* d_out, u64 out_stride, const u8* d_invert_rows) {
    ABY3CU_REQUIRE(d_row_index || !rows, "bit_transpose_gather: null row index");
    return launch_transpose(ctx, d_in, d_row_index, rows, cols, in_stride, d_out, out_stride, d_invert_rows);
}

int aby3cu_bin_level(aby3cu_ctx* ctx, const u32* d_gates, u32 n_gates, void* d_mem0, void* d_mem1, u64 row_bytes,
                     const u8 key_prev[16], const u8 key_next[16], u64 and_index0) {
    ABY3CU_REQUIRE(ctx && ((d_gates && d_mem0 && (d_mem1 || !key_prev)) || !n_gates), "bin_level: null argument");
    ABY3CU_REQUIRE(row_bytes % 16 == 0, "bin_level: row_bytes must be a multiple of 16");
    ABY3CU_REQUIRE((key_prev == nullptr) == (key_next == nullptr), "bin_level: give both keys or neither");
    if (!n_gates || !row_bytes) return 0;
    DeviceGuard g(ctx->device);
    const u64 rw = row_bytes / 8, chunks = rw / 2;
    if (key_prev) {
        AesKey kp, kn; host_expand_key(key_prev, &kp); host_expand_key(key_next, &kn);
        ABY3CU_CHECK(cudaFuncSetAttribute(k_bin_level<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesTableBytes));
        if (prefer_max_smem(k_bin_level<true>)) return 1;
        const unsigned grid = ew_grid(ctx, chunks, 256, 3);
        k_bin_level<true><<<grid, 256, kAesTableBytes, ctx->stream>>>((const uint4*)d_gates, n_gates, (u64*)d_mem0, (u64*)d_mem1, rw, kp, kn, and_index0);
    } else {
        static const AesKey zero = {};
        const unsigned grid = ew_grid(ctx, chunks, 256, 8);
        if (!d_mem1) k_bin_level<false, true><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_gates, n_gates, (u64*)d_mem0, nullptr, rw, zero, zero, and_index0);
        else k_bin_level<false><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_gates, n_gates, (u64*)d_mem0, (u64*)d_mem1, rw, zero, zero, and_index0);
    }
    return post_launch(ctx, "k_bin_level");
}

int aby3cu_bin_linear_plane0(aby3cu_ctx* ctx, const u32* d_gates, u32 n_gates, const u8* d_batch_first, void* d_mem0, u64 row_bytes) {
    ABY3CU_REQUIRE(ctx && ((d_gates && d_batch_first && d_mem0) || !n_gates), "bin_linear_plane0: null argument");
    ABY3CU_REQUIRE(row_bytes % 16 == 0, "bin_linear_plane0: row_bytes must be a multiple of 16");
    if (!n_gates || !row_bytes) return 0;
    DeviceGuard g(ctx->device);
    const u64 rw = row_bytes / 8, chunks = rw / 2;
    k_bin_linear_plane0<<<ew_grid(ctx, chunks, 256, 8), 256, 0, ctx->stream>>>((const uint4*)d_gates, d_batch_first, n_gates, (u64*)d_mem0, rw);
    return post_launch(ctx, "k_bin_linear_plane0");
}

int aby3cu_bin_and_layer(aby3cu_ctx* ctx, const u32* d_gates, u32 n_gates, void* d_mem0, const void* d_mem1, u64 row_bytes,
                         const u8 key_prev[16], const u8 key_next[16], u64 and_index0) {
    ABY3CU_REQUIRE(ctx && key_prev && key_next && ((d_gates && d_mem0 && d_mem1) || !n_gates), "bin_and_layer: null argument");
    ABY3CU_REQUIRE(row_bytes % 16 == 0, "bin_and_layer: row_bytes must be a multiple of 16");
    if (!n_gates || !row_bytes) return 0;
    DeviceGuard g(ctx->device);
    AesKey kp, kn; host_expand_key(key_prev, &kp); host_expand_key(key_next, &kn);
    const u64 chunks = row_bytes / 16;
    static const bool wide_on = [] { const char* e = getenv("ABY3CU_AES_WIDE"); return !(e && e[0] == '0'); }();
    if (wide_on && !ctx->corun && chunks * n_gates >= (1u << 16) && chunks >= 2048) {
        ABY3CU_CHECK(cudaFuncSetAttribute(k_bin_and_layer<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesWideTableBytes));
        if (prefer_max_smem(k_bin_and_layer<true>)) return 1;
        // one 512-thread CTA per SM, one wave: the warps walk the flat list of (gate, run of 256 chunks)
        u64 gx = (((chunks + 255) / 256) * n_gates + 15) / 16;
        if (gx > (u64)ctx->sm_count) gx = ctx->sm_count;
        k_bin_and_layer<true><<<(unsigned)gx, 512, kAesWideTableBytes, ctx->stream>>>((const uint4*)d_gates, n_gates, (u64*)d_mem0,
                                                                                                          (const u64*)d_mem1, row_bytes / 8, kp, kn, and_index0);
        return post_launch(ctx, "k_bin_and_layer");
    }
    ABY3CU_CHECK(cudaFuncSetAttribute(k_bin_and_layer<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesTableBytes));
    if (prefer_max_smem(k_bin_and_layer<false>)) return 1;
    const u64 gx = (chunks + 255) / 256 < 32 ? (chunks + 255) / 256 : 32;
    u64 gy = ((u64)ctx->sm_count * 3 + gx - 1) / gx;
    if (gy > n_gates) gy = n_gates;
    k_bin_and_layer<false><<<dim3((unsigned)gx, (unsigned)gy), 256, kAesTableBytes, ctx->stream>>>((const uint4*)d_gates, n_gates, (u64*)d_mem0,
                                                                                                   (const u64*)d_mem1, row_bytes / 8, kp, kn, and_index0);
    return post_launch(ctx, "k_bin_and_layer");
}

int aby3cu_bin_bitwise_rowmajor(aby3cu_ctx* ctx, uint32_t gate_type, const int64_t* d_a0, const int64_t* d_a1, const int64_t* d_b0,
                                const int64_t* d_b1, int64_t* d_out0, int64_t* d_out_copy, uint64_t n, uint32_t bits, uint64_t row_bytes,
                                const u8 key_prev[16], const u8 key_next[16], uint64_t and_index0) {
    ABY3CU_REQUIRE(ctx && key_prev && key_next && ((d_a0 && d_a1 && d_b0 && d_b1 && d_out0) || !n), "bin_bitwise_rowmajor: null argument");
    ABY3CU_REQUIRE(gate_type == 8 || gate_type == 14, "bin_bitwise_rowmajor: gate type must be And (8) or Or (14)");
    ABY3CU_REQUIRE(bits >= 1 && bits <= 64, "bin_bitwise_rowmajor: 1..64 bits per instance");
    ABY3CU_REQUIRE(row_bytes % 16 == 0 && row_bytes * 8 >= n, "bin_bitwise_rowmajor: row_bytes must be the wire-row size of the circuit (a multiple of 16 covering n bits)");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey kp, kn; host_expand_key(key_prev, &kp); host_expand_key(key_next, &kn);
    const u64 tiles = (n + 127) / 128;
    static const bool wide_on = [] { const char* e = getenv("ABY3CU_AES_WIDE"); return !(e && e[0] == '0'); }();
    // (measured: with four keystreams per lane and the transposes the wide form is no faster here -- 0.30 vs 0.28 ms at
    // 2^23 instances -- so it stays behind an opt-in switch)
    static const bool wide_bitwise = [] { const char* e = getenv("ABY3CU_AES_WIDE_BITWISE"); return e && e[0] == '1'; }();
    if (wide_on && wide_bitwise && !ctx->corun && tiles >= 4096) {
        ABY3CU_CHECK(cudaFuncSetAttribute(k_bitwise_rowmajor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesWideTableBytes));
        if (prefer_max_smem(k_bitwise_rowmajor<true>)) return 1;
        k_bitwise_rowmajor<true><<<ctx->sm_count, 512, kAesWideTableBytes, ctx->stream>>>((const u64*)d_a0, (const u64*)d_a1, (const u64*)d_b0, (const u64*)d_b1,
                                                                                         (u64*)d_out0, (u64*)d_out_copy, n, bits, row_bytes / 16, kp, kn, and_index0, gate_type);
        return post_launch(ctx, "k_bitwise_rowmajor");
    }
    ABY3CU_CHECK(cudaFuncSetAttribute(k_bitwise_rowmajor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesTableBytes));
    if (prefer_max_smem(k_bitwise_rowmajor<false>)) return 1;
    // 128 registers x 256 threads: two CTAs per SM are resident; a grid of three per SM ran as 1.5 waves (ncu), the second one half empty
    const unsigned grid = ew_grid(ctx, tiles * 32, 256, 2);
    k_bitwise_rowmajor<false><<<grid, 256, kAesTableBytes, ctx->stream>>>((const u64*)d_a0, (const u64*)d_a1, (const u64*)d_b0, (const u64*)d_b1,
                                                                          (u64*)d_out0, (u64*)d_out_copy, n, bits, row_bytes / 16, kp, kn, and_index0, gate_type);
    return post_launch(ctx, "k_bitwise_rowmajor");
}

int aby3cu_bin_maxmin_rowmajor(aby3cu_ctx* ctx, const int64_t* d_c0, const int64_t* d_c1, const int64_t* d_a0, const int64_t* d_a1,
                               const int64_t* d_b0, const int64_t* d_b1, int64_t* d_min0, int64_t* d_max0, uint64_t n, uint64_t row_bytes,
                               const u8 key_prev1[16], const u8 key_next1[16], const u8 key_prev2[16], const u8 key_next2[16], uint32_t not_plane) {
    ABY3CU_REQUIRE(ctx && key_prev1 && key_next1 && key_prev2 && key_next2 && ((d_c0 && d_c1 && d_a0 && d_a1 && d_b0 && d_b1 && d_min0 && d_max0) || !n),
                   "bin_maxmin_rowmajor: null argument");
    ABY3CU_REQUIRE(not_plane <= 2, "bin_maxmin_rowmajor: not_plane is 0 (none), 1 (plane 0) or 2 (plane 1)");
    ABY3CU_REQUIRE(row_bytes % 16 == 0 && row_bytes * 8 >= 2 * n, "bin_maxmin_rowmajor: row_bytes must be the wire-row size of a circuit over 2 n instances");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey kp1, kn1, kp2, kn2;
    host_expand_key(key_prev1, &kp1); host_expand_key(key_next1, &kn1); host_expand_key(key_prev2, &kp2); host_expand_key(key_next2, &kn2);
    ABY3CU_CHECK(cudaMemsetAsync(d_min0, 0, n * 8, ctx->stream));
    ABY3CU_CHECK(cudaMemsetAsync(d_max0, 0, n * 8, ctx->stream));
    const u64 tiles = (2 * n + 127) / 128;
    static const bool wide_on = [] { const char* e = getenv("ABY3CU_AES_WIDE"); return !(e && e[0] == '0'); }();
    if (wide_on && !ctx->corun && tiles >= 4096) {
        ABY3CU_CHECK(cudaFuncSetAttribute(k_maxmin_rowmajor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesWideTableBytes));
        if (prefer_max_smem(k_maxmin_rowmajor<true>)) return 1;
        k_maxmin_rowmajor<true><<<ew_grid(ctx, tiles * 32, 512, 1), 512, kAesWideTableBytes, ctx->stream>>>(
            (const u64*)d_c0, (const u64*)d_c1, (const u64*)d_a0, (const u64*)d_a1, (const u64*)d_b0, (const u64*)d_b1, (u64*)d_min0, (u64*)d_max0, n,
            row_bytes / 16, kp1, kn1, kp2, kn2, not_plane);
        return post_launch(ctx, "k_maxmin_rowmajor");
    }
    ABY3CU_CHECK(cudaFuncSetAttribute(k_maxmin_rowmajor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAesTableBytes));
    if (prefer_max_smem(k_maxmin_rowmajor<false>)) return 1;
    const unsigned grid = ew_grid(ctx, tiles * 32, 256, 2);
    k_maxmin_rowmajor<false><<<grid, 256, kAesTableBytes, ctx->stream>>>((const u64*)d_c0, (const u64*)d_c1, (const u64*)d_a0, (const u64*)d_a1, (const u64*)d_b0,
                                                                  (const u64*)d_b1, (u64*)d_min0, (u64*)d_max0, n, row_bytes / 16, kp1, kn1, kp2, kn2, not_plane);
    return post_launch(ctx, "k_maxmin_rowmajor");
}

// The shadow evaluator of the reference's BINARY_ENGINE_DEBUG build (Sh3BinaryEvaluator.cpp:1469-1601), on the device:
// with all three share planes of the wire memory at hand, every gate is re-evaluated on the reconstructed wire values
// and compared with the reconstructed output wire; mismatching instances are counted.
__global__ void __launch_bounds__(256) k_bin_check(const uint4* __restrict__ gates, const u8* __restrict__ skip, u32 n_gates,
                                                   const u64* __restrict__ m0, const u64* __restrict__ m1, const u64* __restrict__ m2,
                                                   u64 rw, u64 width, unsigned long long* __restrict__ bad, u32* __restrict__ first_bad) {
    const u64 words = (width + 63) / 64;
    for (u32 g = blockIdx.y; g < n_gates; g += gridDim.y) {
        if (skip && skip[g]) continue;
        const uint4 G = gates[g];
        for (u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x; c < words; c += (u64)gridDim.x * blockDim.x) {
            const u64 a = m0[(u64)G.x * rw + c] ^ m1[(u64)G.x * rw + c] ^ m2[(u64)G.x * rw + c];
            const u64 b = m0[(u64)G.y * rw + c] ^ m1[(u64)G.y * rw + c] ^ m2[(u64)G.y * rw + c];
            const u64 o = m0[(u64)G.z * rw + c] ^ m1[(u64)G.z * rw + c] ^ m2[(u64)G.z * rw + c];
            u64 e;
            switch (G.w) {
            case 6: e = a ^ b; break;
            case 8: e = a & b; break;
            case 1: e = ~(a | b); break;
            case 14: e = a | b; break;
            case 9: e = ~(a ^ b); break;
            case 10: e = a; break;
            case 4: e = ~a & b; break;
            default: e = ~o; break;                 // unsupported type: always a mismatch
            }
            u64 diff = e ^ o;
            if (c == words - 1 && (width & 63)) diff &= (1ull << (width & 63)) - 1;
            if (diff) {
                atomicAdd(bad, (unsigned long long)__popcll(diff));
                atomicMin(first_bad, g);
            }
        }
    }
}

static int rows_copy(aby3cu_ctx* ctx, bool pack, void* mem, u64 row_bytes, const u32* locs, u32 n_locs, u64 nbytes, void* buf,
                     const u8* invert) {
    ABY3CU_REQUIRE(ctx && ((mem && locs && buf) || !(n_locs * nbytes)), "bin rows: null argument");
    ABY3CU_REQUIRE(nbytes <= row_bytes, "bin rows: nbytes exceeds the row");
    if (!(n_locs * nbytes)) return 0;
    DeviceGuard g(ctx->device);
    const uintptr_t bits = reinterpret_cast<uintptr_t>(mem) | reinterpret_cast<uintptr_t>(buf) | nbytes | row_bytes;
    const int vec = (bits & 15) == 0 ? 16 : (bits & 7) == 0 ? 8 : 1;
    const u64 per = nbytes / vec;
    const unsigned gx = (unsigned)((per + 255) / 256 < 64 ? (per + 255) / 256 : 64);
    const u64 want_y = ((u64)ctx->sm_count * 8 + gx - 1) / gx;
    const dim3 grid(gx ? gx : 1, (unsigned)(n_locs < want_y ? n_locs : (want_y ? want_y : 1)));
#define ABY3CU_ROWS(V, P) k_bin_rows<V, P><<<grid, 256, 0, ctx->stream>>>((u8*)mem, row_bytes, locs, n_locs, nbytes, (u8*)buf, invert)
    if (vec == 16) { if (pack) ABY3CU_ROWS(16, true); else ABY3CU_ROWS(16, false); }
    else if (vec == 8) { if (pack) ABY3CU_ROWS(8, true); else ABY3CU_ROWS(8, false); }
    else { if (pack) ABY3CU_ROWS(1, true); else ABY3CU_ROWS(1, false); }
#undef ABY3CU_ROWS
    return post_launch(ctx, "k_bin_rows");
}

int aby3cu_bin_pack_rows(aby3cu_ctx* ctx, const void* d_mem, u64 row_bytes, const u32* d_locs, u32 n_locs, u64 nbytes, void* d_out,
                         const u8* d_invert) {
    return rows_copy(ctx, true, const_cast<void*>(d_mem), row_bytes, d_locs, n_locs, nbytes, d_out, d_invert);
}

int aby3cu_bin_scatter_rows(aby3cu_ctx* ctx, void* d_mem, u64 row_bytes, const u32* d_locs, u32 n_locs, u64 nbytes, const void* d_in) {
    return rows_copy(ctx, false, d_mem, row_bytes, d_locs, n_locs, nbytes, const_cast<void*>(d_in), nullptr);
}

int aby3cu_bin_check_gates(aby3cu_ctx* ctx, const u32* d_gates, const u8* d_skip, u32 n_gates, const void* d_plane_a, const void* d_plane_b,
                           const void* d_plane_c, u64 row_bytes, u64 width, u64* d_bad_count, u32* d_first_bad_gate) {
    ABY3CU_REQUIRE(ctx && d_bad_count && d_first_bad_gate, "bin_check_gates: null argument");
    ABY3CU_REQUIRE(row_bytes % 8 == 0, "bin_check_gates: row_bytes must be a multiple of 8");
    DeviceGuard g(ctx->device);
    ABY3CU_CHECK(cudaMemsetAsync(d_bad_count, 0, 8, ctx->stream));
    ABY3CU_CHECK(cudaMemsetAsync(d_first_bad_gate, 0xFF, 4, ctx->stream));
    if (!n_gates || !width) return 0;
    ABY3CU_REQUIRE(d_gates && d_plane_a && d_plane_b && d_plane_c, "bin_check_gates: null argument");
    const u64 words = (width + 63) / 64;
    const unsigned gx = (unsigned)((words + 255) / 256 < 64 ? (words + 255) / 256 : 64);
    u64 gy = ((u64)ctx->sm_count * 8 + gx - 1) / gx;
    if (gy > n_gates) gy = n_gates;
    k_bin_check<<<dim3(gx, (unsigned)gy), 256, 0, ctx->stream>>>((const uint4*)d_gates, d_skip, n_gates, (const u64*)d_plane_a, (const u64*)d_plane_b,
                                                                   (const u64*)d_plane_c, row_bytes / 8, width, (unsigned long long*)d_bad_count,
                                                                   d_first_bad_gate);
    return post_launch(ctx, "k_bin_check");
}

}  // extern "C"
