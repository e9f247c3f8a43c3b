// elementwise.cu -- the HBM-bound kernels of the arithmetic path: AES-CTR fill,
// zero sharing, Hadamard multiplication (with and without truncation pair),
// truncation tuple / finish, share add/sub/xor, reveal combine, transpose,
// row gather.  Every thread moves 16 bytes per array per step (LDG.128/STG.128),
// grids are a multiple of the SM count, keystreams are produced in registers so
// no random bytes ever touch HBM in the fused kernels.
#include "aes.cuh"

namespace aby3cu {

namespace {

constexpr int kThreads = 256;
constexpr int kCtasPerSm = 3;     // 3 x 64 KiB AES tables fit in 227 KiB of shared memory
constexpr int kWideThreads = 512; // the wide AES form: one CTA per SM (128 KiB of tables)
constexpr size_t kWideMinPairs = size_t(1) << 16;   // below this the 128 KiB table fill is not amortised

__device__ __forceinline__ u64 sar(u64 x, unsigned s) { return (u64)((i64)x >> s); }

struct alignas(16) U64x2 { u64 a, b; };

__device__ __forceinline__ U64x2 ld2(const i64* p, size_t i, size_t n, bool vec) {
    U64x2 v;
    if (vec) {
        v = *reinterpret_cast<const U64x2*>(p + i);
    } else {
        v.a = (u64)p[i];
        v.b = (i + 1 < n) ? (u64)p[i + 1] : 0;
    }
    return v;
}
__device__ __forceinline__ void st2(i64* p, size_t i, size_t n, bool vec, U64x2 v) {
    if (vec) {
        *reinterpret_cast<U64x2*>(p + i) = v;
    } else {
        p[i] = (i64)v.a;
        if (i + 1 < n) p[i + 1] = (i64)v.b;
    }
}

__host__ __device__ inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------
// keystream fill: out[i] = KS[e0 + i]
// ---------------------------------------------------------------------------------
template <bool WIDE>
__global__ void __launch_bounds__(WIDE ? kWideThreads : kThreads, 1) k_fill(const __grid_constant__ AesKey key, u64 e0, i64* __restrict__ out, size_t n, int vec) {
    aes_tables_init<WIDE>();
    __syncthreads();
    const u32 Tl = (threadIdx.x & 31) * 4;
    AesStream<WIDE> ks;
    aes_for_each<WIDE>((n + 1) / 2, [&](size_t p) {
        U64x2 v;
        ks.pair(Tl, key, e0 + 2 * p, v.a, v.b);
        const size_t i = 2 * p;
        st2(out, i, n, vec && i + 1 < n, v);
    });
}

// ---------------------------------------------------------------------------------
// zero share: out[i] = addend[i] (+|^) (KSp[e0+i] (-|^) KSn[e0+i])
// ---------------------------------------------------------------------------------
template <bool BINARY, bool WIDE>
__global__ void __launch_bounds__(WIDE ? kWideThreads : kThreads, 1) k_zero_share(const __grid_constant__ AesKey kp, const __grid_constant__ AesKey kn, u64 e0,
                                                         const i64* __restrict__ addend, i64* __restrict__ out, size_t n, int vec) {
    aes_tables_init<WIDE>();
    __syncthreads();
    const u32 Tl = (threadIdx.x & 31) * 4;
    AesStream<WIDE> sp, sn;
    aes_for_each<WIDE>((n + 1) / 2, [&](size_t p) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        U64x2 a = {0, 0};
        if (addend) a = ld2(addend, i, n, v2);
        u64 p0, p1, q0, q1;
        sp.pair(Tl, kp, e0 + i, p0, p1);
        sn.pair(Tl, kn, e0 + i, q0, q1);
        U64x2 r;
        if (BINARY) { r.a = a.a ^ p0 ^ q0; r.b = a.b ^ p1 ^ q1; }
        else        { r.a = a.a + (p0 - q0); r.b = a.b + (p1 - q1); }
        st2(out, i, n, v2, r);
    });
}

// ---------------------------------------------------------------------------------
// Hadamard multiplication, non-truncating: C0 = A0*B0 + A0*B1 + A1*B0 + z
// ---------------------------------------------------------------------------------
__device__ __forceinline__ u64 cross(u64 a0, u64 a1, u64 b0, u64 b1) { return a0 * (b0 + b1) + a1 * b0; }

template <bool MASK, bool WIDE>
__global__ void __launch_bounds__(WIDE ? kWideThreads : kThreads, 1) k_mul_hadamard(const i64* __restrict__ A0, const i64* __restrict__ A1,
                                                           const i64* __restrict__ B0, const i64* __restrict__ B1,
                                                           const __grid_constant__ AesKey kp, const __grid_constant__ AesKey kn,
                                                           u64 e0, i64* __restrict__ C0, size_t n, int vec) {
    if (MASK) { aes_tables_init<WIDE>(); __syncthreads(); }
    const u32 Tl = (threadIdx.x & 31) * 4;
    AesStream<WIDE> sp, sn;
    aes_for_each<WIDE>((n + 1) / 2, [&](size_t p) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        U64x2 a0 = ld2(A0, i, n, v2), a1 = ld2(A1, i, n, v2), b0 = ld2(B0, i, n, v2), b1 = ld2(B1, i, n, v2);
        U64x2 r;
        r.a = cross(a0.a, a1.a, b0.a, b1.a);
        r.b = cross(a0.b, a1.b, b0.b, b1.b);
        if (MASK) {
            u64 p0, p1, q0, q1;
            sp.pair(Tl, kp, e0 + i, p0, p1);
            sn.pair(Tl, kn, e0 + i, q0, q1);
            r.a += p0 - q0;
            r.b += p1 - q1;
        }
        st2(C0, i, n, v2, r);
    });
}

// ---------------------------------------------------------------------------------
// Truncation tuple, optionally fused with the Hadamard cross term.
//   t0 = KS_next[en+i], t1 = KS_prev[ep+i]; r = t0>>2; RT0 = t0>>(d+2); RT1 = t1>>(d+2)
//   CROSS: V = cross - r
//   else : R = r (if R), NEGR = -r (if NEGR)
// ---------------------------------------------------------------------------------
template <bool CROSS, bool RAND, bool WIDE>
__global__ void __launch_bounds__(WIDE ? kWideThreads : kThreads, 1) k_trunc(const i64* __restrict__ A0, const i64* __restrict__ A1,
                                                    const i64* __restrict__ B0, const i64* __restrict__ B1,
                                                    const __grid_constant__ AesKey knext, u64 en,
                                                    const __grid_constant__ AesKey kprev, u64 ep, unsigned d2,
                                                    i64* __restrict__ V, i64* __restrict__ R, i64* __restrict__ NEGR,
                                                    i64* __restrict__ RT0, i64* __restrict__ RT1, size_t n, int vec,
                                                    const u64* __restrict__ iter, u64 iter_stride) {
    if (RAND) { aes_tables_init<WIDE>(); __syncthreads(); }
    // replayed as a CUDA-graph node: the stream offsets advance with a device-resident iteration counter
    if (iter) { const u64 it = *iter; en += it * iter_stride; ep += it * iter_stride; }
    const u32 Tl = (threadIdx.x & 31) * 4;
    AesStream<WIDE> sn, sp;
    aes_for_each<WIDE>((n + 1) / 2, [&](size_t p) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        u64 t00 = 0, t01 = 0, t10 = 0, t11 = 0;
        if (RAND) {
            sn.pair(Tl, knext, en + i, t00, t01);
            sp.pair(Tl, kprev, ep + i, t10, t11);
        }
        const u64 r0 = sar(t00, 2), r1 = sar(t01, 2);
        U64x2 rt0 = {sar(t00, d2), sar(t01, d2)};
        U64x2 rt1 = {sar(t10, d2), sar(t11, d2)};
        st2(RT0, i, n, v2, rt0);
        st2(RT1, i, n, v2, rt1);
        if (CROSS) {
            U64x2 a0 = ld2(A0, i, n, v2), a1 = ld2(A1, i, n, v2), b0 = ld2(B0, i, n, v2), b1 = ld2(B1, i, n, v2);
            U64x2 v = {cross(a0.a, a1.a, b0.a, b1.a) - r0, cross(a0.b, a1.b, b0.b, b1.b) - r1};
            st2(V, i, n, v2, v);
        } else {
            if (R) { U64x2 v = {r0, r1}; st2(R, i, n, v2, v); }
            if (NEGR) { U64x2 v = {0 - r0, 0 - r1}; st2(NEGR, i, n, v2, v); }
        }
    });
}

// C += (s0+s1+s2) >> shift
__global__ void __launch_bounds__(kThreads) k_trunc_finish(const i64* __restrict__ s0, const i64* __restrict__ s1, const i64* __restrict__ s2,
                                                           i64* __restrict__ C, size_t n, unsigned shift, int vec) {
    const size_t pairs = (n + 1) / 2;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        U64x2 a = ld2(s0, i, n, v2), b = ld2(s1, i, n, v2), c = ld2(s2, i, n, v2), o = ld2(C, i, n, v2);
        o.a += sar(a.a + b.a + c.a, shift);
        o.b += sar(a.b + b.b + c.b, shift);
        st2(C, i, n, v2, o);
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_share_op(const i64* __restrict__ x, const i64* __restrict__ y, i64* __restrict__ out, size_t n, int vec) {
    const size_t pairs = (n + 1) / 2;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        U64x2 a = ld2(x, i, n, v2), b = ld2(y, i, n, v2), o;
        if (OP == ABY3CU_OP_ADD) { o.a = a.a + b.a; o.b = a.b + b.b; }
        else if (OP == ABY3CU_OP_SUB) { o.a = a.a - b.a; o.b = a.b - b.b; }
        else { o.a = a.a ^ b.a; o.b = a.b ^ b.b; }
        st2(out, i, n, v2, o);
    }
}

// both share planes of a replicated sharing in one launch (blockIdx.y = plane)
struct Planes2 { const i64* x[2]; const i64* y[2]; i64* out[2]; };
template <int OP>
__global__ void __launch_bounds__(kThreads) k_share_op2(Planes2 p, size_t n, int vec) {
    const i64* __restrict__ x = p.x[blockIdx.y];
    const i64* __restrict__ y = p.y[blockIdx.y];
    i64* __restrict__ out = p.out[blockIdx.y];
    const size_t pairs = (n + 1) / 2;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 2 * q;
        const bool v2 = vec && i + 1 < n;
        U64x2 a = ld2(x, i, n, v2), b = ld2(y, i, n, v2), o;
        if (OP == ABY3CU_OP_ADD) { o.a = a.a + b.a; o.b = a.b + b.b; }
        else if (OP == ABY3CU_OP_SUB) { o.a = a.a - b.a; o.b = a.b - b.b; }
        else { o.a = a.a ^ b.a; o.b = a.b ^ b.b; }
        st2(out, i, n, v2, o);
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_combine3(const i64* __restrict__ x0, const i64* __restrict__ x1, const i64* __restrict__ x2,
                                                       i64* __restrict__ out, size_t n, int vec) {
    const size_t pairs = (n + 1) / 2;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        U64x2 a = ld2(x0, i, n, v2), b = ld2(x1, i, n, v2), c = ld2(x2, i, n, v2), o;
        if (OP == ABY3CU_OP_ADD) { o.a = a.a + b.a + c.a; o.b = a.b + b.b + c.b; }
        else { o.a = a.a ^ b.a ^ c.a; o.b = a.b ^ b.b ^ c.b; }
        st2(out, i, n, v2, o);
    }
}

// out = a * x + b (wrapping); x may be NULL (constant fill)
__global__ void __launch_bounds__(kThreads) k_axpb(u64 a, const i64* __restrict__ x, u64 b, i64* __restrict__ out, size_t n, int vec) {
    const size_t pairs = (n + 1) / 2;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (size_t)gridDim.x * blockDim.x) {
        const size_t i = 2 * p;
        const bool v2 = vec && i + 1 < n;
        U64x2 v = {0, 0};
        if (x) v = ld2(x, i, n, v2);
        U64x2 o = {a * v.a + b, a * v.b + b};
        st2(out, i, n, v2, o);
    }
}

// 32x32 tile transpose of 8-byte elements through padded shared memory
__global__ void __launch_bounds__(256) k_transpose(const i64* __restrict__ in, u64 rows, u64 cols, i64* __restrict__ out) {
    __shared__ i64 tile[32][33];
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

// 32x32 tile transpose of both share planes (blockIdx.y = plane)
__global__ void __launch_bounds__(256) k_transpose2(const i64* __restrict__ in0, const i64* __restrict__ in1, u64 rows, u64 cols,
                                                    i64* __restrict__ out0, i64* __restrict__ out1) {
    __shared__ i64 tile[32][33];
    const i64* __restrict__ in = blockIdx.y ? in1 : in0;
    i64* __restrict__ out = blockIdx.y ? out1 : out0;
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

// several row gathers that share one index vector, in one launch (blockIdx.y = job): the mini-batch
// extraction of SGD takes the same rows of X and Y, both share planes
struct GatherJobs { const i64* in[ABY3CU_MAX_GATHER_JOBS]; i64* out[ABY3CU_MAX_GATHER_JOBS]; u64 cols[ABY3CU_MAX_GATHER_JOBS]; };
__global__ void __launch_bounds__(256) k_gather_rows_multi(GatherJobs jobs, const u64* __restrict__ idx, u64 nrows, const u64* __restrict__ iter) {
    if (iter) idx += *iter * nrows;            // graph replay: batch number `*iter` of a long index vector
    const i64* __restrict__ in = jobs.in[blockIdx.y];
    i64* __restrict__ out = jobs.out[blockIdx.y];
    const u64 cols = jobs.cols[blockIdx.y];
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cols == 1) {
        for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (u64)gridDim.x * blockDim.x) out[r] = in[idx[r]];
        return;
    }
    const bool v2 = (cols & 1) == 0 && al16(in) && al16(out);
    for (u64 r = warp; r < nrows; r += nwarps) {
        const i64* src = in + idx[r] * cols;
        i64* dst = out + r * cols;
        if (v2) {
            const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(src);
            ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
            for (u64 c = lane; c < cols / 2; c += 32) d2[c] = s2[c];
        } else {
            for (u64 c = lane; c < cols; c += 32) dst[c] = src[c];
        }
    }
}

// out[r,:] = in[idx[r],:]; one warp per row, 16 B per lane per step
__global__ void __launch_bounds__(256) k_gather_rows(const i64* __restrict__ in, u64 cols, const u64* __restrict__ idx, u64 nrows, i64* __restrict__ out) {
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cols == 1) {
        for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (u64)gridDim.x * blockDim.x) out[r] = in[idx[r]];
        return;
    }
    for (u64 r = warp; r < nrows; r += nwarps) {
        const i64* src = in + idx[r] * cols;
        i64* dst = out + r * cols;
        for (u64 c = lane; c < cols; c += 32) dst[c] = src[c];
    }
}

// One compare-exchange stage of the merge network (aby3-Basic/Sort.cpp:366-393) pairs positions r0 + 2i and r0 + d + 2i
// with d ODD: the two operand vectors are the two parities of ONE contiguous range.  Pair p = (src[r0 + 2p], src[r0 + 2p + 1])
// holds X[p] and Y[p - (d - 1) / 2], so the range is read (or written) exactly once, contiguously, instead of four strided
// passes through index vectors.  PLANES share planes per launch (blockIdx.y).
struct CmpxPlanes { const i64* src[2]; i64* x[2]; i64* y[2]; };
__global__ void __launch_bounds__(256) k_cmpx_gather(CmpxPlanes P, u64 r0, u64 h /* (d - 1) / 2 */, u64 m) {
    const i64* __restrict__ src = P.src[blockIdx.y] + r0;
    i64* __restrict__ X = P.x[blockIdx.y];
    i64* __restrict__ Y = P.y[blockIdx.y];
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < m + h; p += (u64)gridDim.x * blockDim.x) {
        if (p < m) X[p] = __ldcs(src + 2 * p);
        if (p >= h) Y[p - h] = __ldcs(src + 2 * p + 1);
    }
}
struct CmpxScatter { i64* dst[2]; const i64* x[2]; const i64* y[2]; };
__global__ void __launch_bounds__(256) k_cmpx_scatter(CmpxScatter P, u64 r0, u64 h, u64 m) {
    i64* __restrict__ dst = P.dst[blockIdx.y] + r0;
    const i64* __restrict__ X = P.x[blockIdx.y];
    const i64* __restrict__ Y = P.y[blockIdx.y];
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < m + h; p += (u64)gridDim.x * blockDim.x) {
        if (p < m) dst[2 * p] = __ldcs(X + p);
        if (p >= h) dst[2 * p + 1] = __ldcs(Y + p - h);
    }
}

__global__ void __launch_bounds__(256) k_iota(u64 start, u64 step, u64* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = start + step * i;
}

// out[idx[r],:] = in[r,:]
__global__ void __launch_bounds__(256) k_scatter_rows(const i64* __restrict__ in, u64 cols, const u64* __restrict__ idx, u64 nrows, i64* __restrict__ out) {
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cols == 1) {
        for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (u64)gridDim.x * blockDim.x) out[idx[r]] = in[r];
        return;
    }
    for (u64 r = warp; r < nrows; r += nwarps) {
        const i64* src = in + r * cols;
        i64* dst = out + idx[r] * cols;
        for (u64 c = lane; c < cols; c += 32) dst[c] = src[c];
    }
}

}  // namespace

int upload_aes_constants() {
    u8 sbox[256];
    host_sbox(sbox);
    u32 te0[256];
    for (int x = 0; x < 256; ++x) {
        u8 s = sbox[x];
        u8 s2 = (u8)((s << 1) ^ ((s & 0x80) ? 0x1B : 0));
        u8 s3 = (u8)(s2 ^ s);
        te0[x] = (u32)s2 | ((u32)s << 8) | ((u32)s << 16) | ((u32)s3 << 24);
    }
    ABY3CU_CHECK(cudaMemcpyToSymbol(c_Te0, te0, sizeof(te0)));
    return 0;
}

template <class K>
static int enable_big_smem(K kernel, int bytes = kAesTableBytes) {
    ABY3CU_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (prefer_max_smem(kernel)) return 1;
    return 0;
}
// The wide AES form (four tables + CTR folding, aes.cuh) for big launches: whole blocks per pair (even first element), not
// a context whose kernels are shaped to run under a GEMM (64 KiB is what fits next to it).  ABY3CU_AES_WIDE=0 turns it off.
static bool use_wide_aes(const aby3cu_ctx* ctx, size_t pairs, u64 e0, u64 e1 = 0) {
    static const bool on = [] { const char* e = getenv("ABY3CU_AES_WIDE"); return !(e && e[0] == '0'); }();
    return on && !ctx->corun && pairs >= kWideMinPairs && ((e0 | e1) & 1) == 0;
}

}  // namespace aby3cu

using namespace aby3cu;

static const AesKey kZeroKey = {};

extern "C" {

int aby3cu_aes_ctr_fill(aby3cu_ctx* ctx, const u8 key[16], u64 byte_off, void* d_out, size_t nbytes) {
    ABY3CU_REQUIRE(ctx && key && (d_out || !nbytes), "aes_ctr_fill: null argument");
    ABY3CU_REQUIRE(byte_off % 8 == 0 && nbytes % 8 == 0, "aes_ctr_fill: offset and size must be multiples of 8");
    if (!nbytes) return 0;
    DeviceGuard g(ctx->device);
    AesKey k; host_expand_key(key, &k);
    const size_t n = nbytes / 8;
    if (use_wide_aes(ctx, (n + 1) / 2, byte_off / 8)) {
        if (enable_big_smem(k_fill<true>, kAesWideTableBytes)) return 1;
        k_fill<true><<<ctx->sm_count, kWideThreads, kAesWideTableBytes, ctx->stream>>>(k, byte_off / 8, (i64*)d_out, n, al16(d_out));
        return post_launch(ctx, "k_fill");
    }
    if (enable_big_smem(k_fill<false>)) return 1;
    const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, kCtasPerSm);
    k_fill<false><<<grid, kThreads, kAesTableBytes, ctx->stream>>>(k, byte_off / 8, (i64*)d_out, n, al16(d_out));
    return post_launch(ctx, "k_fill");
}

int aby3cu_zero_share(aby3cu_ctx* ctx, const u8 key_prev[16], const u8 key_next[16], u64 elem0,
                      const i64* d_addend, i64* d_out, size_t n, int binary) {
    ABY3CU_REQUIRE(ctx && key_prev && key_next && (d_out || !n), "zero_share: null argument");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    AesKey kp, kn; host_expand_key(key_prev, &kp); host_expand_key(key_next, &kn);
    const int vec = al16(d_out) && (!d_addend || al16(d_addend));
    const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, kCtasPerSm);
    if (use_wide_aes(ctx, (n + 1) / 2, elem0)) {
        if (binary) {
            if (enable_big_smem(k_zero_share<true, true>, kAesWideTableBytes)) return 1;
            k_zero_share<true, true><<<ctx->sm_count, kWideThreads, kAesWideTableBytes, ctx->stream>>>(kp, kn, elem0, d_addend, d_out, n, vec);
        } else {
            if (enable_big_smem(k_zero_share<false, true>, kAesWideTableBytes)) return 1;
            k_zero_share<false, true><<<ctx->sm_count, kWideThreads, kAesWideTableBytes, ctx->stream>>>(kp, kn, elem0, d_addend, d_out, n, vec);
        }
        return post_launch(ctx, "k_zero_share");
    }
    if (binary) {
        if (enable_big_smem(k_zero_share<true, false>)) return 1;
        k_zero_share<true, false><<<grid, kThreads, kAesTableBytes, ctx->stream>>>(kp, kn, elem0, d_addend, d_out, n, vec);
    } else {
        if (enable_big_smem(k_zero_share<false, false>)) return 1;
        k_zero_share<false, false><<<grid, kThreads, kAesTableBytes, ctx->stream>>>(kp, kn, elem0, d_addend, d_out, n, vec);
    }
    return post_launch(ctx, "k_zero_share");
}

int aby3cu_mul_hadamard(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                        const u8 key_prev[16], const u8 key_next[16], u64 elem0, i64* C0, size_t n) {
    ABY3CU_REQUIRE(ctx && ((A0 && A1 && B0 && B1 && C0) || !n), "mul_hadamard: null argument");
    ABY3CU_REQUIRE((key_prev == nullptr) == (key_next == nullptr), "mul_hadamard: give both keys or neither");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    const int vec = al16(A0) && al16(A1) && al16(B0) && al16(B1) && al16(C0);
    if (key_prev) {
        AesKey kp, kn; host_expand_key(key_prev, &kp); host_expand_key(key_next, &kn);
        if (use_wide_aes(ctx, (n + 1) / 2, elem0)) {
            if (enable_big_smem(k_mul_hadamard<true, true>, kAesWideTableBytes)) return 1;
            k_mul_hadamard<true, true><<<ctx->sm_count, kWideThreads, kAesWideTableBytes, ctx->stream>>>(A0, A1, B0, B1, kp, kn, elem0, C0, n, vec);
        } else {
            if (enable_big_smem(k_mul_hadamard<true, false>)) return 1;
            const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, kCtasPerSm);
            k_mul_hadamard<true, false><<<grid, kThreads, kAesTableBytes, ctx->stream>>>(A0, A1, B0, B1, kp, kn, elem0, C0, n, vec);
        }
    } else {
        const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, 8);
        k_mul_hadamard<false, false><<<grid, kThreads, 0, ctx->stream>>>(A0, A1, B0, B1, kZeroKey, kZeroKey, 0, C0, n, vec);
    }
    return post_launch(ctx, "k_mul_hadamard");
}

static int launch_trunc(aby3cu_ctx* ctx, bool crossterm, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                        const u8* key_next, u64 en, const u8* key_prev, u64 ep, u64 d,
                        i64* V, i64* R, i64* NEGR, i64* RT0, i64* RT1, size_t n,
                        const u64* iter = nullptr, u64 iter_stride = 0) {
    ABY3CU_REQUIRE((key_prev == nullptr) == (key_next == nullptr), "trunc: give both keys or neither");
    ABY3CU_REQUIRE(d + 2 < 64, "trunc: shift too large");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    int vec = al16(RT0) && al16(RT1);
    if (crossterm) vec = vec && al16(A0) && al16(A1) && al16(B0) && al16(B1) && al16(V);
    else vec = vec && (!R || al16(R)) && (!NEGR || al16(NEGR));
    const bool rnd = key_prev != nullptr;
    AesKey kn = kZeroKey, kp = kZeroKey;
    if (rnd) { host_expand_key(key_next, &kn); host_expand_key(key_prev, &kp); }
    const unsigned d2 = (unsigned)d + 2;
    // A context marked "corun" issues work that should run UNDER another party's tcgen05 GEMM (148 CTAs x 6 warps x 224
    // registers: two of the four register files of an SM are left with 2048 free registers, one 64-register warp).  A
    // 256-thread CTA needs two warps per register file and is never admitted next to the GEMM; a 128-thread CTA is.
    const bool wide = rnd && !iter && use_wide_aes(ctx, (n + 1) / 2, en, ep);
    const unsigned threads = wide ? kWideThreads : (rnd && ctx->corun) ? 128 : kThreads;
    const unsigned grid = wide ? (unsigned)ctx->sm_count : ew_grid(ctx, (n + 1) / 2, threads, rnd ? kCtasPerSm : 8);
    const size_t smem = wide ? kAesWideTableBytes : rnd ? kAesTableBytes : 0;
    trace_mark(ctx, "(k_trunc ready)");
#define ABY3CU_LAUNCH_TRUNC(C, Rn)                                                                              \
    do {                                                                                                        \
        if (wide) {                                                                                             \
            if (enable_big_smem(k_trunc<C, Rn, Rn>, kAesWideTableBytes)) return 1;                              \
            k_trunc<C, Rn, Rn><<<grid, threads, smem, ctx->stream>>>(A0, A1, B0, B1, kn, en, kp, ep, d2, V, R, NEGR, \
                                                                     RT0, RT1, n, vec, iter, iter_stride);      \
        } else {                                                                                                \
            if (Rn && enable_big_smem(k_trunc<C, Rn, false>)) return 1;                                         \
            k_trunc<C, Rn, false><<<grid, threads, smem, ctx->stream>>>(A0, A1, B0, B1, kn, en, kp, ep, d2, V, R, NEGR, \
                                                                        RT0, RT1, n, vec, iter, iter_stride);   \
        }                                                                                                       \
    } while (0)
    if (crossterm) { if (rnd) ABY3CU_LAUNCH_TRUNC(true, true); else ABY3CU_LAUNCH_TRUNC(true, false); }
    else           { if (rnd) ABY3CU_LAUNCH_TRUNC(false, true); else ABY3CU_LAUNCH_TRUNC(false, false); }
#undef ABY3CU_LAUNCH_TRUNC
    return post_launch(ctx, "k_trunc");
}

int aby3cu_mul_hadamard_trunc(aby3cu_ctx* ctx, const i64* A0, const i64* A1, const i64* B0, const i64* B1,
                              const u8 key_next_common[16], u64 elem_next, const u8 key_prev_common[16], u64 elem_prev,
                              u64 d, i64* V, i64* RT0, i64* RT1, size_t n) {
    ABY3CU_REQUIRE(ctx && ((A0 && A1 && B0 && B1 && V && RT0 && RT1) || !n), "mul_hadamard_trunc: null argument");
    return launch_trunc(ctx, true, A0, A1, B0, B1, key_next_common, elem_next, key_prev_common, elem_prev, d,
                        V, nullptr, nullptr, RT0, RT1, n);
}

int aby3cu_trunc_tuple(aby3cu_ctx* ctx, const u8 key_next_common[16], u64 elem_next, const u8 key_prev_common[16],
                       u64 elem_prev, u64 d, i64* R, i64* NEGR, i64* RT0, i64* RT1, size_t n) {
    ABY3CU_REQUIRE(ctx && ((RT0 && RT1) || !n), "trunc_tuple: null argument");
    return launch_trunc(ctx, false, nullptr, nullptr, nullptr, nullptr, key_next_common, elem_next, key_prev_common,
                        elem_prev, d, nullptr, R, NEGR, RT0, RT1, n);
}

int aby3cu_trunc_tuple_at(aby3cu_ctx* ctx, const u8 key_next_common[16], u64 elem_next, const u8 key_prev_common[16],
                          u64 elem_prev, const u64* d_iter, u64 iter_stride, u64 d, i64* R, i64* NEGR, i64* RT0, i64* RT1, size_t n) {
    ABY3CU_REQUIRE(ctx && ((RT0 && RT1) || !n), "trunc_tuple_at: null argument");
    ABY3CU_REQUIRE(key_next_common && key_prev_common && d_iter, "trunc_tuple_at: keys and the iteration counter are required");
    return launch_trunc(ctx, false, nullptr, nullptr, nullptr, nullptr, key_next_common, elem_next, key_prev_common,
                        elem_prev, d, nullptr, R, NEGR, RT0, RT1, n, d_iter, iter_stride);
}

int aby3cu_trunc_finish(aby3cu_ctx* ctx, const i64* s0, const i64* s1, const i64* s2, i64* C, size_t n, u64 shift) {
    ABY3CU_REQUIRE(ctx && ((s0 && s1 && s2 && C) || !n), "trunc_finish: null argument");
    ABY3CU_REQUIRE(shift < 64, "trunc_finish: shift too large");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    const int vec = al16(s0) && al16(s1) && al16(s2) && al16(C);
    const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, 8);
    k_trunc_finish<<<grid, kThreads, 0, ctx->stream>>>(s0, s1, s2, C, n, (unsigned)shift, vec);
    return post_launch(ctx, "k_trunc_finish");
}

int aby3cu_share_op(aby3cu_ctx* ctx, int op, const i64* x, const i64* y, i64* out, size_t n) {
    ABY3CU_REQUIRE(ctx && ((x && y && out) || !n), "share_op: null argument");
    ABY3CU_REQUIRE(op >= 0 && op <= 2, "share_op: bad op");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    const int vec = al16(x) && al16(y) && al16(out);
    const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, 8);
    if (op == ABY3CU_OP_ADD) k_share_op<ABY3CU_OP_ADD><<<grid, kThreads, 0, ctx->stream>>>(x, y, out, n, vec);
    else if (op == ABY3CU_OP_SUB) k_share_op<ABY3CU_OP_SUB><<<grid, kThreads, 0, ctx->stream>>>(x, y, out, n, vec);
    else k_share_op<ABY3CU_OP_XOR><<<grid, kThreads, 0, ctx->stream>>>(x, y, out, n, vec);
    return post_launch(ctx, "k_share_op");
}

int aby3cu_share_op2(aby3cu_ctx* ctx, int op, const i64* x0, const i64* y0, i64* out0, const i64* x1, const i64* y1, i64* out1, size_t n) {
    ABY3CU_REQUIRE(ctx && ((x0 && y0 && out0 && x1 && y1 && out1) || !n), "share_op2: null argument");
    ABY3CU_REQUIRE(op >= 0 && op <= 2, "share_op2: bad op");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    const int vec = al16(x0) && al16(y0) && al16(out0) && al16(x1) && al16(y1) && al16(out1);
    unsigned gx = ew_grid(ctx, (n + 1) / 2, kThreads, 4);
    const dim3 grid(gx, 2);
    Planes2 p{{x0, x1}, {y0, y1}, {out0, out1}};
    if (op == ABY3CU_OP_ADD) k_share_op2<ABY3CU_OP_ADD><<<grid, kThreads, 0, ctx->stream>>>(p, n, vec);
    else if (op == ABY3CU_OP_SUB) k_share_op2<ABY3CU_OP_SUB><<<grid, kThreads, 0, ctx->stream>>>(p, n, vec);
    else k_share_op2<ABY3CU_OP_XOR><<<grid, kThreads, 0, ctx->stream>>>(p, n, vec);
    return post_launch(ctx, "k_share_op2");
}

int aby3cu_combine3(aby3cu_ctx* ctx, int op, const i64* x0, const i64* x1, const i64* x2, i64* out, size_t n) {
    ABY3CU_REQUIRE(ctx && ((x0 && x1 && x2 && out) || !n), "combine3: null argument");
    ABY3CU_REQUIRE(op == ABY3CU_OP_ADD || op == ABY3CU_OP_XOR, "combine3: bad op");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    const int vec = al16(x0) && al16(x1) && al16(x2) && al16(out);
    const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, 8);
    if (op == ABY3CU_OP_ADD) k_combine3<ABY3CU_OP_ADD><<<grid, kThreads, 0, ctx->stream>>>(x0, x1, x2, out, n, vec);
    else k_combine3<ABY3CU_OP_XOR><<<grid, kThreads, 0, ctx->stream>>>(x0, x1, x2, out, n, vec);
    return post_launch(ctx, "k_combine3");
}

int aby3cu_axpb(aby3cu_ctx* ctx, i64 a, const i64* x, i64 b, i64* out, size_t n) {
    ABY3CU_REQUIRE(ctx && (out || !n), "axpb: null argument");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    const int vec = al16(out) && (!x || al16(x));
    const unsigned grid = ew_grid(ctx, (n + 1) / 2, kThreads, 8);
    k_axpb<<<grid, kThreads, 0, ctx->stream>>>((u64)a, x, (u64)b, out, n, vec);
    return post_launch(ctx, "k_axpb");
}

int aby3cu_transpose_i64(aby3cu_ctx* ctx, const i64* in, u64 rows, u64 cols, i64* out) {
    ABY3CU_REQUIRE(ctx && ((in && out) || !(rows * cols)), "transpose: null argument");
    if (!(rows * cols)) return 0;
    DeviceGuard g(ctx->device);
    const u64 tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    const unsigned grid = (unsigned)(tiles < (u64)ctx->sm_count * 8 ? tiles : (u64)ctx->sm_count * 8);
    k_transpose<<<grid, 256, 0, ctx->stream>>>(in, rows, cols, out);
    return post_launch(ctx, "k_transpose");
}

int aby3cu_transpose_i64_2(aby3cu_ctx* ctx, const i64* in0, const i64* in1, u64 rows, u64 cols, i64* out0, i64* out1) {
    ABY3CU_REQUIRE(ctx && ((in0 && in1 && out0 && out1) || !(rows * cols)), "transpose2: null argument");
    if (!(rows * cols)) return 0;
    DeviceGuard g(ctx->device);
    const u64 tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    const unsigned gx = (unsigned)(tiles < (u64)ctx->sm_count * 4 ? tiles : (u64)ctx->sm_count * 4);
    k_transpose2<<<dim3(gx, 2), 256, 0, ctx->stream>>>(in0, in1, rows, cols, out0, out1);
    return post_launch(ctx, "k_transpose2");
}

int aby3cu_gather_rows_multi(aby3cu_ctx* ctx, int njobs, const i64* const* in, const u64* cols, i64* const* out, const u64* idx, u64 nrows) {
    return aby3cu_gather_rows_multi_at(ctx, njobs, in, cols, out, idx, nrows, nullptr);
}

__global__ void k_counter_add(u64* c, u64 inc) { *c += inc; }

// x[r, cols-1] &= mask for every row (sbMatrix::trim, Sh3Types.h:128-160; Sh3Converter.cpp:97-106)
__global__ void __launch_bounds__(256) k_mask_last_word(i64* __restrict__ x, u64 rows, u64 cols, u64 mask) {
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (u64)gridDim.x * blockDim.x)
        x[r * cols + cols - 1] = (i64)((u64)x[r * cols + cols - 1] & mask);
}

int aby3cu_mask_last_word(aby3cu_ctx* ctx, i64* d_x, u64 rows, u64 cols, u64 mask) {
    ABY3CU_REQUIRE(ctx && (d_x || !(rows * cols)), "mask_last_word: null argument");
    if (!(rows * cols)) return 0;
    DeviceGuard g(ctx->device);
    k_mask_last_word<<<ew_grid(ctx, rows, 256, 8), 256, 0, ctx->stream>>>(d_x, rows, cols, mask);
    return post_launch(ctx, "k_mask_last_word");
}

int aby3cu_counter_add(aby3cu_ctx* ctx, u64* d_counter, u64 inc) {
    ABY3CU_REQUIRE(ctx && d_counter, "counter_add: null argument");
    DeviceGuard g(ctx->device);
    k_counter_add<<<1, 1, 0, ctx->stream>>>(d_counter, inc);
    return post_launch(ctx, "k_counter_add");
}

int aby3cu_gather_rows_multi_at(aby3cu_ctx* ctx, int njobs, const i64* const* in, const u64* cols, i64* const* out, const u64* idx, u64 nrows,
                                const u64* d_iter) {
    ABY3CU_REQUIRE(ctx && in && cols && out, "gather_rows_multi: null argument");
    ABY3CU_REQUIRE(njobs >= 1 && njobs <= ABY3CU_MAX_GATHER_JOBS, "gather_rows_multi: bad job count");
    if (!nrows) return 0;
    ABY3CU_REQUIRE(idx, "gather_rows_multi: null index vector");
    GatherJobs jobs;
    u64 maxc = 1;
    for (int j = 0; j < ABY3CU_MAX_GATHER_JOBS; ++j) {
        const int k = j < njobs ? j : 0;
        ABY3CU_REQUIRE(in[k] && out[k] && cols[k], "gather_rows_multi: null / empty job");
        jobs.in[j] = in[k]; jobs.out[j] = out[k]; jobs.cols[j] = cols[k];
        if (cols[k] > maxc) maxc = cols[k];
    }
    DeviceGuard g(ctx->device);
    const unsigned gx = ew_grid(ctx, maxc == 1 ? nrows : nrows * 32, 256, 4);
    k_gather_rows_multi<<<dim3(gx, (unsigned)njobs), 256, 0, ctx->stream>>>(jobs, idx, nrows, d_iter);
    return post_launch(ctx, "k_gather_rows_multi");
}

int aby3cu_iota_u64(aby3cu_ctx* ctx, u64 start, u64 step, u64* out, size_t n) {
    ABY3CU_REQUIRE(ctx && (out || !n), "iota: null argument");
    if (!n) return 0;
    DeviceGuard g(ctx->device);
    k_iota<<<ew_grid(ctx, n, 256, 8), 256, 0, ctx->stream>>>(start, step, out, n);
    return post_launch(ctx, "k_iota");
}

int aby3cu_cmpx_gather(aby3cu_ctx* ctx, const i64* src0, const i64* src1, u64 r0, u64 d, u64 m, i64* x0, i64* x1, i64* y0, i64* y1) {
    ABY3CU_REQUIRE(ctx && ((src0 && src1 && x0 && x1 && y0 && y1) || !m), "cmpx_gather: null argument");
    ABY3CU_REQUIRE(d & 1, "cmpx_gather: the distance must be odd (the two operands are the two parities of one range)");
    if (!m) return 0;
    DeviceGuard g(ctx->device);
    CmpxPlanes P = {{src0, src1}, {x0, x1}, {y0, y1}};
    k_cmpx_gather<<<dim3(ew_grid(ctx, m + d / 2, 256, 8), 2), 256, 0, ctx->stream>>>(P, r0, d / 2, m);
    return post_launch(ctx, "k_cmpx_gather");
}

int aby3cu_cmpx_scatter(aby3cu_ctx* ctx, const i64* x0, const i64* x1, const i64* y0, const i64* y1, u64 r0, u64 d, u64 m, i64* dst0, i64* dst1) {
    ABY3CU_REQUIRE(ctx && ((dst0 && dst1 && x0 && x1 && y0 && y1) || !m), "cmpx_scatter: null argument");
    ABY3CU_REQUIRE(d & 1, "cmpx_scatter: the distance must be odd");
    if (!m) return 0;
    DeviceGuard g(ctx->device);
    CmpxScatter P = {{dst0, dst1}, {x0, x1}, {y0, y1}};
    k_cmpx_scatter<<<dim3(ew_grid(ctx, m + d / 2, 256, 8), 2), 256, 0, ctx->stream>>>(P, r0, d / 2, m);
    return post_launch(ctx, "k_cmpx_scatter");
}

int aby3cu_scatter_rows(aby3cu_ctx* ctx, const i64* in, u64 cols, const u64* idx, u64 nrows, i64* out) {
    ABY3CU_REQUIRE(ctx && ((in && idx && out) || !(nrows * cols)), "scatter_rows: null argument");
    if (!(nrows * cols)) return 0;
    DeviceGuard g(ctx->device);
    const unsigned grid = ew_grid(ctx, cols == 1 ? nrows : nrows * 32, 256, 8);
    k_scatter_rows<<<grid, 256, 0, ctx->stream>>>(in, cols, idx, nrows, out);
    return post_launch(ctx, "k_scatter_rows");
}

int aby3cu_gather_rows(aby3cu_ctx* ctx, const i64* in, u64 cols, const u64* idx, u64 nrows, i64* out) {
    ABY3CU_REQUIRE(ctx && ((in && idx && out) || !(nrows * cols)), "gather_rows: null argument");
    if (!(nrows * cols)) return 0;
    DeviceGuard g(ctx->device);
    const unsigned grid = ew_grid(ctx, cols == 1 ? nrows : nrows * 32, 256, 8);
    k_gather_rows<<<grid, 256, 0, ctx->stream>>>(in, cols, idx, nrows, out);
    return post_launch(ctx, "k_gather_rows");
}

}  // extern "C"
