/*
 * aby3cu.h -- C ABI of the B200 (sm_100a) implementation of ABY3's
 * replicated-share multiplication hot path.
 *
 * The reference (Fannxy/aby3) has no FFI seam: its hot loops sit inside the C++
 * classes of aby3/sh3/.  Each entry point below replaces one of those CPU loops
 * (cited as path:line relative to the reference root) and is what the sh3
 * facade in aby3_b200/sh3/ binds.  INTEGRATION.md shows the binding a reference
 * maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure;
 *    aby3cu_last_error() gives the message (thread local).  Nothing throws
 *    across the ABI.
 *  - pointers named d_* are device pointers owned by the caller; all work is
 *    enqueued on the context's stream and is asynchronous unless stated.
 *  - "key" is a 16-byte AES-128 key in memory order (oc::block / oc::AES::setKey).
 *  - keystream KS_key = AES_key(toBlock(0)) || AES_key(toBlock(1)) || ...  where
 *    toBlock(c) is c as 8 little-endian bytes followed by 8 zero bytes
 *    (oc::AES::ecbEncCounterMode).  "stream element e" is the little-endian u64
 *    at keystream bytes [8e, 8e+8).
 *  - all arithmetic is wrapping 64-bit; right shifts are arithmetic.
 *  - there is no CPU fallback: with no usable sm_100 device aby3cu_ctx_create fails.
 */
#ifndef ABY3CU_H
#define ABY3CU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ABY3CU_VERSION 1

typedef struct aby3cu_ctx aby3cu_ctx;

/* ---- library / context ---------------------------------------------------- */
int aby3cu_version(void);
const char* aby3cu_last_error(void);
int aby3cu_device_count(int* count);
/* One context per party: a device, a stream, scratch space. */
int aby3cu_ctx_create(int device, aby3cu_ctx** out);
/* Same, but work is enqueued on a caller-owned cudaStream_t (e.g. a torch stream). */
int aby3cu_ctx_create_on_stream(int device, void* cuda_stream, aby3cu_ctx** out);
int aby3cu_ctx_destroy(aby3cu_ctx* ctx);
/* Mark a context whose work is meant to run UNDER another context's tensor-core GEMM on the same GPU (a party's
 * second stream: the truncation pairs of the next product, Sh3Evaluator.cpp:503-566 -- input independent): its
 * keystream kernels use CTAs small enough to be co-resident with the GEMM's. */
int aby3cu_ctx_set_corun(aby3cu_ctx* ctx, int on);
/* Kernel timeline without a profiler (ABY3CU_TRACE=1 in the environment, else both are no-ops): after trace_begin every
 * launch of the library records an event behind itself; trace_dump synchronises the device and writes
 * "stream,kernel,end_ms" (milliseconds since trace_begin) for all of them. */
int aby3cu_trace_begin(aby3cu_ctx* ctx);
/* a named mark on the context's stream (the string must outlive the dump): its time is when the stream got there */
int aby3cu_trace_mark(aby3cu_ctx* ctx, const char* static_name);
int aby3cu_trace_dump(const char* path);
int aby3cu_ctx_device(const aby3cu_ctx* ctx);
void* aby3cu_ctx_stream(const aby3cu_ctx* ctx);
int aby3cu_sync(aby3cu_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t aby3cu_launch_count(const aby3cu_ctx* ctx);

/* ---- memory, copies, events ------------------------------------------------ */
int aby3cu_malloc(aby3cu_ctx* ctx, void** d_ptr, size_t bytes);
int aby3cu_free(aby3cu_ctx* ctx, void* d_ptr);
int aby3cu_memset(aby3cu_ctx* ctx, void* d_ptr, int byte, size_t bytes);
int aby3cu_host_alloc(void** h_ptr, size_t bytes);          /* pinned host memory */
int aby3cu_host_free(void* h_ptr);
int aby3cu_h2d(aby3cu_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int aby3cu_d2h(aby3cu_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* device-to-device; dst_device/src_device may differ (NVLink peer copy) */
int aby3cu_d2d(aby3cu_ctx* ctx, void* d_dst, int dst_device, const void* d_src, int src_device, size_t bytes);
int aby3cu_event_create(aby3cu_ctx* ctx, void** event);
/* CUDA graphs for latency-bound protocol loops (one SGD iteration of aby3-ML/Regression.h:142-171 is ~28 small
 * kernels): capture_begin starts recording the work enqueued on ctx's stream -- and on the streams of other contexts
 * that join it through event_record / event_wait -- capture_end returns an executable graph, graph_launch replays it on
 * ctx's stream (kernel_nodes = kernels inside, added to the launch count).  Kernel arguments are frozen at capture; the
 * *_at entry points read the per-iteration offsets from a device counter instead. */
int aby3cu_capture_begin(aby3cu_ctx* ctx);
int aby3cu_capture_end(aby3cu_ctx* ctx, void** graph_exec);
int aby3cu_graph_launch(aby3cu_ctx* ctx, void* graph_exec, uint64_t kernel_nodes);
int aby3cu_graph_destroy(void* graph_exec);
int aby3cu_event_create_sync(aby3cu_ctx* ctx, void** event);  /* ordering only (no timing): cheaper to record */
int aby3cu_event_destroy(void* event);
int aby3cu_event_record(aby3cu_ctx* ctx, void* event);       /* on ctx's stream */
int aby3cu_event_wait(aby3cu_ctx* ctx, void* event);         /* ctx's stream waits */
int aby3cu_event_sync(void* event);
int aby3cu_event_elapsed_ms(void* start, void* stop, float* ms);

/* ---- AES-CTR keystream ------------------------------------------------------ */
/* Host-side keystream bytes for the small draws the facade makes itself
 * (oc::PRNG::get<block>() at Sh3ShareGen.h:19-20, Sh3Evaluator.cpp:13-14,
 * Sh3BinaryEvaluator.h:99-100).  Pure host code, bounded to 4096 bytes. */
int aby3cu_host_keystream(const uint8_t key[16], uint64_t byte_off, size_t nbytes, uint8_t* out);
/* Bulk oc::PRNG::get(ptr, n) (Sh3Evaluator.cpp:526-527): device fill with
 * keystream bytes [byte_off, byte_off + nbytes); both must be multiples of 8. */
int aby3cu_aes_ctr_fill(aby3cu_ctx* ctx, const uint8_t key[16], uint64_t byte_off,
                        void* d_out, size_t nbytes);

/* ---- zero sharing: Sh3ShareGen::getShare / getBinaryShare loops -------------- */
/* out[i] = addend[i] (+ or ^) (KS_prev[e0+i] (- or ^) KS_next[e0+i]),  i < n.
 * d_addend may be NULL (remoteIntMatrix).  Replaces the per-element loops at
 * Sh3Encryptor.cpp:222-223, 258-259, 303-304 and Sh3ShareGen.h:60-92. */
int aby3cu_zero_share(aby3cu_ctx* ctx, const uint8_t key_prev[16], const uint8_t key_next[16],
                      uint64_t elem0, const int64_t* d_addend, int64_t* d_out, size_t n, int binary);

/* ---- arithmetic multiplication: Sh3Evaluator::asyncMul ------------------------ */
/* Hadamard form of this fork (Sh3Evaluator.cpp:101-105):
 * C0[i] = A0[i]*B0[i] + A0[i]*B1[i] + A1[i]*B0[i] + z(e0+i). */
int aby3cu_mul_hadamard(aby3cu_ctx* ctx, const int64_t* d_A0, const int64_t* d_A1,
                        const int64_t* d_B0, const int64_t* d_B1,
                        const uint8_t key_prev[16], const uint8_t key_next[16], uint64_t elem0,
                        int64_t* d_C0, size_t n);
/* Hadamard form with truncation pair (Sh3Evaluator.cpp:667-673 + 526-537):
 * t0 = KS_nextCommon[e_next+i], t1 = KS_prevCommon[e_prev+i],
 * V[i] = cross(i) - (t0 >> 2), RT0[i] = t0 >> (d+2), RT1[i] = t1 >> (d+2).
 * keys NULL => DEBUG_disable_randomization (r = 0, RT = 0). */
int aby3cu_mul_hadamard_trunc(aby3cu_ctx* ctx, const int64_t* d_A0, const int64_t* d_A1,
                              const int64_t* d_B0, const int64_t* d_B1,
                              const uint8_t key_next_common[16], uint64_t elem_next,
                              const uint8_t key_prev_common[16], uint64_t elem_prev,
                              uint64_t d, int64_t* d_V, int64_t* d_RT0, int64_t* d_RT1, size_t n);
/* Sh3Evaluator::getTruncationTuple (Sh3Evaluator.cpp:503-566).  Writes RT0/RT1 and,
 * when non-NULL, R[i] = t0>>2 and NEGR[i] = -(t0>>2) (NEGR pre-loads the GEMM
 * accumulator so that abMinusR needs no extra pass).  keys NULL => all zero. */
int aby3cu_trunc_tuple(aby3cu_ctx* ctx, const uint8_t key_next_common[16], uint64_t elem_next,
                       const uint8_t key_prev_common[16], uint64_t elem_prev, uint64_t d,
                       int64_t* d_R, int64_t* d_NEGR, int64_t* d_RT0, int64_t* d_RT1, size_t n);
/* getTruncationTuple replayable inside a CUDA graph: the stream offsets are elem_next / elem_prev + *d_iter * iter_stride */
int aby3cu_trunc_tuple_at(aby3cu_ctx* ctx, const uint8_t key_next_common[16], uint64_t elem_next, const uint8_t key_prev_common[16],
                          uint64_t elem_prev, const uint64_t* d_iter, uint64_t iter_stride, uint64_t d,
                          int64_t* d_R, int64_t* d_negR, int64_t* d_RT0, int64_t* d_RT1, size_t n);
/* Open-and-truncate continuation (Sh3Evaluator.cpp:712-718):
 * C[i] += (s0[i] + s1[i] + s2[i]) >> shift  (arithmetic). */
int aby3cu_trunc_finish(aby3cu_ctx* ctx, const int64_t* d_s0, const int64_t* d_s1, const int64_t* d_s2,
                        int64_t* d_C, size_t n, uint64_t shift);

/* The matrix cross term  C (+)= A0*B0 + A0*B1 + A1*B0  over Z_2^64
 * (Sh3Evaluator.cpp:662-665, upstream form of :96-99), A* row-major MxK,
 * B* row-major KxN, C row-major MxN.  accumulate != 0 adds into C (C was
 * pre-loaded with the zero share z or with -r). */
enum { ABY3CU_GEMM_AUTO = 0, ABY3CU_GEMM_IMAD = 1, ABY3CU_GEMM_TCGEN05 = 2 };
int aby3cu_gemm_cross(aby3cu_ctx* ctx, int algo,
                      const int64_t* d_A0, const int64_t* d_A1,
                      const int64_t* d_B0, const int64_t* d_B1,
                      uint64_t M, uint64_t K, uint64_t N, int64_t* d_C, int accumulate);
/* The same; the first kernel that touches C waits for `c_ready` (an aby3cu event, may be NULL): C was pre-loaded on
 * another stream (-r of the truncation pair, produced ahead of the product: Sh3Evaluator.cpp:670-672), the limb
 * pre-pass of the operands does not wait. */
int aby3cu_gemm_cross_after(aby3cu_ctx* ctx, int algo,
                            const int64_t* d_A0, const int64_t* d_A1,
                            const int64_t* d_B0, const int64_t* d_B1,
                            uint64_t M, uint64_t K, uint64_t N, int64_t* d_C, int accumulate, void* c_ready);
/* The same product delivered in ROW BLOCKS: block b = rows [b * block_rows, (b + 1) * block_rows) of C is final when
 * block_events[b] (caller-created events, n_blocks = ceil(M / block_rows), block_rows a multiple of 128) has been
 * reached on the context's stream -- the reshare of xy - r (Sh3Evaluator.cpp:681-684) of block b can then travel on
 * another stream while block b + 1 is still being multiplied.  Every event is recorded even on failure. */
int aby3cu_gemm_cross_blocks(aby3cu_ctx* ctx, int algo, const int64_t* d_A0, const int64_t* d_A1, const int64_t* d_B0,
                             const int64_t* d_B1, uint64_t M, uint64_t K, uint64_t N, int64_t* d_C, int accumulate,
                             void* c_ready_event, uint64_t block_rows, void** block_events, uint32_t n_blocks);
/* algo actually used by the last aby3cu_gemm_cross on this context */
int aby3cu_gemm_last_algo(const aby3cu_ctx* ctx);
/* device time (CUDA events on the context's stream) of the main GEMM kernel of the
 * last aby3cu_gemm_cross call, pre-pass excluded; exact when that call needed a
 * single main-kernel launch (it synchronises on the closing event). */
int aby3cu_gemm_last_main_kernel_ms(aby3cu_ctx* ctx, float* ms);

/* ---- 3-party shared OT and the bit x arithmetic product ---------------- */
/* aby3/OT/SharedOT.cpp:6-28: out[i][h] = msgs[i][h] ^ pad_i[h], pad_i = AES_key(toBlock(idx0+i)),
 * h = 0 low 8 bytes, h = 1 high 8 bytes.  d_msgs / d_out: n pairs of int64, 16-byte aligned. */
int aby3cu_ot_send(aby3cu_ctx* ctx, const uint8_t key[16], uint64_t idx0, const int64_t* d_msgs, int64_t* d_out, size_t n);
/* SharedOT::help (:30-94): out[i] = pad_i[choice[i] & 1] */
int aby3cu_ot_help(aby3cu_ctx* ctx, const uint8_t key[16], uint64_t idx0, const int64_t* d_choice, int64_t* d_out, size_t n);
/* SharedOT::recv (:102-126): out[i] (+)= masked[i][choice[i] & 1] ^ help[i] */
int aby3cu_ot_recv(aby3cu_ctx* ctx, const int64_t* d_masked, const int64_t* d_help, const int64_t* d_choice,
                   int64_t* d_out, size_t n, int accumulate);
/* Message pairs of asyncMul(si64Matrix, sbMatrix) (Sh3Evaluator.cpp:133-160): party 0 draws
 * z, c1 from mPrevCommon (elements elem_prev+2i, +2i+1) and c0 from mNextCommon (elem_next+i). */
int aby3cu_bitmul_msgs_p0(aby3cu_ctx* ctx, const int64_t* d_A0, const int64_t* d_A1, const int64_t* d_B0, const int64_t* d_B1,
                          const uint8_t key_prev_common[16], uint64_t elem_prev,
                          const uint8_t key_next_common[16], uint64_t elem_next,
                          int64_t* d_C0, int64_t* d_C1, int64_t* d_msgs, size_t n);
/* party 2 (:212-240): z, c0 from mNextCommon (elements elem_next+2i, +2i+1) */
int aby3cu_bitmul_msgs_p2(aby3cu_ctx* ctx, const int64_t* d_A1, const int64_t* d_B0, const int64_t* d_B1,
                          const uint8_t key_next_common[16], uint64_t elem_next, int64_t* d_C0, int64_t* d_msgs, size_t n);
/* party 0 of asyncMul(i64 a, sbMatrix) (:430-447): pairs (z, a+z) ordered by b0^b1, z = getShare() */
int aby3cu_bitmul_pub_msgs(aby3cu_ctx* ctx, int64_t a, const int64_t* d_B0, const int64_t* d_B1,
                           const uint8_t key_prev[16], const uint8_t key_next[16], uint64_t elem0, int64_t* d_msgs, size_t n);

/* ---- Sh3Converter::bitInjection (aby3/sh3/Sh3Converter.cpp:211-370) ------------------------ */
/* choice vectors (:244-247, 282-285): out[i*bit_count + j] = bit j of row i of a rows x words binary share plane */
int aby3cu_bits_expand(aby3cu_ctx* ctx, const int64_t* d_in, uint64_t rows, uint64_t words, uint64_t bit_count, int64_t* d_out);
/* party 2, the OT sender (:318-347): d0 / d1 = the next rows*bit_count words of the nextCommon / prevCommon
 * streams (its output share planes); message pair k = (m, m) with m = -d0[k] - d1[k], plus 1 in slot b_k ^ 1,
 * b_k = bit k of in0 ^ in1 */
int aby3cu_bitinj_msgs(aby3cu_ctx* ctx, const int64_t* d_in0, const int64_t* d_in1, uint64_t rows, uint64_t words, uint64_t bit_count,
                       const uint8_t key_next_common[16], uint64_t elem_next, const uint8_t key_prev_common[16], uint64_t elem_prev,
                       int64_t* d_d0, int64_t* d_d1, int64_t* d_msgs);

/* ---- local share arithmetic / reveal ------------------------------------------- */
enum { ABY3CU_OP_ADD = 0, ABY3CU_OP_SUB = 1, ABY3CU_OP_XOR = 2 };
/* out = x op y  (sMatrix +,-: Sh3Types.h:805-820) */
int aby3cu_share_op(aby3cu_ctx* ctx, int op, const int64_t* d_x, const int64_t* d_y, int64_t* d_out, size_t n);
/* the same on BOTH share planes of a replicated sharing in one launch (sMatrix::operator+/-, Sh3Types.h:805-820
 * touch mShares[0] and mShares[1]); latency-bound callers (SGD) halve their launches */
int aby3cu_share_op2(aby3cu_ctx* ctx, int op, const int64_t* d_x0, const int64_t* d_y0, int64_t* d_out0,
                     const int64_t* d_x1, const int64_t* d_y1, int64_t* d_out1, size_t n);
/* out = x0 op x1 op x2 (reveal: Sh3Encryptor.cpp:497-536), op ADD or XOR */
int aby3cu_combine3(aby3cu_ctx* ctx, int op, const int64_t* d_x0, const int64_t* d_x1, const int64_t* d_x2,
                    int64_t* d_out, size_t n);
/* out = a * x + b element-wise (wrapping); x == NULL fills out with b.  Local affine maps of
 * Sh3Piecewise::getFunctionValues / threshold shifts (Sh3Piecewise.cpp:441-451, 541-563). */
int aby3cu_axpb(aby3cu_ctx* ctx, int64_t a, const int64_t* d_x, int64_t b, int64_t* d_out, size_t n);
/* row-major transpose of an int64 matrix (sMatrix::transpose, Sh3Types.h:822-838) */
int aby3cu_transpose_i64(aby3cu_ctx* ctx, const int64_t* d_in, uint64_t rows, uint64_t cols, int64_t* d_out);
/* both share planes in one launch (sMatrix::transpose, Sh3Types.h:822-838) */
int aby3cu_transpose_i64_2(aby3cu_ctx* ctx, const int64_t* d_in0, const int64_t* d_in1, uint64_t rows, uint64_t cols,
                           int64_t* d_out0, int64_t* d_out1);
/* up to ABY3CU_MAX_GATHER_JOBS row gathers sharing one index vector in one launch: job j copies rows
 * idx[0..nrows) of in[j] (cols[j] wide) to out[j]  (extractBatch takes the same rows of X and Y, both planes) */
#define ABY3CU_MAX_GATHER_JOBS 12
int aby3cu_gather_rows_multi(aby3cu_ctx* ctx, int njobs, const int64_t* const* d_in, const uint64_t* cols,
                             int64_t* const* d_out, const uint64_t* d_idx, uint64_t nrows);
/* the same, replayable inside a CUDA graph: the batch is idx[*d_iter * nrows ...) (d_iter: device counter, may be NULL) */
int aby3cu_gather_rows_multi_at(aby3cu_ctx* ctx, int njobs, const int64_t* const* d_in, const uint64_t* cols,
                                int64_t* const* d_out, const uint64_t* d_idx, uint64_t nrows, const uint64_t* d_iter);
/* x[r, cols-1] &= mask for every row of a rows x cols word matrix: keeps the low bitCount % 64 bits of binary shares
 * (sbMatrix::trim, Sh3Types.h:128-160, 383-386; Sh3Converter.cpp:97-106) */
int aby3cu_mask_last_word(aby3cu_ctx* ctx, int64_t* d_x, uint64_t rows, uint64_t cols, uint64_t mask);
/* ---- SGD_Linear (aby3-ML/Regression.h:112-184) for three parties on ONE GPU as ONE persistent kernel --------------
 * `iters` iterations of: XX = X[batch rows], error = mul(XX, w) - YY (shift1 = D), update = mul(XX^T, error) (shift2 =
 * D + log2(B / lr)), w -= update -- the truncating products as in Sh3Evaluator.cpp:651-724 with the truncation pairs drawn
 * from the common keystreams (key_next[p] / key_prev[p], first elements elem_next[p] / elem_prev[p]; an iteration consumes
 * B + F elements of each stream: B for the first product, then F).  d_X, d_Y, d_w: [2 * party + share plane], X rows x F,
 * Y rows x 1, w F x 1 (updated in place); d_batch_idx: iters x B row indices.  F must be even.  d_work: scratch of
 * aby3cu_sgd_linear_colocated_work_bytes(B) bytes.  The grid is launched cooperatively (two grid barriers per iteration). */
size_t aby3cu_sgd_linear_colocated_work_bytes(uint64_t B);
int aby3cu_sgd_linear_colocated(aby3cu_ctx* ctx, const int64_t* const* d_X, const int64_t* const* d_Y, int64_t* const* d_w,
                                const uint64_t* d_batch_idx, uint64_t F, uint64_t B, uint64_t iters, uint64_t shift1,
                                uint64_t shift2, const uint8_t* const* key_next, const uint64_t* elem_next,
                                const uint8_t* const* key_prev, const uint64_t* elem_prev, void* d_work);
/* *d_counter += inc on the context's stream (the iteration counter of a replayed graph) */
int aby3cu_counter_add(aby3cu_ctx* ctx, uint64_t* d_counter, uint64_t inc);
/* ---- the small kernels of a protocol step, batched over independent problems (blockIdx.y) -------------------------
 * Parties that share a GPU run the SAME kernel on three sets of pointers at every step of a latency-bound loop (SGD,
 * aby3-ML/Regression.h:142-171); one launch for all of them shortens the replayed CUDA graph (aby3_b200/ml/SgdGraph.h).
 * Arithmetic and keystream offsets are those of the single-problem entry points; at most ABY3CU_MAX_BATCH problems. */
#define ABY3CU_MAX_BATCH 6
/* getTruncationTuple (Sh3Evaluator.cpp:503-566) for njobs (key pair, offsets, shift, count) tuples: negr = -(t0>>2),
 * rt0 = t0 >> (shift+2), rt1 = t1 >> (shift+2); stream offsets elem + *d_iter * iter_stride (d_iter may be NULL) */
int aby3cu_trunc_tuple_batch_at(aby3cu_ctx* ctx, int njobs, const uint8_t* const* keys_next, const uint64_t* elem_next,
                                const uint8_t* const* keys_prev, const uint64_t* elem_prev, const uint64_t* shifts,
                                const uint64_t* counts, int64_t* const* d_negr, int64_t* const* d_rt0, int64_t* const* d_rt1,
                                const uint64_t* d_iter, uint64_t iter_stride);
/* out[p] = x[p] op y[p] for nplanes share planes of n elements (Sh3Types.h:805-820) */
int aby3cu_share_op_batch(aby3cu_ctx* ctx, int op, int nplanes, const int64_t* const* d_x, const int64_t* const* d_y,
                          int64_t* const* d_out, size_t n);
/* out[p] = in[p]^T for nplanes row-major rows x cols planes (Sh3Types.h:822-838) */
int aby3cu_transpose_i64_batch(aby3cu_ctx* ctx, int nplanes, const int64_t* const* d_in, uint64_t rows, uint64_t cols,
                               int64_t* const* d_out);
/* C[j] += (s0[j] + s1[j] + s2[j]) >> shift  (Sh3Evaluator.cpp:712-718) for njobs parties */
int aby3cu_trunc_finish_batch(aby3cu_ctx* ctx, int njobs, const int64_t* const* d_s0, const int64_t* const* d_s1,
                              const int64_t* const* d_s2, int64_t* const* d_C, size_t n, uint64_t shift);
/* the cross term with a one-column right operand, C[j] (+)= A0[j] (B0[j] + B1[j]) + A1[j] B0[j], A M x K
 * (Sh3Evaluator.cpp:662-665 at the shapes of aby3-ML/Regression.h:157,166) */
int aby3cu_gemv_cross_batch(aby3cu_ctx* ctx, int njobs, const int64_t* const* d_A0, const int64_t* const* d_A1,
                            const int64_t* const* d_B0, const int64_t* const* d_B1, uint64_t M, uint64_t K,
                            int64_t* const* d_C, int accumulate);

/* The same cross terms (Sh3Evaluator.cpp:662-665, N = 1) for the THREE parties of one product when they share a GPU, every
 * share plane of A read once: party p's second plane of A is the previous party's first plane (replicated sharing), so only the
 * three first planes are passed.  C_p (+)= A0_p (B0_p + B1_p) + A0_{p-1} B0_p.  Same words as three aby3cu_gemm_cross calls
 * on consistent sharings (logistic inference, aby3-ML/aby3ML.h:102-139: X is 2^22 x 512). */
int aby3cu_gemv_ring(aby3cu_ctx* ctx, const int64_t* const* d_A0, const int64_t* const* d_B0, const int64_t* const* d_B1,
                     uint64_t M, uint64_t K, int64_t* const* d_C, int accumulate);

/* gather rows: out[r,:] = in[idx[r],:]  (extractBatch, aby3-ML/Regression.h:43-58) */
int aby3cu_gather_rows(aby3cu_ctx* ctx, const int64_t* d_in, uint64_t cols, const uint64_t* d_idx,
                       uint64_t nrows, int64_t* d_out);

/* out[i] = start + step * i: index vectors of the merge network (Sort.cpp:366-371) */
int aby3cu_iota_u64(aby3cu_ctx* ctx, uint64_t start, uint64_t step, uint64_t* d_out, size_t n);
/* scatter rows: out[idx[r],:] = in[r,:]  (compare-exchange write-back of aby3-Basic's
 * odd_even_merge, aby3-Basic/Sort.cpp:388-393) */
int aby3cu_scatter_rows(aby3cu_ctx* ctx, const int64_t* d_in, uint64_t cols, const uint64_t* d_idx,
                        uint64_t nrows, int64_t* d_out);

/* One compare-exchange stage of odd_even_merge (aby3-Basic/Sort.cpp:366-393) on both share planes in one pass:
 * gather  x[i] = src[r0 + 2 i],  y[i] = src[r0 + d + 2 i]   (i < m; the index vectors of :366-371 are these progressions),
 * scatter dst[r0 + 2 i] = x[i],  dst[r0 + d + 2 i] = y[i]    (:388-393).  d must be odd (it is q - 1 or 1 in the network): the
 * operands are then the two parities of one contiguous range, which is read / written once. */
int aby3cu_cmpx_gather(aby3cu_ctx* ctx, const int64_t* d_src0, const int64_t* d_src1, uint64_t r0, uint64_t d, uint64_t m,
                       int64_t* d_x0, int64_t* d_x1, int64_t* d_y0, int64_t* d_y1);
int aby3cu_cmpx_scatter(aby3cu_ctx* ctx, const int64_t* d_x0, const int64_t* d_x1, const int64_t* d_y0, const int64_t* d_y1,
                        uint64_t r0, uint64_t d, uint64_t m, int64_t* d_dst0, int64_t* d_dst1);

/* ---- binary engine: Sh3BinaryEvaluator ----------------------------------------- */
/* row stride (bytes) of the bit-sliced wire memory for `width` instances
 * (mMem.reset(width, wires, 8) with 256-bit blocks, Sh3BinaryEvaluator.cpp:84). */
uint64_t aby3cu_bin_row_bytes(uint64_t width);
/* oc::transpose (Sh3BinaryEvaluator.cpp:252,1365): out bit (r,c) = in bit (c,r),
 * LSB first.  in: rows x cols bits, in_stride bytes per row; out: cols x rows
 * bits, out_stride bytes per row; strides multiples of 4.  Bits of each written
 * out row beyond `rows` (up to the next 32-bit word) are zeroed; if invert_rows
 * (device, one byte per OUT... see below) is non-NULL it flags IN rows to complement
 * (getOutput's isInvert, :1329-1335) -- indexed by in row. */
int aby3cu_bit_transpose(aby3cu_ctx* ctx, const void* d_in, uint64_t rows, uint64_t cols, uint64_t in_stride,
                         void* d_out, uint64_t out_stride, const uint8_t* d_invert_rows);
/* transpose where IN rows are gathered through row_index (wire ids): in row r is
 * at d_in + row_index[r]*in_stride (getOutput gathers output wires, :1300-1327) */
int aby3cu_bit_transpose_gather(aby3cu_ctx* ctx, const void* d_in, const uint32_t* d_row_index, uint64_t rows,
                                uint64_t cols, uint64_t in_stride, void* d_out, uint64_t out_stride,
                                const uint8_t* d_invert_rows);
/* One AND-depth level of roundCallback (Sh3BinaryEvaluator.cpp:671-1080).
 * d_gates: [n][4] u32 = in0, in1, out, type (cryptoTools GateType encoding:
 * Nor=1, na_And=4, Xor=6, And=8, Nxor=9, a(copy)=10, Or=14), executed in order.
 * mem0/mem1: the two share planes, wires x row_bytes.  Nonlinear gates use
 * z = KS_prev ^ KS_next starting at AES block  and_index0*(row_bytes/16)
 * (getShares, :1406-1442), and_index0 = number of nonlinear gates before this level.
 * d_mem1 == NULL (no keys, LINEAR gates only): plane 0 alone is evaluated -- the second plane of every wire is the previous
 * party's first plane, and a party that can read it there (same GPU) need not recompute it. */
int aby3cu_bin_level(aby3cu_ctx* ctx, const uint32_t* d_gates, uint32_t n_gates,
                     void* d_mem0, void* d_mem1, uint64_t row_bytes,
                     const uint8_t key_prev[16], const uint8_t key_next[16], uint64_t and_index0);
/* The LINEAR gates (Xor = 6, Nxor = 9, copy = 10) of one level on plane 0 only, in batches of mutually independent gates:
 * d_batch_first[g] != 0 marks the first gate of a batch (at most 8 gates are taken together); no gate of a batch may read or
 * write a wire that another gate of the same batch writes.  All operand rows of a batch are loaded before its outputs are
 * stored -- the same values as aby3cu_bin_level(..., d_mem1 = NULL) walking the list in order (roundCallback's gate loop,
 * Sh3BinaryEvaluator.cpp:671-1080, restricted to the linear types). */
int aby3cu_bin_linear_plane0(aby3cu_ctx* ctx, const uint32_t* d_gates, uint32_t n_gates, const uint8_t* d_batch_first,
                             void* d_mem0, uint64_t row_bytes);
/* The nonlinear gates of one level when they are listed AFTER the level's linear gates (they are
 * then mutually independent): gate x instance parallel.  Gate g of the list uses the zero-share
 * block range of nonlinear gate and_index0 + g, exactly as aby3cu_bin_level would. */
int aby3cu_bin_and_layer(aby3cu_ctx* ctx, const uint32_t* d_gates, uint32_t n_gates, void* d_mem0, const void* d_mem1,
                         uint64_t row_bytes, const uint8_t key_prev[16], const uint8_t key_next[16], uint64_t and_index0);
/* A ONE-LEVEL bitwise circuit (oc::BetaLibrary::int_int_bitwiseAnd / bitwiseOr as aby3-Basic/BoolBasic.cpp:100-123,
 * 193-209 evaluate them: gate g combines bit g of every instance) on ROW-MAJOR share words: setInput's transposes
 * (:252), the gate loop (:729-798 / :912-981), getShares (:1406-1442) and getOutput's transpose (:1365) in one pass.
 * out0[j] = f(a, b)[j] ^ z[j], bit g of z[j] = bit j of the zero-share blocks of nonlinear gate and_index0 + g
 * (row_bytes = aby3cu_bin_row_bytes(n): the wire-row size that fixes the block ranges).  Same values as
 * bit_transpose + bin_and_layer + bit_transpose_gather.  d_out_copy (may be NULL) receives a second copy (the message). */
int aby3cu_bin_bitwise_rowmajor(aby3cu_ctx* ctx, uint32_t gate_type /* 8 = And, 14 = Or */, const int64_t* d_a0, const int64_t* d_a1,
                                const int64_t* d_b0, const int64_t* d_b1, int64_t* d_out0, int64_t* d_out_copy, uint64_t n, uint32_t bits,
                                uint64_t row_bytes, const uint8_t key_prev[16], const uint8_t key_next[16], uint64_t and_index0);
/* aby3-Basic/BoolBasic.cpp:275-312 (bool_cipher_max_min_split) after its comparison, one pass per party on row-major words:
 * with m = -c (the comparison bit's shares widened to 0 / -1 masks, :279-284) the reference runs int_int_bitwiseAnd(64) twice
 * over the stacked 2n-row operands, t1 = [m; m] & [A; B] under keys (key_prev1, key_next1) and t2 = NOT[m; m] & [A; B] under
 * (key_prev2, key_next2) -- NOT complements share x_1 only (:315-343): not_plane = 1 at party 1 (its plane 0), 2 at party 2
 * (its plane 1), 0 at party 0 -- and xors  min = t1[0:n] ^ t2[n:2n],  max = t1[n:2n] ^ t2[0:n].  This entry writes plane 0 of
 * min and max (the same words as two aby3cu_bin_bitwise_rowmajor runs with and_index0 = 0 + aby3cu_share_op xors); plane 1 of
 * either is the previous party's plane 0.  row_bytes = aby3cu_bin_row_bytes(2 n). */
int aby3cu_bin_maxmin_rowmajor(aby3cu_ctx* ctx, const int64_t* d_c0, const int64_t* d_c1, const int64_t* d_a0, const int64_t* d_a1,
                               const int64_t* d_b0, const int64_t* d_b1, int64_t* d_min0, int64_t* d_max0, uint64_t n, uint64_t row_bytes,
                               const uint8_t key_prev1[16], const uint8_t key_next1[16], const uint8_t key_prev2[16],
                               const uint8_t key_next2[16], uint32_t not_plane);
/* sendBuff packing (:795-796) and getOutput(sPackedBin) (:1213-1283):
 * out[j*nbytes .. ) = first nbytes of row locs[j] of mem, complemented where d_invert[j] != 0
 * (d_invert may be NULL). */
int aby3cu_bin_pack_rows(aby3cu_ctx* ctx, const void* d_mem, uint64_t row_bytes, const uint32_t* d_locs,
                         uint32_t n_locs, uint64_t nbytes, void* d_out, const uint8_t* d_invert);
/* receive scatter (:555-573): first nbytes of row locs[j] of mem = in[j*nbytes ..) */
int aby3cu_bin_scatter_rows(aby3cu_ctx* ctx, void* d_mem, uint64_t row_bytes, const uint32_t* d_locs,
                            uint32_t n_locs, uint64_t nbytes, const void* d_in);
/* Shadow evaluation (the reference's BINARY_ENGINE_DEBUG checker, Sh3BinaryEvaluator.cpp:1469-1601): given ALL THREE share
 * planes of the wire memory, re-evaluates the n_gates gates ([n][4] = in0, in1, out, type) on the reconstructed wires and
 * counts the instances whose output wire disagrees (skip[g] != 0 exempts a gate).  *d_bad_count = mismatching instance-gates,
 * *d_first_bad_gate = lowest failing gate index (0xFFFFFFFF: none). */
int aby3cu_bin_check_gates(aby3cu_ctx* ctx, const uint32_t* d_gates, const uint8_t* d_skip, uint32_t n_gates, const void* d_plane_a,
                           const void* d_plane_b, const void* d_plane_c, uint64_t row_bytes, uint64_t width,
                           uint64_t* d_bad_count, uint32_t* d_first_bad_gate);

#ifdef __cplusplus
}
#endif
#endif /* ABY3CU_H */
