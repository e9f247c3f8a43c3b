/*
 * oracle/oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a plain CPU restatement of the
 * reference algorithm (Fannxy/aby3, aby3/sh3) for the replicated-share
 * multiplication hot path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under
 * aby3_b200/ links, imports or calls it.
 *
 * PARITY STATUS: pinned against the REFERENCE'S OWN CODE at the protocol level:
 * oracle/_ref/libaby3ref.so is the reference's sh3 sources (Sh3Runtime,
 * Sh3Encryptor, Sh3ShareGen, Sh3Evaluator, SharedOT, Sh3BinaryEvaluator,
 * Sh3Piecewise, CircuitLibrary) compiled unmodified from /root/reference by
 * oracle/Makefile (target _ref), and tests/test_ref_parity.py checks this
 * restatement against it share plane for share plane on the same seeds
 * (sharing, reveal, asyncMul, getTruncationTuple, truncating asyncMul,
 * bit x arithmetic products over SharedOT, the binary engine, scheduler order).
 * Still "parity unpinned": the third-party PRIMITIVES under that code.  The
 * reference's PRNG/AES/transpose live in osu-crypto/libOTe @
 * cf537295c47a3924c13030a9b796cee9d6ebeace (cryptoTools submodule), absent from
 * /root/reference and this image; oracle/shim/ restates them from their public
 * interface (oc::PRNG = one contiguous AES-128-CTR keystream over toBlock(ctr),
 * toBlock(hi, lo) byte order, LSB-first bit transpose) and both sides of the
 * comparison share those assumptions.  AES-128 itself is pinned by FIPS-197
 * Appendix B/C.1 and SP 800-38A F.5.1; the reference's own reconstruction-level
 * test assertions are re-expressed in tests/test_oracle_*.py.
 *
 * Conventions: party i holds (x_i, x_{i-1}); plane 0 is
 * the party's own share, plane 1 the previous party's.  All "shares" arrays
 * are int64 laid out [party(3)][plane(2)][n] row-major.
 */
#ifndef ABY3_ORACLE_H
#define ABY3_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- AES-128 / AES-CTR (cryptoTools oc::AES restatement) ---------------- */
/* force_soft=1 uses the portable table-free software path; 0 uses AES-NI when
 * compiled in.  Both must agree (tests check it). */
void orc_aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16], int force_soft);
/* oc::AES::ecbEncCounterMode(baseIdx, n, out): out[i] = AES_key(toBlock(baseIdx+i)),
 * toBlock(x) = 16 bytes: little-endian x in bytes 0..7, zero in 8..15. */
void orc_aes_ctr_blocks(const uint8_t key[16], uint64_t base_idx, uint64_t nblocks, uint8_t* out);
/* bytes [byte_off, byte_off+nbytes) of the stream AES_key(0)||AES_key(1)||... */
void orc_keystream(const uint8_t key[16], uint64_t byte_off, uint64_t nbytes, uint8_t* out);

/* ---- oc::PRNG restatement (buffered, stateful) -------------------------- */
typedef struct orc_prng orc_prng;
orc_prng* orc_prng_new(const uint8_t seed[16], uint64_t buffer_blocks /* 256 */);
void orc_prng_free(orc_prng*);
void orc_prng_get(orc_prng*, uint8_t* dst, uint64_t nbytes);
uint64_t orc_prng_bytes_consumed(const orc_prng*);

/* ---- three-party session (Sh3Encryptor + Sh3Evaluator state x3) --------- */
typedef struct orc_session orc_session;
/* seeds: [party][0=prev,1=next][16].  Mirrors enc.init(i, prev, next) and
 * eval.init(i, prev, next) of aby3_tests/Sh3EvaluatorTests.cpp:41-47. */
orc_session* orc_session_new(const uint8_t enc_seeds[3][2][16], const uint8_t eval_seeds[3][2][16]);
void orc_session_free(orc_session*);
void orc_session_set_disable_randomization(orc_session*, int on);
/* cursors[0]=enc zero-share element idx, [1]=eval zero-share element idx,
 * [2]=eval prevCommon byte cursor, [3]=eval nextCommon byte cursor,
 * [4]=enc prevCommon byte cursor, [5]=enc nextCommon byte cursor */
void orc_session_cursors(const orc_session*, int party, uint64_t cursors[6]);

/* Sh3Encryptor::localIntMatrix (owner) + remoteIntMatrix (others). */
void orc_share_int(orc_session*, int owner, const int64_t* plain, int64_t* shares, uint64_t n);
/* Sh3Encryptor::localBinMatrix / remoteBinMatrix. */
void orc_share_bin(orc_session*, int owner, const int64_t* plain, int64_t* shares, uint64_t n);
/* Sh3Encryptor::reveal as seen by `party`: x[0]+x[1]+next.x[0]  (xor if bin). */
void orc_reveal(const int64_t* shares, uint64_t n, int party, int binary, int64_t* out);

/* Sh3Evaluator::asyncMul(si64Matrix) -- mode 0: matrix product (upstream form,
 * Sh3Evaluator.cpp:96-99), mode 1: Hadamard (fork form :101-105).  One zero
 * share per output element, row-major.  nthreads: 1 = sequential, 3 = one
 * thread per party (the reference's model), >3 = also split GEMM rows. */
void orc_mul(orc_session*, const int64_t* A, const int64_t* B, int64_t* C,
             uint64_t M, uint64_t K, uint64_t N, int mode, int nthreads);
/* Sh3Evaluator::asyncMul(si64Matrix, shift) -- Sh3Evaluator.cpp:651-730. */
void orc_mul_trunc(orc_session*, const int64_t* A, const int64_t* B, int64_t* C,
                   uint64_t M, uint64_t K, uint64_t N, int mode, uint64_t shift, int nthreads);
/* Sh3Evaluator::getTruncationTuple for one party (advances its cursors). */
void orc_trunc_tuple(orc_session*, int party, uint64_t n, uint64_t d,
                     int64_t* R, int64_t* RT0, int64_t* RT1);

/* Sh3Evaluator::asyncMul(si64Matrix, sbMatrix) / asyncMul(i64, sbMatrix) with SharedOT
 * (Sh3Evaluator.cpp:119-263, 418-501; aby3/OT/SharedOT.cpp).  B holds ONE bit per row. */
void orc_mul_bit(orc_session*, const int64_t* A, const int64_t* B, int64_t* C, uint64_t n);
void orc_mul_bit_pub(orc_session*, int64_t a, const int64_t* B, int64_t* C, uint64_t n);

/* local share arithmetic (Sh3Types.h:805-820): op 0 add, 1 sub, 2 xor */
void orc_share_op(const int64_t* X, const int64_t* Y, int64_t* Z, uint64_t n, int op);

/* ---- plain helpers ------------------------------------------------------- */
/* C = A*B (mode 0) or A.*B (mode 1) over Z_2^64 */
void orc_plain_mul(const int64_t* A, const int64_t* B, int64_t* C,
                   uint64_t M, uint64_t K, uint64_t N, int mode, int nthreads);
/* the single-party cross term A0*B0 + A0*B1 + A1*B0 (three products, as the
 * reference executes them) */
void orc_cross_term(const int64_t* A0, const int64_t* A1, const int64_t* B0, const int64_t* B1,
                    int64_t* C, uint64_t M, uint64_t K, uint64_t N, int mode, int nthreads);

/* ---- binary engine (Sh3BinaryEvaluator restatement) ---------------------- */
/* gate types follow cryptoTools GateType truth-table encoding */
enum { ORC_GATE_NOR = 1, ORC_GATE_NA_AND = 4, ORC_GATE_XOR = 6, ORC_GATE_AND = 8,
       ORC_GATE_NXOR = 9, ORC_GATE_COPY = 10, ORC_GATE_OR = 14 };
typedef struct {
    uint32_t wire_count;
    uint32_t gate_count;
    const uint32_t* gates;        /* [gate_count][4] = in0, in1, out, type ; already in level order */
    uint32_t level_count;
    const uint32_t* level_gates;  /* gates per level */
    uint32_t num_inputs;
    const uint32_t* input_first;  /* first wire of each input bundle (contiguous) */
    const uint32_t* input_bits;   /* bits in each bundle */
    uint32_t num_outputs;
    const uint32_t* output_off;   /* offset of each output bundle into output_wires */
    const uint32_t* output_bits;
    const uint32_t* output_wires; /* wire ids */
    const uint8_t* output_invert; /* per output wire: 1 = inverted */
} orc_circuit;
/* bit-matrix transpose, LSB-first (oc::transpose): in is rows x cols bits with
 * in_stride bytes per row; out is cols x rows bits with out_stride bytes per row */
void orc_bit_transpose(const uint8_t* in, uint64_t rows, uint64_t cols, uint64_t in_stride,
                       uint8_t* out, uint64_t out_stride);
/* row stride in bytes of the bit-sliced wire memory for `width` instances:
 * 256 * ceil(width / 2048)   (Sh3Types.h:537-543 with T=__m256i, alignment 8) */
uint64_t orc_bin_row_bytes(uint64_t width);
/* Full evaluation: setCir(cir,width,eval.mShareGen) on each party, setInput for
 * every input, all rounds with reshare, getOutput.  inputs[k] / outputs[k] are
 * sbMatrix share arrays [3][2][width * ceil(bits/64)].  If mem_dump != NULL it
 * receives the final wire memory [3][2][wire_count * row_bytes]. */
void orc_bin_eval(orc_session*, const orc_circuit* cir, uint64_t width,
                  const int64_t* const* inputs, int64_t* const* outputs, uint8_t* mem_dump);

/* ---- Sh3Converter restatement (aby3/sh3/Sh3Converter.h:24-41, Sh3Converter.cpp:63-411) ---- */
/* conv.init(rt, eval.mShareGen) on every party (draws the OT seeds from the common PRNGs) */
void orc_conv_init(orc_session*);
/* toBinaryMatrix(si64Matrix): the two inputs of the adder circuit, binary shares [3][2][n] */
void orc_conv_a2b_inputs(orc_session*, const int64_t* X, uint64_t n, int64_t* x0, int64_t* x1);
/* bitInjection (one round form): B binary shares [3][2][rows*words] -> Y arithmetic shares [3][2][rows*bits] */
void orc_conv_bit_injection(orc_session*, const int64_t* B, uint64_t rows, uint64_t words, uint64_t bits, int64_t* Y);

/* self test: FIPS-197 vectors, soft == AES-NI, PRNG == keystream.  0 = ok */
int orc_selftest(void);
int orc_has_aesni(void);

#ifdef __cplusplus
}
#endif
#endif
