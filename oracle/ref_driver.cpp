// oracle/ref_driver.cpp -- C entry points over the REFERENCE's own protocol code.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  This file is ours; everything it
// calls -- aby3::Sh3Runtime, Sh3Encryptor, Sh3Evaluator, SharedOT, Sh3BinaryEvaluator,
// Sh3Piecewise, CircuitLibrary -- is compiled UNMODIFIED from /root/reference by
// oracle/Makefile (target `_ref`), against the stand-in third-party headers under
// oracle/shim/ (Eigen, Boost, function2, cryptoTools/libOTe are absent from the
// reference tree and this image; oracle/shim/README.md lists what is assumed).
// Three parties run as three threads over in-process channels, exactly as the
// reference's unit tests do (aby3_tests/Sh3EvaluatorTests.cpp:20-135).
//
// tests/test_ref_parity.py checks oracle.cpp (the restatement the GPU path is
// compared with) against this library share for share.
//
// Share arrays: int64 [party(3)][plane(2)][n], plane 0 = own share, as oracle.h.
#include <aby3/sh3/Sh3BinaryEvaluator.h>
#include <aby3/sh3/Sh3Converter.h>
#include <aby3/sh3/Sh3Encryptor.h>
#include <aby3/sh3/Sh3Evaluator.h>
#include <aby3/sh3/Sh3Piecewise.h>
#include <aby3/sh3/Sh3Runtime.h>

#include <aby3-ML/aby3ML.h>
#include <aby3-ML/LinearModelGen.h>
#include <aby3-ML/Regression.h>
#include <aby3-Basic/Basics.h>
#include <aby3-Basic/BuildingBlocks.h>
#include <aby3-Basic/Sort.h>
#include <cryptoTools/Common/CLP.h>
#include <aby3_tests/Test.h>
#include <fstream>
#include <map>

#include <atomic>
#include <chrono>
#include <thread>

int linear_main_3pc_sh(oc::CLP& cmd);        // aby3-ML/main-linear.cpp:173 (compiled unmodified)

using namespace aby3;

namespace {
thread_local std::string g_err;

struct RefParty {
    CommPkg comm;
    Sh3Runtime rt;
    Sh3Encryptor enc;
    Sh3Evaluator eval;
    Sh3Converter conv;
};
}  // namespace

struct ref_session {
    RefParty p[3];

    // run f(party) on three threads (the reference's model: one thread per party)
    int run(const std::function<void(int)>& f) {
        std::string errs[3];
        std::thread th[3];
        for (int i = 0; i < 3; ++i)
            th[i] = std::thread([&, i] {
                try { f(i); } catch (const std::exception& e) { errs[i] = e.what(); } catch (...) { errs[i] = "unknown exception"; }
            });
        for (auto& t : th) t.join();
        for (int i = 0; i < 3; ++i)
            if (!errs[i].empty()) { g_err = "party " + std::to_string(i) + ": " + errs[i]; return 1; }
        return 0;
    }
};

namespace {
block blk(const uint8_t* p) { block b; memcpy(&b, p, 16); return b; }

void loadInt(si64Matrix& m, const int64_t* shares, int party, u64 rows, u64 cols) {
    m.resize(rows, cols);
    const u64 n = rows * cols;
    for (int s = 0; s < 2; ++s) memcpy(m.mShares[s].data(), shares + (u64(party) * 2 + s) * n, n * 8);
}
void storeInt(const si64Matrix& m, int64_t* shares, int party) {
    const u64 n = m.size();
    for (int s = 0; s < 2; ++s) memcpy(shares + (u64(party) * 2 + s) * n, m.mShares[s].data(), n * 8);
}
void loadBin(sbMatrix& m, const int64_t* shares, int party, u64 rows, u64 bits) {
    m.resize(rows, bits);
    const u64 n = m.i64Size();
    for (int s = 0; s < 2; ++s) memcpy(m.mShares[s].data(), shares + (u64(party) * 2 + s) * n, n * 8);
}
void storeBin(const sbMatrix& m, int64_t* shares, int party) {
    const u64 n = m.i64Size();
    for (int s = 0; s < 2; ++s) memcpy(shares + (u64(party) * 2 + s) * n, m.mShares[s].data(), n * 8);
}
}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

// seeds: [party][0 = prev, 1 = next][16]; enc.init / eval.init as aby3_tests/Sh3EvaluatorTests.cpp:41-47
ref_session* ref_session_new(const uint8_t* enc_seeds, const uint8_t* eval_seeds) {
    try {
        std::unique_ptr<ref_session> s(new ref_session);
        auto c01 = oc::Channel::makePair(), c02 = oc::Channel::makePair(), c12 = oc::Channel::makePair();
        s->p[0].comm = CommPkg{c02.first, c01.first};          // {prev, next}
        s->p[1].comm = CommPkg{c01.second, c12.first};
        s->p[2].comm = CommPkg{c12.second, c02.second};
        for (int i = 0; i < 3; ++i) {
            RefParty& P = s->p[i];
            P.rt.init(i, P.comm);
            P.enc.init(i, blk(enc_seeds + (2 * i) * 16), blk(enc_seeds + (2 * i + 1) * 16));
            P.eval.init(i, blk(eval_seeds + (2 * i) * 16), blk(eval_seeds + (2 * i + 1) * 16));
        }
        return s.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_session_free(ref_session* s) { delete s; }
void ref_session_set_disable_randomization(ref_session* s, int on) {
    for (auto& P : s->p) P.eval.DEBUG_disable_randomization = on != 0;
}

// Sh3Encryptor::localIntMatrix at `owner`, remoteIntMatrix elsewhere
int ref_share_int(ref_session* s, int owner, const int64_t* plain, int64_t* shares, uint64_t rows, uint64_t cols) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix m(rows, cols);
        if (i == owner) {
            i64Matrix pl(rows, cols);
            memcpy(pl.data(), plain, rows * cols * 8);
            P.enc.localIntMatrix(P.comm, pl, m);
        } else P.enc.remoteIntMatrix(P.comm, m);
        storeInt(m, shares, i);
    });
}
// Sh3Encryptor::localBinMatrix / remoteBinMatrix; `words` 64-bit words per row
int ref_share_bin(ref_session* s, int owner, const int64_t* plain, int64_t* shares, uint64_t rows, uint64_t words) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        sbMatrix m(rows, words * 64);
        if (i == owner) {
            i64Matrix pl(rows, words);
            memcpy(pl.data(), plain, rows * words * 8);
            P.enc.localBinMatrix(P.comm, pl, m);
        } else P.enc.remoteBinMatrix(P.comm, m);
        storeBin(m, shares, i);
    });
}
// Sh3Encryptor::revealAll on every party; out = [3][n] (what each party reconstructs)
int ref_reveal_all(ref_session* s, const int64_t* shares, uint64_t rows, uint64_t cols, int binary, int64_t* out) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        i64Matrix dest(rows, cols);          // the reference requires a pre-sized destination (Sh3Encryptor.cpp:581-582)
        if (!binary) {
            si64Matrix m;
            loadInt(m, shares, i, rows, cols);
            P.enc.revealAll(P.comm, m, dest);
        } else {
            sbMatrix m;
            loadBin(m, shares, i, rows, cols * 64);
            P.enc.revealAll(P.comm, m, dest);
        }
        memcpy(out + u64(i) * rows * cols, dest.data(), rows * cols * 8);
    });
}

// Sh3Evaluator::asyncMul(dep, si64Matrix A, B, C).get()  (Sh3Evaluator.cpp:92-116: this fork's element-wise form)
int ref_mul(ref_session* s, const int64_t* A, const int64_t* B, int64_t* C, uint64_t rows, uint64_t cols) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix a, b, c(rows, cols);
        loadInt(a, A, i, rows, cols);
        loadInt(b, B, i, rows, cols);
        P.eval.asyncMul(P.rt.noDependencies(), a, b, c).get();
        storeInt(c, C, i);
    });
}
// Sh3Evaluator::asyncMul(dep, si64Matrix A, B, C, shift).get()  (Sh3Evaluator.cpp:651-730)
int ref_mul_trunc(ref_session* s, const int64_t* A, const int64_t* B, int64_t* C, uint64_t rows, uint64_t cols, uint64_t shift) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix a, b, c;
        loadInt(a, A, i, rows, cols);
        loadInt(b, B, i, rows, cols);
        P.eval.asyncMul(P.rt.noDependencies(), a, b, c, shift).get();
        storeInt(c, C, i);
    });
}
// Sh3Evaluator::getTruncationTuple on one party (Sh3Evaluator.cpp:503-566)
int ref_trunc_tuple(ref_session* s, int party, uint64_t rows, uint64_t cols, uint64_t d, int64_t* R, int64_t* RT0, int64_t* RT1) {
    try {
        TruncationPair t = s->p[party].eval.getTruncationTuple(rows, cols, d);
        const u64 n = rows * cols;
        memcpy(R, t.mR.data(), n * 8);
        memcpy(RT0, t.mRTrunc.mShares[0].data(), n * 8);
        memcpy(RT1, t.mRTrunc.mShares[1].data(), n * 8);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// asyncMul(dep, si64Matrix A, sbMatrix B, C) with SharedOT (Sh3Evaluator.cpp:119-263); B: one bit per row
int ref_mul_bit(ref_session* s, const int64_t* A, const int64_t* B, int64_t* C, uint64_t n) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix a, c(n, 1);
        sbMatrix b;
        loadInt(a, A, i, n, 1);
        loadBin(b, B, i, n, 1);
        P.eval.asyncMul(P.rt.noDependencies(), a, b, c).get();
        storeInt(c, C, i);
    });
}
// asyncMul(dep, i64 a, sbMatrix B, C) (Sh3Evaluator.cpp:418-501)
int ref_mul_bit_pub(ref_session* s, int64_t a, const int64_t* B, int64_t* C, uint64_t n) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix c(n, 1);
        sbMatrix b;
        loadBin(b, B, i, n, 1);
        const i64 aa = a;
        P.eval.asyncMul(P.rt.noDependencies(), aa, b, c).get();
        storeInt(c, C, i);
    });
}

// same flat layout as orc_circuit (oracle.h); gates already in level order
typedef struct {
    uint32_t wire_count, gate_count;
    const uint32_t* gates;
    uint32_t level_count;
    const uint32_t* level_gates;
    uint32_t num_inputs;
    const uint32_t* input_first;
    const uint32_t* input_bits;
    uint32_t num_outputs;
    const uint32_t* output_off;
    const uint32_t* output_bits;
    const uint32_t* output_wires;
    const uint8_t* output_invert;
} ref_circuit;

// Sh3BinaryEvaluator: setCir(cir, width, eval.mShareGen), setInput(sbMatrix) for every input,
// asyncEvaluate(rt).get(), getOutput(sbMatrix).  inputs[k]/outputs[k]: [3][2][width * ceil(bits/64)]
int ref_bin_eval(ref_session* s, const ref_circuit* fc, uint64_t width, const int64_t* const* inputs, int64_t* const* outputs) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        oc::BetaCircuit cir;
        cir.mWireCount = fc->wire_count;
        cir.mWireFlags.assign(fc->wire_count, oc::BetaWireFlag::Wire);
        for (u32 g = 0; g < fc->gate_count; ++g) {
            cir.mGates.emplace_back(fc->gates[4 * g], fc->gates[4 * g + 1], (oc::GateType)fc->gates[4 * g + 3], fc->gates[4 * g + 2]);
            if (!oc::isLinear(cir.mGates.back().mType)) ++cir.mNonlinearGateCount;
        }
        u64 pos = 0;
        for (u32 l = 0; l < fc->level_count; ++l) {
            u64 ands = 0;
            for (u64 g = pos; g < pos + fc->level_gates[l]; ++g) ands += !oc::isLinear(cir.mGates[g].mType);
            cir.mLevelCounts.push_back(fc->level_gates[l]);
            cir.mLevelAndCounts.push_back(ands);
            pos += fc->level_gates[l];
        }
        for (u32 k = 0; k < fc->num_inputs; ++k) {
            oc::BetaBundle b(fc->input_bits[k]);
            for (u32 j = 0; j < fc->input_bits[k]; ++j) b[j] = fc->input_first[k] + j;
            cir.mInputs.push_back(b);
        }
        for (u32 k = 0; k < fc->num_outputs; ++k) {
            oc::BetaBundle b(fc->output_bits[k]);
            for (u32 j = 0; j < fc->output_bits[k]; ++j) {
                b[j] = fc->output_wires[fc->output_off[k] + j];
                if (fc->output_invert && fc->output_invert[fc->output_off[k] + j]) cir.mWireFlags[b[j]] = oc::BetaWireFlag::InvWire;
            }
            cir.mOutputs.push_back(b);
        }
        Sh3BinaryEvaluator ev;
        ev.setCir(&cir, width, P.eval.mShareGen);
        std::vector<sbMatrix> in(fc->num_inputs);
        for (u32 k = 0; k < fc->num_inputs; ++k) {
            loadBin(in[k], inputs[k], i, width, fc->input_bits[k]);
            ev.setInput(k, in[k]);
        }
        ev.asyncEvaluate(P.rt.noDependencies()).get();
        for (u32 k = 0; k < fc->num_outputs; ++k) {
            sbMatrix out(width, fc->output_bits[k]);
            ev.getOutput(k, out);
            storeBin(out, outputs[k], i);
        }
    });
}

// Sh3Piecewise::eval(dep, si64Matrix in, out, D, evaluator) on an n x 1 sharing (Sh3Piecewise.cpp:184-379).
// Region r has coef_counts[r] coefficients, constant first; coefficient k is coef_int[k] if coef_is_int[k] else coef_dbl[k].
int ref_piecewise(ref_session* s, const int64_t* X, uint64_t n, const double* thresholds, int n_thresholds, const int* coef_counts,
                  const int* coef_is_int, const int64_t* coef_int, const double* coef_dbl, uint64_t D, int64_t* Y) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        Sh3Piecewise pw;
        for (int t = 0; t < n_thresholds; ++t) pw.mThresholds.emplace_back(thresholds[t]);
        pw.mCoefficients.resize(n_thresholds + 1);
        int k = 0;
        for (int r = 0; r <= n_thresholds; ++r)
            for (int c = 0; c < coef_counts[r]; ++c, ++k) {
                if (coef_is_int[k]) pw.mCoefficients[r].emplace_back((i64)coef_int[k]);
                else pw.mCoefficients[r].emplace_back(coef_dbl[k]);
            }
        si64Matrix in, out(n, 1);
        loadInt(in, X, i, n, 1);
        pw.eval(P.rt.noDependencies(), in, out, D, P.eval).get();
        storeInt(out, Y, i);
    });
}

// The plaintext piecewise evaluator the reference's own test pins (aby3_tests/Sh3PiecewiseTests.cpp:13-80)
int ref_piecewise_plain(const int64_t* x, uint64_t n, const double* thresholds, int n_thresholds, const int* coef_counts,
                        const int* coef_is_int, const int64_t* coef_int, const double* coef_dbl, uint64_t D, int64_t* y) {
    try {
        Sh3Piecewise pw;
        for (int t = 0; t < n_thresholds; ++t) pw.mThresholds.emplace_back(thresholds[t]);
        pw.mCoefficients.resize(n_thresholds + 1);
        int k = 0;
        for (int r = 0; r <= n_thresholds; ++r)
            for (int c = 0; c < coef_counts[r]; ++c, ++k) {
                if (coef_is_int[k]) pw.mCoefficients[r].emplace_back((i64)coef_int[k]);
                else pw.mCoefficients[r].emplace_back(coef_dbl[k]);
            }
        i64Matrix in(n, 1), out(n, 1);
        memcpy(in.data(), x, n * 8);
        pw.eval(in, out, D);
        memcpy(y, out.data(), n * 8);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// ---- Sh3Converter (aby3/sh3/Sh3Converter.cpp) -----------------------------------------------------
// conv.init(rt, eval.mShareGen) on every party
int ref_conv_init(ref_session* s) {
    try {
        for (auto& P : s->p) P.conv.init(P.rt, P.eval.mShareGen);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// conv.toBinaryMatrix(rt, si64Matrix in, sbMatrix dest).get(); dest left empty -> 64 bits per input word
int ref_conv_a2b(ref_session* s, const int64_t* X, uint64_t rows, uint64_t cols, int64_t* Y) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix in;
        loadInt(in, X, i, rows, cols);
        sbMatrix dest;
        P.conv.toBinaryMatrix(P.rt.noDependencies(), in, dest).get();
        P.rt.runAll();                       // the trailing state-keeping continuations (Sh3Converter.cpp:112)
        if (dest.rows() != rows || dest.i64Cols() != cols) throw std::runtime_error("unexpected a2b output shape");
        storeBin(dest, Y, i);
    });
}
// conv.bitInjection(rt, sbMatrix in (rows x bits), si64Matrix dest).get(); Y: [3][2][rows * bits]
int ref_conv_bit_injection(ref_session* s, const int64_t* B, uint64_t rows, uint64_t bits, int64_t* Y) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        sbMatrix in;
        loadBin(in, B, i, rows, bits);
        si64Matrix dest;
        P.conv.bitInjection(P.rt.noDependencies(), in, dest).get();
        storeInt(dest, Y, i);
    });
}
// conv.toPackedBin(in, pk) then conv.toBinaryMatrix(pk, out) on one share plane pair; packed: [2][bits * simd] words
int ref_conv_packed_roundtrip(const int64_t* in2, uint64_t rows, uint64_t bits, int64_t* packed2, uint64_t simd, int64_t* out2) {
    try {
        Sh3Converter conv;
        sbMatrix in(rows, bits), out;
        const u64 n = in.i64Size();
        for (int p = 0; p < 2; ++p) memcpy(in.mShares[p].data(), in2 + p * n, n * 8);
        sPackedBin pk;
        conv.toPackedBin(in, pk);
        if (pk.simdWidth() != simd || pk.bitCount() != bits) throw std::runtime_error("unexpected packed shape");
        for (int p = 0; p < 2; ++p) memcpy(packed2 + p * bits * simd, pk.mShares[p].data(), bits * simd * 8);
        conv.toBinaryMatrix(pk, out);
        for (int p = 0; p < 2; ++p) memcpy(out2 + p * n, out.mShares[p].data(), n * 8);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// Sh3Encryptor::localPackedBinary (owner) / remotePackedBinary, then revealAll(comm, sPackedBin, dest) on every party.
// plain: rows x cols words; shares: [3][2][64*cols * simd] words; revealed: [3][rows * cols]
int ref_share_reveal_packed(ref_session* s, int owner, const int64_t* plain, uint64_t rows, uint64_t cols, int64_t* shares, int64_t* revealed) {
    return s->run([&](int i) {
        RefParty& P = s->p[i];
        sPackedBin pk(rows, 64 * cols);
        if (i == owner) {
            i64Matrix pl(rows, cols);
            memcpy(pl.data(), plain, rows * cols * 8);
            P.enc.localPackedBinary(P.comm, pl, pk);
        } else P.enc.remotePackedBinary(P.comm, pk);
        const u64 n = pk.mShares[0].size();
        for (int p = 0; p < 2; ++p) memcpy(shares + (u64(i) * 2 + p) * n, pk.mShares[p].data(), n * 8);
        i64Matrix dest;
        P.enc.revealAll(P.comm, pk, dest);
        memcpy(revealed + u64(i) * rows * cols, dest.data(), rows * cols * 8);
    });
}

// Time the reference's three-party truncating product on this host (bench.py --impl reference / cpu_baseline):
// asyncMul(dep, A (M x K), B (K x N), C, shift).get() on three party threads -- the three Eigen-form products
// of Sh3Evaluator.cpp:662-665 (here: the shim's blocked loop), the fork's element-wise overwrite (:667-668,
// needs K >= max(M, N)), truncation pair, the open to parties 0/1 and the final pass.  Seconds per call.
double ref_time_mul_trunc(ref_session* s, uint64_t M, uint64_t K, uint64_t N, uint64_t shift, int reps) {
    // the fork's element-wise overwrite (:667-668) indexes A0 and B0 with the RESULT's linear index: that stays inside A
    // when N <= K and inside B (plus the 8192 zero elements of slack the Eigen stand-in keeps behind every matrix) when
    // M * N <= K * N + 8192
    if (N > K || M * N > K * N + 8192) { g_err = "ref_time_mul_trunc: the fork's element-wise pass would read outside the operands for this shape"; return -1.0; }
    std::vector<si64Matrix> a(3), b(3), c(3);
    for (int i = 0; i < 3; ++i) {
        a[i].resize(M, K); b[i].resize(K, N);
        for (int p = 0; p < 2; ++p) {
            for (u64 j = 0; j < M * K; ++j) a[i].mShares[p](j) = i64((j * 0x9E3779B97F4A7C15ull + i + p) >> 20);
            for (u64 j = 0; j < K * N; ++j) b[i].mShares[p](j) = i64((j * 0xC2B2AE3D27D4EB4Full + i + p) >> 20);
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        int rc = s->run([&](int i) {
            RefParty& P = s->p[i];
            P.eval.asyncMul(P.rt.noDependencies(), a[i], b[i], c[i], shift).get();
        });
        if (rc) return -1.0;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}


// ---- the reference's APPLICATIONS on top of its own sh3 layer (CPU baselines of BASELINE configs 3 and 5, and the
// ---- expected values of tests/test_compat.py, where the same unmodified sources run on the B200 facade) --------------

// aby3-ML/main-linear.cpp's own entry point, argv-style ("-N 1000 -D 100 -B 32 -I 50 -testN 100"): three party
// threads, loopback sessions, LinearModelGen data, SGD_Linear; it prints its own iters/s line (main-linear.cpp:147-149).
int ref_main_linear(int argc, const char* const* argv) {
    try {
        oc::CLP cmd;
        cmd.parse(argc, argv);
        return linear_main_3pc_sh(cmd);
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// aby3-ML linear regression exactly as main-linear.cpp:31-133 sets it up (aby3ML engine seeded toBlock(pIdx), D16,
// lr = 2^-10, party 0 inputs everything), with the data passed in instead of LinearModelGen's, timing SGD_Linear
// (Regression.h:112-184) alone.  x: N x F doubles, y: N doubles.  w_shares: [3][2][F] (may be null).
// Returns seconds for `iters` iterations, < 0 on error.
double ref_sgd_linear(const double* x, const double* y, uint64_t N, uint64_t F, uint64_t B, uint64_t iters, double lr, int64_t* w_shares) {
    static std::atomic<int> instance{0};
    const std::string tag = "sgd" + std::to_string(instance++);
    oc::IOService ios;
    std::string errs[3];
    double secs[3] = {0, 0, 0};
    std::thread th[3];
    for (int i = 0; i < 3; ++i)
        th[i] = std::thread([&, i] {
            try {
                const u64 next = (i + 1) % 3, prev = (i + 2) % 3;
                auto name = [&](u64 a, u64 b) { return tag + std::to_string(std::min(a, b)) + std::to_string(std::max(a, b)); };
                oc::Session epNext(ios, "127.0.0.1", 1212 + std::min<u64>(i, next), (u64)i < next ? oc::SessionMode::Server : oc::SessionMode::Client, name(i, next));
                oc::Session epPrev(ios, "127.0.0.1", 1212 + std::min<u64>(i, prev), (u64)i < prev ? oc::SessionMode::Server : oc::SessionMode::Client, name(i, prev));
                const Decimal D = D16;
                aby3ML p;
                p.mPrint = false;
                p.init(i, epPrev, epNext, oc::toBlock(i));
                sf64Matrix<D> X, Y, W;
                if (i == 0) {
                    eMatrix<double> vx(N, F), vy(N, 1), vw(F, 1);
                    memcpy(vx.data(), x, N * F * sizeof(double));
                    memcpy(vy.data(), y, N * sizeof(double));
                    vw.setZero();
                    X = p.localInput<D>(vx); Y = p.localInput<D>(vy); W = p.localInput<D>(vw);
                } else {
                    X = p.remoteInput<D>(0); Y = p.remoteInput<D>(0); W = p.remoteInput<D>(0);
                }
                RegressionParam params;
                params.mBatchSize = B; params.mIterations = iters; params.mLearningRate = lr;
                auto t0 = std::chrono::steady_clock::now();
                SGD_Linear(params, p, X, Y, W);
                secs[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (w_shares) storeInt(W.i64Cast(), w_shares, i);
            } catch (const std::exception& e) { errs[i] = e.what(); } catch (...) { errs[i] = "unknown exception"; }
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < 3; ++i)
        if (!errs[i].empty()) { g_err = "party " + std::to_string(i) + ": " + errs[i]; return -1.0; }
    return std::max(secs[0], std::max(secs[1], secs[2]));
}

// aby3-Basic on binary sharings of 64-bit values (n rows): op 0 = bool_cipher_lt (BoolBasic.cpp:20-40), 1 = bool_cipher_eq,
// 2 = bool_cipher_and, 3 = bool_cipher_or, 4 = bool_cipher_add, 5 = bool_cipher_max, 6 = bool_cipher_min.
// A, B, out: [3][2][n] (lt / eq: out holds one bit per row in bit 0).  seconds (wall, slowest party) in *secs.
int ref_basic_bool(ref_session* s, int op, const int64_t* A, const int64_t* B, uint64_t n, int64_t* out, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        RefParty& P = s->p[i];
        sbMatrix a, b, r;
        loadBin(a, A, i, n, 64);
        loadBin(b, B, i, n, 64);
        auto t0 = std::chrono::steady_clock::now();
        switch (op) {
        case 0: bool_cipher_lt(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 1: bool_cipher_eq(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 2: bool_cipher_and(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 3: bool_cipher_or(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 4: bool_cipher_add(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 5: bool_cipher_max(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 6: bool_cipher_min(i, a, b, r, P.enc, P.eval, P.rt); break;
        default: throw std::runtime_error("ref_basic_bool: unknown op");
        }
        t[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r.rows() != n) throw std::runtime_error("ref_basic_bool: unexpected result shape");
        storeBin(r, out, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
// cipher_gt on arithmetic sharings (BuildingBlocks.cpp:525-532: MSB of b - a); out: [3][2][n], bit 0
int ref_basic_cipher_gt(ref_session* s, const int64_t* A, const int64_t* B, uint64_t n, int64_t* out, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        RefParty& P = s->p[i];
        si64Matrix a, b;
        sbMatrix r;
        loadInt(a, A, i, n, 1);
        loadInt(b, B, i, n, 1);
        auto t0 = std::chrono::steady_clock::now();
        cipher_gt(i, a, b, r, P.eval, P.rt);
        t[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r.rows() != n) throw std::runtime_error("ref_basic_cipher_gt: unexpected result shape");
        storeBin(r, out, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
// bool_cipher_max_min_split (BoolBasic.cpp:275-312); mx, mn: [3][2][n]
int ref_basic_max_min_split(ref_session* s, const int64_t* A, const int64_t* B, uint64_t n, int64_t* mx, int64_t* mn, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        RefParty& P = s->p[i];
        sbMatrix a, b, hi, lo;
        loadBin(a, A, i, n, 64);
        loadBin(b, B, i, n, 64);
        auto t0 = std::chrono::steady_clock::now();
        bool_cipher_max_min_split(i, a, b, hi, lo, P.enc, P.eval, P.rt);
        t[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        storeBin(hi, mx, i);
        storeBin(lo, mn, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
// odd_even_merge of two sorted binary sharings (Sort.cpp:327-406); out: [3][2][n1 + n2]
int ref_basic_odd_even_merge(ref_session* s, const int64_t* A, uint64_t n1, const int64_t* B, uint64_t n2, int64_t* out, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        RefParty& P = s->p[i];
        sbMatrix a, b, r;
        loadBin(a, A, i, n1, 64);
        loadBin(b, B, i, n2, 64);
        auto t0 = std::chrono::steady_clock::now();
        odd_even_merge(a, b, r, i, P.enc, P.eval, P.rt);
        t[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r.rows() != n1 + n2) throw std::runtime_error("ref_basic_odd_even_merge: unexpected result shape");
        storeBin(r, out, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}


// Time aby3-ML's logistic inference step y = logisticFunc(X * W) (aby3-ML/aby3ML.h:102-139) on the reference's own code:
// asyncMul(X (rows x F), W (F x 1), shift) then Sh3Piecewise::eval with thresholds +-0.5.  Seconds per pass.
double ref_time_logistic(ref_session* s, uint64_t rows, uint64_t F, uint64_t D, int reps) {
    if (rows > F + 8192) { g_err = "ref_time_logistic: rows must stay <= F + 8192 (see ref_time_mul_trunc)"; return -1.0; }
    std::vector<si64Matrix> x(3), w(3), z(3), y(3);
    for (int i = 0; i < 3; ++i) {
        x[i].resize(rows, F); w[i].resize(F, 1);
        for (int p = 0; p < 2; ++p) {
            for (u64 j = 0; j < rows * F; ++j) x[i].mShares[p](j) = i64((j * 0x9E3779B97F4A7C15ull + i + p) >> 44);
            for (u64 j = 0; j < F; ++j) w[i].mShares[p](j) = i64((j * 0xC2B2AE3D27D4EB4Full + i + p) >> 50);
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        int rc = s->run([&](int i) {
            RefParty& P = s->p[i];
            P.eval.asyncMul(P.rt.noDependencies(), x[i], w[i], z[i], D).get();
            Sh3Piecewise pw;
            pw.mThresholds.resize(2);
            pw.mThresholds[0] = -0.5; pw.mThresholds[1] = 0.5;
            pw.mCoefficients.resize(3);
            pw.mCoefficients[1].resize(2);
            pw.mCoefficients[1][0] = 0.5; pw.mCoefficients[1][1] = 1;
            pw.mCoefficients[2].resize(1);
            pw.mCoefficients[2][0] = 1;
            y[i].resize(rows, 1);
            pw.eval(P.rt.noDependencies(), z[i], y[i], D, P.eval).get();
        });
        if (rc) return -1.0;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}

}  // extern "C"

// ---- the fork's own role tests (aby3_tests/Test.cpp, BoolTest.cpp, SortTest.cpp; frontend/main.cpp:16-54 dispatches them
// ---- by flag, Eval/dis_exec.sh starts one process per role).  Here: the named test, compiled unmodified, on three
// ---- threads of this process with "-role i"; its check_result() lines are counted from DEBUG_FILE.
namespace {
typedef int (*RoleTest)(oc::CLP&);
RoleTest find_role_test(const std::string& n) {
    static const std::map<std::string, RoleTest> t = {
        {"arith_basic_test", arith_basic_test}, {"bool_basic_test", bool_basic_test}, {"bool_basic_test2", bool_basic_test2},
        {"bool_aggregation_test", bool_aggregation_test}, {"get_first_zero_test", get_first_zero_test},
        {"share_conversion_test", share_conversion_test}, {"initialization_test", initialization_test},
        {"bc_sort_test", bc_sort_test}, {"bc_sort_corner_test", bc_sort_corner_test}, {"bc_sort_multiple_times", bc_sort_multiple_times},
        {"quick_sort_test", quick_sort_test}, {"quick_sort_with_duplicate_elements_test", quick_sort_with_duplicate_elements_test},
        {"odd_even_merge_test", odd_even_merge_test}, {"shuffle_test", shuffle_test}, {"correlation_test", correlation_test}};
    auto it = t.find(n);
    return it == t.end() ? nullptr : it->second;
}
}  // namespace

extern "C" int ref_role_test(const char* name, int* n_success, int* n_error) {
    RoleTest fn = find_role_test(name);
    if (!fn) { g_err = std::string("unknown role test ") + name; return 1; }
    { std::ofstream trunc(DEBUG_FILE, std::ios_base::trunc); }
    std::string errs[3];
    std::thread th[3];
    for (int i = 0; i < 3; ++i)
        th[i] = std::thread([&, i] {
            try {
                const std::string role = std::to_string(i);
                const char* argv[] = {"frontend", "-role", role.c_str()};
                oc::CLP cmd;
                cmd.parse(3, argv);
                fn(cmd);
                
            } catch (const std::exception& e) { errs[i] = e.what(); } catch (...) { errs[i] = "unknown exception"; }
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < 3; ++i)
        if (!errs[i].empty()) { g_err = "party " + std::to_string(i) + ": " + errs[i]; return 1; }
    int ok = 0, bad = 0;
    std::ifstream in(DEBUG_FILE);
    for (std::string line; std::getline(in, line);) {
        if (line.find("SUCCESS") != std::string::npos) ++ok;
        if (line.find("ERROR") != std::string::npos) ++bad;
    }
    if (n_success) *n_success = ok;
    if (n_error) *n_error = bad;
    return 0;
}
