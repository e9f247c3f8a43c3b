// Basics.h -- the comparison / selection / merge building blocks of aby3-Basic that sit
// directly on the binary engine (aby3-Basic/BoolBasic.cpp:20-40, 100-123, 193-209, 228-312,
// BuildingBlocks.cpp:407-532, Sort.cpp:327-406), with the reference's calling convention
//     func(role, inputA, inputB, res, Sh3Encryptor&, Sh3Evaluator&, Sh3Runtime&).
// The reference walks the vectors element by element on the host around every engine call
// (mask expansion, concatenation, gather / scatter of the merge network); here those loops
// are device kernels, so a compare-exchange stage never touches host memory.
#pragma once
#include <cmath>
#include <cstdlib>

#include "../sh3/Sh3BinaryEvaluator.h"
#include "../sh3/Sh3Encryptor.h"
#include "../sh3/Sh3Evaluator.h"

namespace aby3 {
namespace basic {

constexpr u64 BITSIZE = 64;

namespace detail {
inline void runCircuit(oc::BetaCircuit* cir, const sbMatrix& in0, const sbMatrix& in1, sbMatrix& res, u64 outBits,
                       Sh3Evaluator& eval, Sh3Runtime& runtime) {
    Sh3BinaryEvaluator binEng;
    const u64 n = in0.rows();
    binEng.sharePlanes(runtime.mComm);   // in0 / in1 are sharings the protocols produced: plane 1 is the neighbour's plane 0
    binEng.setCir(cir, n, eval.mShareGen);
    binEng.setInputRef(0, in0);          // in0 / in1 outlive the .get() below
    binEng.setInputRef(1, in1);
    binEng.asyncEvaluate(runtime).then([&](Sh3Task) {
        res.resize(n, outBits);
        binEng.getOutput(0, res);
    }).get();
}
inline oc::BetaLibrary& library() { static thread_local oc::BetaLibrary lib; return lib; }
}  // namespace detail

// res = (A < B) as a one-bit sharing, signed comparison (BoolBasic.cpp:20-40; pinned by
// aby3_tests/BoolTest.cpp:122, 573: bool_cipher_lt(Y, X) reveals x > y)
inline void bool_cipher_lt(int, sbMatrix& A, sbMatrix& B, sbMatrix& res, Sh3Encryptor&, Sh3Evaluator& eval, Sh3Runtime& rt) {
    detail::runCircuit(detail::library().int_int_lt_ab(A.bitCount(), B.bitCount()), A, B, res, 1, eval, rt);
}
inline void bool_cipher_eq(int, sbMatrix& A, sbMatrix& B, sbMatrix& res, Sh3Encryptor&, Sh3Evaluator& eval, Sh3Runtime& rt) {
    detail::runCircuit(detail::library().int_eq(A.bitCount()), A, B, res, 1, eval, rt);
}
inline void bool_cipher_and(int, sbMatrix& A, sbMatrix& B, sbMatrix& res, Sh3Encryptor&, Sh3Evaluator& eval, Sh3Runtime& rt) {
    const u64 b = A.bitCount();
    detail::runCircuit(detail::library().int_int_bitwiseAnd(b, b, b), A, B, res, b, eval, rt);
}
inline void bool_cipher_or(int, sbMatrix& A, sbMatrix& B, sbMatrix& res, Sh3Encryptor&, Sh3Evaluator& eval, Sh3Runtime& rt) {
    const u64 b = A.bitCount();
    detail::runCircuit(detail::library().int_int_bitwiseOr(b, b, b), A, B, res, b, eval, rt);
}
inline void bool_cipher_add(int, sbMatrix& A, sbMatrix& B, sbMatrix& res, Sh3Encryptor&, Sh3Evaluator& eval, Sh3Runtime& rt) {
    const u64 b = A.bitCount();
    detail::runCircuit(detail::library().int_int_add(b, b, b, oc::BetaLibrary::Optimized::Depth), A, B, res, b, eval, rt);
}

// bitwise NOT of a sharing: exactly one of the three shares is complemented -- x_1, which is
// plane 0 at party 1 and plane 1 at party 2 (BoolBasic.cpp:315-343).
inline void bool_cipher_not(int pIdx, sbMatrix& A, sbMatrix& res) {
    gpu::Context* ctx = gpu::current();
    const u64 n = A.i64Size();
    if (&res != &A) res = A;
    if (pIdx == 1) gpu::check(aby3cu_axpb(ctx->h(), -1, res.mShares[0].dev(), -1, res.mShares[0].devMut(), n));
    else if (pIdx == 2) gpu::check(aby3cu_axpb(ctx->h(), -1, res.mShares[1].dev(), -1, res.mShares[1].devMut(), n));
    else if (pIdx != 0) throw std::runtime_error("bool_cipher_not: pIdx out of range.");
}

// (max, min) of two vectors of 64-bit values, element-wise (BoolBasic.cpp:275-312): one lt circuit,
// the comparison bit widened to a 0 / -1 mask share by share, two bitwiseAnd circuits over the
// stacked operands [A; B], and local xors.
inline void bool_cipher_max_min_split(int pIdx, sbMatrix& A, sbMatrix& B, sbMatrix& res_max, sbMatrix& res_min,
                                      Sh3Encryptor& enc, Sh3Evaluator& eval, Sh3Runtime& rt) {
    gpu::Context* ctx = gpu::current();
    const u64 n = A.rows();
    sbMatrix comp;
    bool_cipher_lt(pIdx, A, B, comp, enc, eval, rt);
    static const bool fused = [] { const char* e = std::getenv("ABY3_FUSED_MAXMIN"); return !(e && e[0] == '0'); }();
    if (fused && n && A.i64Cols() == 1 && B.i64Cols() == 1 && B.rows() == n && pIdx >= 0 && pIdx <= 2) {
        // Both bitwiseAnd runs, the NOT between them and the xors as one pass per party (aby3cu_bin_maxmin_rowmajor):
        // the two engine runs would draw their zero-share keys in this order (Sh3BinaryEvaluator.h:98-101), and since xor is
        // linear the second plane of min / max is the previous party's first plane -- one reshare instead of two.
        // Same share words as the path below (ABY3_FUSED_MAXMIN=0).
        const block p1 = eval.mShareGen.mPrevCommon.get<block>(), n1 = eval.mShareGen.mNextCommon.get<block>();
        const block p2 = eval.mShareGen.mPrevCommon.get<block>(), n2 = eval.mShareGen.mNextCommon.get<block>();
        res_max.resize(n, BITSIZE);
        res_min.resize(n, BITSIZE);
        gpu::check(aby3cu_bin_maxmin_rowmajor(ctx->h(), comp.mShares[0].dev(), comp.mShares[1].dev(), A.mShares[0].dev(), A.mShares[1].dev(),
                                              B.mShares[0].dev(), B.mShares[1].dev(), res_min.mShares[0].devOut(), res_max.mShares[0].devOut(), n,
                                              aby3cu_bin_row_bytes(2 * n), p1.data(), n1.data(), p2.data(), n2.data(), (u32)pIdx));
        rt.mComm.mNext.asyncSendDevice(res_min.mShares[0].dev(), n * 8);
        rt.mComm.mNext.asyncSendDevice(res_max.mShares[0].dev(), n * 8);
        auto f0 = rt.mComm.mPrev.asyncRecvDevice(res_min.mShares[1].devOut(), n * 8);
        auto f1 = rt.mComm.mPrev.asyncRecvDevice(res_max.mShares[1].devOut(), n * 8);
        f0.get();
        f1.get();
        return;
    }
    sbMatrix extComp(2 * n, BITSIZE), extAB(2 * n, BITSIZE), tmp1, tmp2;
    for (int s = 0; s < 2; ++s) {
        i64* m = extComp.mShares[s].devOut();
        // a share of the bit is 0 or 1: its mask share is 0 or -1 (:279-284)
        gpu::check(aby3cu_axpb(ctx->h(), -1, comp.mShares[s].dev(), 0, m, n));
        gpu::check(aby3cu_d2d(ctx->h(), m + n, ctx->device(), m, ctx->device(), n * 8));
        i64* ab = extAB.mShares[s].devOut();
        gpu::check(aby3cu_d2d(ctx->h(), ab, ctx->device(), A.mShares[s].dev(), ctx->device(), n * 8));
        gpu::check(aby3cu_d2d(ctx->h(), ab + n, ctx->device(), B.mShares[s].dev(), ctx->device(), n * 8));
    }
    bool_cipher_and(pIdx, extComp, extAB, tmp1, enc, eval, rt);
    bool_cipher_not(pIdx, extComp, extComp);
    bool_cipher_and(pIdx, extComp, extAB, tmp2, enc, eval, rt);
    res_max.resize(n, BITSIZE);
    res_min.resize(n, BITSIZE);
    for (int s = 0; s < 2; ++s) {
        const i64* t1 = tmp1.mShares[s].dev();
        const i64* t2 = tmp2.mShares[s].dev();
        gpu::check(aby3cu_share_op(ctx->h(), ABY3CU_OP_XOR, t1, t2 + n, res_min.mShares[s].devOut(), n));   // c ? A : B
        gpu::check(aby3cu_share_op(ctx->h(), ABY3CU_OP_XOR, t1 + n, t2, res_max.mShares[s].devOut(), n));   // c ? B : A
    }
}

inline void bool_cipher_max(int pIdx, sbMatrix& A, sbMatrix& B, sbMatrix& res, Sh3Encryptor& enc, Sh3Evaluator& eval, Sh3Runtime& rt) {
    sbMatrix mn;
    bool_cipher_max_min_split(pIdx, A, B, res, mn, enc, eval, rt);
}

// most significant bit of an arithmetic sharing (BuildingBlocks.cpp:407-448): x0+x1 enters the
// circuit from party 0 (re-shared to party 1 without a mask, as the reference does), x2 from
// parties 1/2; the circuit is the MSB of their 64-bit sum.
inline int fetch_msb(int pIdx, si64Matrix& diffAB, sbMatrix& res, Sh3Evaluator& eval, Sh3Runtime& runtime) {
    gpu::Context* ctx = gpu::current();
    const u64 n = diffAB.size();
    sbMatrix in0(n, 64), in1(n, 64);
    auto zero = [&](eMatrix<i64>& m) { gpu::check(aby3cu_memset(ctx->h(), m.devOut(), 0, n * 8)); };
    switch (pIdx) {
    case 0:
        gpu::check(aby3cu_share_op(ctx->h(), ABY3CU_OP_ADD, diffAB.mShares[0].dev(), diffAB.mShares[1].dev(), in0.mShares[0].devOut(), n));
        zero(in1.mShares[0]); zero(in1.mShares[1]);
        break;
    case 1:
        gpu::check(aby3cu_d2d(ctx->h(), in1.mShares[0].devOut(), ctx->device(), diffAB.mShares[0].dev(), ctx->device(), n * 8));
        zero(in1.mShares[1]); zero(in0.mShares[0]);
        break;
    default:
        zero(in1.mShares[0]);
        gpu::check(aby3cu_d2d(ctx->h(), in1.mShares[1].devOut(), ctx->device(), diffAB.mShares[1].dev(), ctx->device(), n * 8));
        zero(in0.mShares[0]);
    }
    runtime.mComm.mNext.asyncSendDevice(in0.mShares[0].dev(), n * 8);
    runtime.mComm.mPrev.asyncRecvDevice(in0.mShares[1].devOut(), n * 8).get();
    detail::runCircuit(detail::library().int_comp_helper(64), in0, in1, res, 1, eval, runtime);
    return 0;
}

// res = (A > B) = MSB(B - A)   (BuildingBlocks.cpp:525-532)
inline int cipher_gt(int pIdx, si64Matrix& A, si64Matrix& B, sbMatrix& res, Sh3Evaluator& eval, Sh3Runtime& runtime) {
    si64Matrix diffAB = B - A;
    return fetch_msb(pIdx, diffAB, res, eval, runtime);
}

// Merge of two sorted share vectors by the reference's compare-exchange network (Sort.cpp:327-406):
// interleave, pad with the larger last element, then stages (d = 1, q-1, q/2-1, ..., 1) of
// compare-exchanges between positions i and i+d.
inline int odd_even_merge(sbMatrix& data1, sbMatrix& data2, sbMatrix& res, int pIdx, Sh3Encryptor& enc, Sh3Evaluator& eval, Sh3Runtime& rt) {
    gpu::Context* ctx = gpu::current();
    const u64 l1 = data1.rows(), l2 = data2.rows(), length = std::max(l1, l2);
    sbMatrix result(length * 2, BITSIZE);
    // pad value: max of the two last (largest) elements (:336-345)
    sbMatrix max1(1, BITSIZE), max2(1, BITSIZE), maxEle;
    for (int s = 0; s < 2; ++s) {
        gpu::check(aby3cu_d2d(ctx->h(), max1.mShares[s].devOut(), ctx->device(), data1.mShares[s].dev() + (l1 - 1), ctx->device(), 8));
        gpu::check(aby3cu_d2d(ctx->h(), max2.mShares[s].devOut(), ctx->device(), data2.mShares[s].dev() + (l2 - 1), ctx->device(), 8));
    }
    bool_cipher_max(pIdx, max1, max2, maxEle, enc, eval, rt);
    // index vectors are arithmetic progressions: built on the device
    auto iota = [&](u64 start, u64 step, u64 n) {
        gpu::Buffer b(ctx, std::max<size_t>(n * 8, 16));
        gpu::check(aby3cu_iota_u64(ctx->h(), start, step, (u64*)b.ptr(), n));
        return b;
    };
    gpu::Buffer dEven = iota(0, 2, l1), dOdd = iota(1, 2, l2);
    for (int s = 0; s < 2; ++s) {
        const i64 pad = maxEle.mShares[s](0, 0);
        i64* r = result.mShares[s].devOut();
        gpu::check(aby3cu_axpb(ctx->h(), 0, nullptr, pad, r, 2 * length));
        gpu::check(aby3cu_scatter_rows(ctx->h(), data1.mShares[s].dev(), 1, (const u64*)dEven.ptr(), l1, r));
        gpu::check(aby3cu_scatter_rows(ctx->h(), data2.mShares[s].dev(), 1, (const u64*)dOdd.ptr(), l2, r));
    }
    u64 t = (u64)std::ceil(std::log2((double)length) + 1);
    u64 q = (u64)1 << (t - 1), d = 1, r0 = 0;
    while (d > 0) {
        // compare positions i and i + d for i = r0, r0 + 2, ... < 2*length - d   (:366-371)
        const u64 m = (length * 2 > r0 + d) ? (length * 2 - d - r0 + 1) / 2 : 0;
        if (m && (d & 1)) {
            // d is q - 1 or 1, i.e. odd: X and Y are the two parities of ONE contiguous range of `result`, read and
            // written once for both planes (aby3cu_cmpx_gather / _scatter) instead of eight indexed passes
            sbMatrix X(m, BITSIZE), Y(m, BITSIZE), mx, mn;
            gpu::check(aby3cu_cmpx_gather(ctx->h(), result.mShares[0].dev(), result.mShares[1].dev(), r0, d, m, X.mShares[0].devOut(),
                                          X.mShares[1].devOut(), Y.mShares[0].devOut(), Y.mShares[1].devOut()));
            bool_cipher_max_min_split(pIdx, X, Y, mx, mn, enc, eval, rt);
            gpu::check(aby3cu_cmpx_scatter(ctx->h(), mn.mShares[0].dev(), mn.mShares[1].dev(), mx.mShares[0].dev(), mx.mShares[1].dev(), r0, d, m,
                                           result.mShares[0].devMut(), result.mShares[1].devMut()));
        } else if (m) {
            gpu::Buffer dX = iota(r0, 2, m), dY = iota(r0 + d, 2, m);
            sbMatrix X(m, BITSIZE), Y(m, BITSIZE), mx, mn;
            for (int s = 0; s < 2; ++s) {
                gpu::check(aby3cu_gather_rows(ctx->h(), result.mShares[s].dev(), 1, (const u64*)dX.ptr(), m, X.mShares[s].devOut()));
                gpu::check(aby3cu_gather_rows(ctx->h(), result.mShares[s].dev(), 1, (const u64*)dY.ptr(), m, Y.mShares[s].devOut()));
            }
            bool_cipher_max_min_split(pIdx, X, Y, mx, mn, enc, eval, rt);
            for (int s = 0; s < 2; ++s) {
                i64* r = result.mShares[s].devMut();
                gpu::check(aby3cu_scatter_rows(ctx->h(), mn.mShares[s].dev(), 1, (const u64*)dX.ptr(), m, r));
                gpu::check(aby3cu_scatter_rows(ctx->h(), mx.mShares[s].dev(), 1, (const u64*)dY.ptr(), m, r));
            }
        }
        d = q - 1;
        q >>= 1;
        r0 = 1;
    }
    res.resize(l1 + l2, BITSIZE);
    for (int s = 0; s < 2; ++s)
        gpu::check(aby3cu_d2d(ctx->h(), res.mShares[s].devOut(), ctx->device(), result.mShares[s].dev(), ctx->device(), (l1 + l2) * 8));
    return 0;
}

}  // namespace basic
}  // namespace aby3
