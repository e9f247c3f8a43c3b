// Sh3Evaluator.h -- arithmetic multiplication protocols
// (aby3/sh3/Sh3Evaluator.h:11-126, Sh3Evaluator.cpp:9-116, 503-730).
// Signatures, ownership (operands captured by reference, must outlive .get())
// and message pattern are the reference's; the cross term, the zero share, the
// truncation pair and the open-and-truncate step are aby3cu kernels.
//
// Shape rule for the matrix overloads: A.cols() == B.rows() is a
// matrix product (upstream semantics, what the reference's unit tests, aby3-ML
// and the tutorial expect); otherwise identical shapes multiply element-wise
// (this fork's loop at Sh3Evaluator.cpp:101-105 / :667-668, what aby3-Basic's
// n x 1 vectors expect); anything else throws.
#pragma once
#include "Sh3FixedPoint.h"
#include "Sh3Runtime.h"
#include "Sh3ShareGen.h"

namespace aby3 {

struct TruncationPair {
    i64Matrix mR;          // additive share of r, subtracted before the value is opened
    si64Matrix mRTrunc;    // replicated share of ~ r >> d, added back after truncation
};

// 3-party shared OT endpoint (aby3/OT/SharedOT.h); only its seed is part of this
// round's path (it consumes one block of each common PRNG in init), the OT
// protocol itself belongs to the bit x arithmetic multiplication.
struct SharedOT {
    void setSeed(const block& s) { mKey = s; mIdx = 0; }
    block mKey;
    u64 mIdx = 0;
};

class Sh3Evaluator {
public:
    void init(u64 partyIdx, block prevSeed, block nextSeed, u64 buffSize = 256);
    void init(u64 partyIdx, CommPkg& comm, block seed, u64 buffSize = 256);

    bool DEBUG_disable_randomization = false;
    // force a GEMM algorithm (ABY3CU_GEMM_*); AUTO picks tcgen05 for dense shapes
    int mGemmAlgo = ABY3CU_GEMM_AUTO;

    Sh3Task asyncMul(Sh3Task dependency, const si64& A, const si64& B, si64& C);
    Sh3Task asyncMul(Sh3Task dependency, const si64Matrix& A, const si64Matrix& B, si64Matrix& C);
    Sh3Task asyncMul(Sh3Task dependency, const si64Matrix& A, const si64Matrix& B, si64Matrix& C, u64 shift);
    Sh3Task asyncMul(Sh3Task dependency, const si64& A, const si64& B, si64& C, u64 shift);

    template <Decimal D>
    Sh3Task asyncMul(Sh3Task dependency, const sf64<D>& A, const sf64<D>& B, sf64<D>& C) {
        return asyncMul(dependency, A.i64Cast(), B.i64Cast(), C.i64Cast(), D);
    }
    template <Decimal D>
    Sh3Task asyncMul(Sh3Task dependency, const sf64Matrix<D>& A, const sf64Matrix<D>& B, sf64Matrix<D>& C, u64 shift) {
        return asyncMul(dependency, A.i64Cast(), B.i64Cast(), C.i64Cast(), D + shift);
    }
    template <Decimal D>
    Sh3Task asyncMul(Sh3Task dependency, const sf64Matrix<D>& A, const sf64Matrix<D>& B, sf64Matrix<D>& C) {
        return asyncMul(dependency, A.i64Cast(), B.i64Cast(), C.i64Cast(), D);
    }

    // bit x arithmetic products need SharedOT -- SURVEY 8(f) item 1, not in this round
    Sh3Task asyncMul(Sh3Task dep, const si64Matrix& A, const sbMatrix& B, si64Matrix& C);
    Sh3Task asyncMul(Sh3Task dep, const i64& a, const sbMatrix& B, si64Matrix& C);

    TruncationPair getTruncationTuple(u64 xSize, u64 ySize, u64 d);

    u64 mPartyIdx = (u64)-1, mTruncationIdx = 0;
    Sh3ShareGen mShareGen;
    SharedOT mOtPrevRecver;   // seed shared with the next party
    SharedOT mOtNextRecver;   // seed shared with the previous party

private:
    enum class MulMode { Matmul, Hadamard };
    static MulMode mulMode(const si64Matrix& A, const si64Matrix& B);
    static u64 streamElem(const oc::PRNG& p);
};

}  // namespace aby3
