// Sh3Converter.h -- conversions between the share representations next to the binary engine
// (aby3/sh3/Sh3Converter.h:12-57, Sh3Converter.cpp:12-411):
//   toPackedBin / toBinaryMatrix(sPackedBin)  row-major <-> bit-sliced binary shares (bit transposes)
//   toBinaryMatrix(si64Matrix)                arithmetic -> binary: x = (x0 + x2) + x1 through a 64-bit adder circuit
//   bitInjection                              binary -> arithmetic, one output element per input bit, over SharedOT
// Same names, argument meaning and draw order from the common PRNGs as the reference; all data stays in HBM.
#pragma once
#include "Sh3BinaryEvaluator.h"
#include "Sh3Evaluator.h"
#include "Sh3Runtime.h"
#include "Sh3ShareGen.h"

namespace aby3 {

class Sh3Converter {
public:
    oc::BetaLibrary mLib;
    Sh3ShareGen* mRandGen = nullptr;
    Sh3BinaryEvaluator mBin;
    SharedOT mOT12, mOT02;
    oc::BetaCircuit mCir;

    // Sh3Converter.h:24-41: the OT between parties {1,2} serves receiver 0, the one between {0,2} serves receiver 1
    void init(Sh3Runtime& rt, Sh3ShareGen& gen) {
        mRandGen = &gen;
        mOT12.mIdx = rt.mPartyIdx;
        mOT02.mIdx = rt.mPartyIdx;
        if (rt.mPartyIdx == 0) mOT02.setSeed(mRandGen->mPrevCommon.get());
        if (rt.mPartyIdx == 1) mOT12.setSeed(mRandGen->mNextCommon.get());
        if (rt.mPartyIdx == 2) {
            mOT12.setSeed(mRandGen->mPrevCommon.get());
            mOT02.setSeed(mRandGen->mNextCommon.get());
        }
    }

    void toPackedBin(const sbMatrix& in, sPackedBin& dest);
    void toBinaryMatrix(const sPackedBin& in, sbMatrix& dest);
    Sh3Task toBinaryMatrix(Sh3Task dep, const si64Matrix& in, sbMatrix& dest);
    Sh3Task bitInjection(Sh3Task dep, const sbMatrix& in, si64Matrix& dest, bool twoRounds = false);

    oc::BetaCircuit getArithToBinCircuit(u64 base, u64 bitCount);
};

}  // namespace aby3
