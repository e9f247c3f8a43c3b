// aby3ML.h -- the ML engine of aby3-ML on top of the sh3 facade
// (aby3-ML/aby3ML.h:11-139, aby3ML.cpp:4-17): input sharing, mul / mulTruncate,
// reveal.  The reference takes two oc::Session objects and opens its channels on
// them; here the caller hands over the CommPkg directly (in-process / NVLink
// channels, see sh3/Channel.h).  logisticFunc needs Sh3Piecewise.
#pragma once
#include "../sh3/Sh3Encryptor.h"
#include "../sh3/Sh3Evaluator.h"
#include "../sh3/Sh3Piecewise.h"

namespace aby3 {

class aby3ML {
public:
    Sh3Encryptor mEnc;
    Sh3Evaluator mEval;
    Sh3Runtime mRt;
    bool mPrint = true;

    u64 partyIdx() { return mRt.mPartyIdx; }

    // aby3ML.cpp:4-17: both generators are seeded from PRNG(seed); each sends its seed to
    // the next party and receives the previous party's.
    void init(u64 partyIdx, CommPkg& comm, block seed) {
        mRt.init(partyIdx, comm);
        oc::PRNG prng(seed);
        mEnc.init(partyIdx, comm, prng.get<block>());
        mEval.init(partyIdx, comm, prng.get<block>());
    }

    template <Decimal D>
    sf64Matrix<D> localInput(const f64Matrix<D>& val) {
        std::array<u64, 2> size{{val.rows(), val.cols()}};
        mRt.mComm.mNext.asyncSendCopy(size);
        mRt.mComm.mPrev.asyncSendCopy(size);
        sf64Matrix<D> dest(size[0], size[1]);
        mEnc.localFixedMatrix(mRt.noDependencies(), val, dest).get();
        return dest;
    }
    template <Decimal D>
    sf64Matrix<D> localInput(const eMatrix<double>& vals) {
        f64Matrix<D> v2(vals.rows(), vals.cols());
        for (u64 i = 0; i < vals.size(); ++i) v2(i) = vals(i);
        return localInput(v2);
    }
    template <Decimal D>
    sf64Matrix<D> remoteInput(u64 partyIdx) {
        std::array<u64, 2> size;
        if (partyIdx == (mRt.mPartyIdx + 1) % 3) mRt.mComm.mNext.recv(size);
        else if (partyIdx == (mRt.mPartyIdx + 2) % 3) mRt.mComm.mPrev.recv(size);
        else throw RTE_LOC;
        sf64Matrix<D> dest(size[0], size[1]);
        mEnc.remoteFixedMatrix(mRt.noDependencies(), dest).get();
        return dest;
    }
    void preprocess(u64, Decimal) {}

    template <Decimal D>
    eMatrix<double> reveal(const sf64Matrix<D>& vals) {
        f64Matrix<D> temp(vals.rows(), vals.cols());
        mEnc.revealAll(mRt.noDependencies(), vals, temp).get();
        eMatrix<double> ret(vals.rows(), vals.cols());
        for (u64 i = 0; i < ret.size(); ++i) ret(i) = static_cast<double>(temp(i));
        return ret;
    }
    template <Decimal D>
    double reveal(const sf64<D>& val) {
        f64<D> dest;
        mEnc.revealAll(mRt.noDependencies(), val, dest).get();
        return static_cast<double>(dest);
    }
    template <Decimal D>
    sf64Matrix<D> mul(const sf64Matrix<D>& left, const sf64Matrix<D>& right) {
        sf64Matrix<D> dest;
        mEval.asyncMul(mRt.noDependencies(), left, right, dest).get();
        return dest;
    }
    template <Decimal D>
    sf64Matrix<D> mulTruncate(const sf64Matrix<D>& left, const sf64Matrix<D>& right, u64 shift) {
        sf64Matrix<D> dest;
        mEval.asyncMul(mRt.noDependencies(), left, right, dest, shift).get();
        return dest;
    }

    // aby3ML.h:119-139: the piecewise approximation f(x) = 0 | x + 0.5 | 1 with cuts at -0.5 and 0.5
    Sh3Piecewise mLogistic;
    static void setLogistic(Sh3Piecewise& pw) {
        if (pw.mThresholds.size()) return;
        pw.mThresholds.resize(2);
        pw.mThresholds[0] = -0.5;
        pw.mThresholds[1] = 0.5;
        pw.mCoefficients.resize(3);
        pw.mCoefficients[1].resize(2);
        pw.mCoefficients[1][0] = 0.5;
        pw.mCoefficients[1][1] = 1;
        pw.mCoefficients[2].resize(1);
        pw.mCoefficients[2][0] = 1;
    }
    template <Decimal D>
    sf64Matrix<D> logisticFunc(const sf64Matrix<D>& Y) {
        setLogistic(mLogistic);
        sf64Matrix<D> out(Y.rows(), Y.cols());
        mLogistic.eval<D>(mRt.noDependencies(), Y, out, mEval).get();
        return out;
    }
};

}  // namespace aby3
