// Sh3Piecewise.cpp -- see Sh3Piecewise.h.
#include "Sh3Piecewise.h"

namespace aby3 {

namespace {
// fixed-point product with a 128-bit intermediate (Sh3Piecewise.cpp:20-35)
i64 fxmul(i64 op1, i64 op2, i64 shift) {
    __int128 v = (__int128)op1 * (__int128)op2;
    return (i64)(v >> shift);
}
}  // namespace

std::vector<u8> Sh3Piecewise::getInputRegions(const i64Matrix& inputs, u64 decimal) {
    const u64 T = mThresholds.size();
    std::vector<u8> regions(inputs.rows() * (T + 1)), th(T);
    for (u64 i = 0; i < inputs.rows(); ++i) {
        const i64 in = inputs(i);
        for (u64 t = 0; t < T; ++t) th[t] = in < mThresholds[t].getFixedPoint(decimal) ? 1 : 0;
        regions[i * (T + 1)] = th[0];
        for (u64 t = 1; t < T; ++t) regions[i * (T + 1) + t] = (u8)((1 ^ th[t - 1]) * th[t]);
        regions[i * (T + 1) + T] = (u8)(1 ^ th.back());
    }
    return regions;
}

void Sh3Piecewise::eval(const i64Matrix& inputs, i64Matrix& outputs, u64 decimal, bool) {
    if (inputs.cols() != 1 || outputs.cols() != 1) throw std::runtime_error(LOCATION);
    if (outputs.size() != inputs.size()) throw std::runtime_error(LOCATION);
    if (mThresholds.size() == 0) throw std::runtime_error(LOCATION);
    if (mCoefficients.size() != mThresholds.size() + 1) throw std::runtime_error(LOCATION);
    const std::vector<u8> regions = getInputRegions(inputs, decimal);
    const u64 R = mCoefficients.size();
    for (u64 i = 0; i < inputs.rows(); ++i) {
        const i64 in = inputs(i);
        i64 out = 0;
        for (u64 t = 0; t < R; ++t) {
            i64 ft = 0, inPower = (1ll << decimal);
            for (u64 c = 0; c < mCoefficients[t].size(); ++c) {
                ft += fxmul(mCoefficients[t][c].getFixedPoint(decimal), inPower, (i64)decimal);
                inPower = fxmul(in, inPower, (i64)decimal);
            }
            out += regions[i * R + t] * ft;
        }
        outputs(i) = out;
    }
}

// Sh3Piecewise.cpp:381-516.  The secret x = x0+x1+x2 enters the circuit as two binary
// "sharings": (x0+x1) held by P0 alone (re-shared to P1, no mask -- as the reference does),
// with the public threshold subtracted at parties 0 and 1, and x2 held by P1/P2.
Sh3Task Sh3Piecewise::getInputRegions(const si64Matrix& inputs, u64 decimal, CommPkg& comm, Sh3Task& self, Sh3ShareGen& gen, bool) {
    gpu::Context* ctx = gpu::current();
    const u64 n = inputs.size(), T = mThresholds.size();
    const u64 pIdx = self.getRuntime().mPartyIdx;
    circuitInput0.resize(T);
    circuitInput1.resize(n, 64);
    for (auto& c : circuitInput0) c.resize(n, 64);
    auto zero = [&](eMatrix<i64>& m) { gpu::check(aby3cu_memset(ctx->h(), m.devOut(), 0, n * 8)); };
    auto& c0s0 = circuitInput0[0].mShares[0];
    auto& c0s1 = circuitInput0[0].mShares[1];
    switch (pIdx) {
    case 0:
        gpu::check(aby3cu_share_op(ctx->h(), ABY3CU_OP_ADD, inputs.mShares[0].dev(), inputs.mShares[1].dev(), c0s0.devOut(), n));
        zero(circuitInput1.mShares[0]);
        zero(circuitInput1.mShares[1]);
        break;
    case 1:
        circuitInput1.mShares[0] = inputs.mShares[0];
        zero(circuitInput1.mShares[1]);
        zero(c0s0);
        break;
    default:
        zero(circuitInput1.mShares[0]);
        circuitInput1.mShares[1] = inputs.mShares[1];
        zero(c0s0);
    }
    comm.mNext.asyncSendDevice(c0s0.dev(), n * 8);
    comm.mPrev.asyncRecvDevice(c0s1.devOut(), n * 8).get();

    for (u64 t = 1; t < T; ++t) circuitInput0[t] = circuitInput0[0];
    if (pIdx < 2)
        for (u64 t = 0; t < T; ++t) {
            auto& v = circuitInput0[t].mShares[pIdx];
            gpu::check(aby3cu_axpb(ctx->h(), 1, v.dev(), -mThresholds[t].getFixedPoint(decimal), v.devMut(), n));
        }

    auto cir = lib.int_Sh3Piecewise_helper(64, T);
    binEng.sharePlanes(comm);            // both circuit inputs are consistent sharings (plane 1 == the previous party's plane 0)
    binEng.setCir(cir, n, gen);
    binEng.setInput(T, circuitInput1);
    for (u64 t = 0; t < T; ++t) binEng.setInput(t, circuitInput0[t]);
    binEng.asyncEvaluate(self).then([&, n](Sh3Task) {
        for (u64 t = 0; t < mInputRegions.size(); ++t) {
            mInputRegions[t].resize(n, 1);
            binEng.getOutput(t, mInputRegions[t]);
        }
        binEng.releaseSharedPlanes();
    }, "binEval-continuation").get();
    return self.getRuntime();
}

// local affine maps c0 + c1 * x with an integer slope (Sh3Piecewise.cpp:518-567)
Sh3Task Sh3Piecewise::getFunctionValues(const si64Matrix& inputs, CommPkg&, Sh3Task self, u64 decimal, span<si64Matrix> outputs) {
    gpu::Context* ctx = gpu::current();
    i64 maxDegree = 0;
    for (auto& c : mCoefficients) maxDegree = std::max<i64>(maxDegree, (i64)c.size());
    if (maxDegree - 1 > 1) throw std::runtime_error("not implemented" LOCATION);
    const u64 pIdx = self.getRuntime().mPartyIdx, n = inputs.size();
    for (u64 c = 0; c < mCoefficients.size(); ++c) {
        if (mCoefficients[c].size() > 1) {
            if (!mCoefficients[c][1].mIsInteger) throw std::runtime_error("not implemented" LOCATION);
            const i64 constant = mCoefficients[c][0].getFixedPoint(decimal), slope = mCoefficients[c][1].getInteger();
            outputs[c].resize(inputs.rows(), inputs.cols());
            for (u64 s = 0; s < 2; ++s) {
                // the public constant is added to share 0 only: plane 0 at P0, plane 1 at P1
                const i64 add = (pIdx < 2 && s == pIdx) ? constant : 0;
                gpu::check(aby3cu_axpb(ctx->h(), slope, inputs.mShares[s].dev(), add, outputs[c].mShares[s].devOut(), n));
            }
        }
    }
    return self;
}

Sh3Task Sh3Piecewise::eval(Sh3Task dep, const si64Matrix& inputs, si64Matrix& outputs, u64 D, Sh3Evaluator& evaluator, bool print) {
    if (inputs.cols() != 1 || outputs.cols() != 1) throw std::runtime_error(LOCATION);
    if (outputs.size() != inputs.size()) throw std::runtime_error(LOCATION);
    if (mThresholds.size() == 0) throw std::runtime_error(LOCATION);
    if (mCoefficients.size() != mThresholds.size() + 1) throw std::runtime_error(LOCATION);
    gpu::Context* ctx = gpu::current();
    const u64 n = inputs.size();
    mInputRegions.resize(mCoefficients.size());
    getInputRegions(inputs, D, dep.getRuntime().mComm, dep, evaluator.mShareGen, print);
    functionOutputs.resize(mCoefficients.size());
    auto combineTask = getFunctionValues(inputs, dep.getRuntime().mComm, dep, D, span<si64Matrix>(functionOutputs.data(), functionOutputs.size()));
    for (int s = 0; s < 2; ++s) gpu::check(aby3cu_memset(ctx->h(), outputs.mShares[s].devOut(), 0, n * 8));
    for (u64 c = 0; c < mCoefficients.size(); ++c) {
        if (mCoefficients[c].empty()) continue;                       // an all-zero region costs nothing (:294)
        if (mCoefficients[c].size() > 1) {
            evaluator.asyncMul(combineTask, functionOutputs[c], mInputRegions[c], functionOutputs[c])
                .then([&outputs, this, c](Sh3Task) { outputs = outputs + functionOutputs[c]; })
                .get();
        } else {
            functionOutputs[c].resize(inputs.rows(), inputs.cols());
            const i64 k = mCoefficients[c][0].getFixedPoint(D);
            evaluator.asyncMul(combineTask, k, mInputRegions[c], functionOutputs[c])
                .then([&outputs, this, c](Sh3Task) { outputs = outputs + functionOutputs[c]; })
                .get();
        }
    }
    return dep;
}

}  // namespace aby3
