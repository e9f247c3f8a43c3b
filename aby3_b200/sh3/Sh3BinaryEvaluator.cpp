// Sh3BinaryEvaluator.cpp -- see Sh3BinaryEvaluator.h.
#include "Sh3BinaryEvaluator.h"

namespace aby3 {

using oc::GateType;

// gate g: (input 0 bit g, input 1 bit g) -> output bit g, all gates the same nonlinear type, one level
bool Sh3BinaryEvaluator::fastEligible(const oc::BetaCircuit& cir, u32& type, u32& bits) {
    static const bool on = [] { const char* e = std::getenv("ABY3_BIN_ROWMAJOR"); return !(e && e[0] == '0'); }();
    if (!on || cir.mLevelCounts.size() != 1 || cir.mInputs.size() != 2 || cir.mOutputs.size() != 1) return false;
    const u64 n = cir.mInputs[0].size();
    if (!n || n > 64 || cir.mInputs[1].size() != n || cir.mOutputs[0].size() != n || cir.mGates.size() != n) return false;
    const GateType t = cir.mGates[0].mType;
    if (t != GateType::And && t != GateType::Or) return false;
    for (u64 g = 0; g < n; ++g) {
        const auto& G = cir.mGates[g];
        if (G.mType != t || G.mInput[0] != cir.mInputs[0][g] || G.mInput[1] != cir.mInputs[1][g] || G.mOutput != cir.mOutputs[0][g]) return false;
        if (cir.isInvert(G.mOutput)) return false;
    }
    type = (u32)t;
    bits = (u32)n;
    return true;
}

void Sh3BinaryEvaluator::sharePlanes(CommPkg& comm) {
    static const bool on = [] { const char* e = std::getenv("ABY3_BIN_SHARED_PLANES"); return !(e && e[0] == '0'); }();
    mShareWanted = on && comm.mPrev.colocated() && comm.mNext.colocated();
}

void Sh3BinaryEvaluator::releaseBorrows() {
    // (every kernel of ours that reads the neighbour's plane has been enqueued: the reader events are recorded now)
    for (auto& b : mBorrows) b.release(mCtx);
    mBorrows.clear();
    mPrevMem = nullptr;
}

// One message per AND level instead of the level's rows: the next party reads plane 0 in place once everything we have
// enqueued so far (the level's linear gates and all earlier levels) has run; `logicalBytes` = what the copying path moves.
void Sh3BinaryEvaluator::exchangeReady(CommPkg& comm, u64 logicalBytes) {
    comm.mNext.asyncSendDeviceShared(mMem0Shared, logicalBytes);
    oc::Borrowed b;
    comm.mPrev.asyncRecvDeviceBorrow(logicalBytes, &b).get();
    if (!b.shared) throw std::runtime_error("binary engine (shared planes): the previous party did not share its wire memory " LOCATION);
    mPrevMem = (const u8*)b.ptr;
    mBorrows.push_back(std::move(b));
}

void Sh3BinaryEvaluator::allocWireMemory() {
    const u64 planeBytes = std::max<u64>((u64)mCir->mWireCount * mRowBytes, 16);
    // Rows a gate writes are overwritten in full (every 16-byte chunk of the row) before anything reads them; only the other
    // wires -- the input bundles, until setInput fills them, and wires no gate drives -- start as zero like the reference's
    // freshly reset wire memory (Sh3BinaryEvaluator.cpp:84).  For the 64-bit comparison that is 128 of 505 rows.
    std::vector<u8> written(mCir->mWireCount, 0);
    for (auto& G : mCir->mGates) written[G.mOutput] = 1;
    auto clear = [&](void* base) {
        if (!mCir->mWireCount) { gpu::check(aby3cu_memset(mCtx->h(), base, 0, planeBytes)); return; }
        for (u64 w = 0; w < mCir->mWireCount;) {
            if (written[w]) { ++w; continue; }
            u64 e = w;
            while (e < mCir->mWireCount && !written[e]) ++e;
            gpu::check(aby3cu_memset(mCtx->h(), (u8*)base + w * mRowBytes, 0, (e - w) * mRowBytes));
            w = e;
        }
    };
    if (mShare) {
        mMem[0].free(); mMem[1].free();
        mMem0Shared = std::make_shared<gpu::SharedBuffer>(mCtx, planeBytes);
        clear(mMem0Shared->ptr());
        return;
    }
    mMem0Shared.reset();
    for (int s = 0; s < 2; ++s) mMem[s].reset(mCtx, planeBytes);
    clear(mMem[0].ptr());
    // plane 1 of an AND output arrives as ceil(width / 8) bytes (rounded to 16): the rest of its 256-byte-aligned row is
    // never written, so this plane is cleared as a whole (pad bits only, but they stay deterministic)
    gpu::check(aby3cu_memset(mCtx->h(), mMem[1].ptr(), 0, planeBytes));
}

void Sh3BinaryEvaluator::setCir(oc::BetaCircuit* cir, u64 width, block prevSeed, block nextSeed) {
    if (cir->mLevelCounts.size() == 0) cir->levelByAndDepth();        // .cpp:69-78
    mCir = cir;
    mCtx = gpu::current();
    mWidth = width;
    mRowBytes = aby3cu_bin_row_bytes(width);                           // mMem.reset(width, wires, 8)  :84
    mShareIdx = 0;
    mLevel = 0;
    mShareAES[0].setKey(prevSeed);                                     // :87-88
    mShareAES[1].setKey(nextSeed);

    releaseBorrows();
    mFast = !mDebug && width && fastEligible(*cir, mFastType, mFastBits);
    mFastTaken = false;
    mFastPtr = {};
    for (auto& in : mFastIn) for (auto& b : in) b.free();
    for (auto& b : mFastOut) b.free();
    mShare = false;
    // gate list and the per-level AND output wires, uploaded once
    std::vector<u32> flat(4 * cir->mGates.size()), locs;
    for (u64 g = 0; g < cir->mGates.size(); ++g) {
        const auto& G = cir->mGates[g];
        switch (G.mType) {
        case GateType::Xor: case GateType::And: case GateType::Nor: case GateType::Or:
        case GateType::Nxor: case GateType::a: case GateType::na_And: break;
        default: throw std::runtime_error("BinaryEngine unsupported GateType " LOCATION);   // :1066-1079
        }
        if (G.mOutput == G.mInput[0] || G.mOutput == G.mInput[1]) throw RTE_LOC;             // :684-689
        if (G.mInput[0] == G.mInput[1] && G.mType != GateType::a) throw RTE_LOC;
        flat[4 * g] = G.mInput[0]; flat[4 * g + 1] = G.mInput[1]; flat[4 * g + 2] = G.mOutput; flat[4 * g + 3] = (u32)G.mType;
        if (!oc::isLinear(G.mType)) locs.push_back(G.mOutput);
    }
    mLevelGateOff.assign(1, 0);
    mLevelAndOff.assign(1, 0);
    for (u64 l = 0; l < cir->mLevelCounts.size(); ++l) {
        mLevelGateOff.push_back(mLevelGateOff.back() + cir->mLevelCounts[l]);
        mLevelAndOff.push_back(mLevelAndOff.back() + cir->mLevelAndCounts[l]);
    }
    // The gate x column parallel AND layer assumes that no gate of a level overwrites a wire another gate of the level
    // reads or writes (every wire written once).  Circuits that reuse wires (hand-made gate lists; validateMemory already
    // reckons with them) take the sequential in-order kernel for the affected levels.
    std::vector<u32> writes(cir->mWireCount, 0);
    for (auto& G : cir->mGates) ++writes[G.mOutput];
    for (auto& in : cir->mInputs) for (auto w : in.mWires) ++writes[w];
    mLevelLinear.clear();
    for (u64 l = 0; l < cir->mLevelCounts.size(); ++l) {
        u64 g = mLevelGateOff[l];
        const u64 end = mLevelGateOff[l + 1];
        bool reused = false;
        for (u64 k = g; k < end; ++k) {
            const auto& G = cir->mGates[k];
            reused |= writes[G.mOutput] > 1 || writes[G.mInput[0]] > 1 || writes[G.mInput[1]] > 1;
        }
        if (reused) { mLevelLinear.push_back(-1); continue; }
        while (g < end && oc::isLinear(cir->mGates[g].mType)) ++g;
        const u64 nLinear = g - mLevelGateOff[l];
        while (g < end && !oc::isLinear(cir->mGates[g].mType)) ++g;
        mLevelLinear.push_back(g == end ? (i64)nLinear : -1);
    }
    // shared planes need every level in the split form (linear gates, then independent nonlinear ones): a level evaluated
    // by the in-order kernel would need the neighbour's SAME kernel to have finished -- around the ring, a deadlock
    mShare = mShareWanted && !mDebug && !mFast && width != 0;
    for (i64 nl : mLevelLinear) mShare = mShare && nl >= 0;
    if (mFast) { mMem[0].free(); mMem[1].free(); mMem0Shared.reset(); }
    else allocWireMemory();
    if (mShare) {
        // batches of mutually independent linear gates for aby3cu_bin_linear_plane0: a gate that reads or writes a wire an
        // earlier gate of the running batch writes starts a new batch (so does the ninth gate and every level)
        std::vector<u8> first(cir->mGates.size(), 1);
        static const bool batched = [] { const char* e = std::getenv("ABY3_BIN_LINEAR_BATCH"); return !(e && e[0] == '0'); }();
        if (batched)
            for (u64 l = 0; l < cir->mLevelCounts.size(); ++l) {
                const u64 g0 = mLevelGateOff[l], nLin = (u64)mLevelLinear[l];
                std::vector<u32> written;
                u64 inBatch = 0;
                for (u64 g = g0; g < g0 + nLin; ++g) {
                    const auto& G = cir->mGates[g];
                    bool dep = inBatch == 0 || inBatch == 8;
                    for (u32 w : written) dep = dep || w == G.mInput[0] || w == G.mInput[1] || w == G.mOutput;
                    if (dep) { written.clear(); inBatch = 0; }
                    first[g] = dep ? 1 : 0;
                    written.push_back(G.mOutput);
                    ++inBatch;
                }
            }
        mLinBatchDev.reset(mCtx, std::max<size_t>(first.size(), 16));
        if (!first.empty()) gpu::check(aby3cu_h2d(mCtx->h(), mLinBatchDev.ptr(), first.data(), first.size()));
    }
    mGatesDev.reset(mCtx, std::max<size_t>(flat.size() * 4, 16));
    mAndLocsDev.reset(mCtx, std::max<size_t>(locs.size() * 4, 16));
    if (!flat.empty()) gpu::check(aby3cu_h2d(mCtx->h(), mGatesDev.ptr(), flat.data(), flat.size() * 4));
    if (!locs.empty()) gpu::check(aby3cu_h2d(mCtx->h(), mAndLocsDev.ptr(), locs.data(), locs.size() * 4));
    // (h2d from pageable memory is staged before the call returns: `flat` / `locs` may die now)
}

void Sh3BinaryEvaluator::setInput(u64 i, const sbMatrix& in) {
    if (!mCir || i >= mCir->mInputs.size()) throw std::runtime_error(LOCATION);
    setInput(mCir->mInputs[i], in);
}

// transpose both share planes of `in` into the wire rows of the bundle (.cpp:200-253)
void Sh3BinaryEvaluator::setInputRef(u64 i, const sbMatrix& in) {
    if (!mCir || i >= mCir->mInputs.size()) throw std::runtime_error(LOCATION);
    if (mFast) fastSetInput(mCir->mInputs[i], in, false);
    else setInput(mCir->mInputs[i], in);
}

void Sh3BinaryEvaluator::fastSetInput(const oc::BetaBundle& inWires, const sbMatrix& in, bool copy) {
    mLevel = 0;
    if (in.bitCount() != inWires.size()) throw std::invalid_argument("input data wrong size");
    if (in.rows() != mWidth) throw std::invalid_argument("incorrect number of rows");
    const int k = inWires.front() == mCir->mInputs[0].mWires.front() ? 0 : 1;
    for (int s = 0; s < 2; ++s) {
        if (copy) {
            mFastIn[k][s].reset(mCtx, mWidth * 8);
            gpu::check(aby3cu_d2d(mCtx->h(), mFastIn[k][s].ptr(), mCtx->device(), in.mShares[s].dev(), mCtx->device(), mWidth * 8));
            mFastPtr[k][s] = (const i64*)mFastIn[k][s].ptr();
        } else {
            mFastPtr[k][s] = in.mShares[s].dev();
        }
    }
}

// leave the row-major path (a call it does not cover arrived): wire memory as the general engine expects it
void Sh3BinaryEvaluator::materialize() {
    if (!mFast) return;
    mFast = false;
    allocWireMemory();
    for (int k = 0; k < 2; ++k) {
        if (!mFastPtr[k][0]) continue;
        for (int s = 0; s < 2; ++s) {
            u8* dst = (u8*)mMem[s].ptr() + (u64)mCir->mInputs[k].mWires.front() * mRowBytes;
            gpu::check(aby3cu_bit_transpose(mCtx->h(), mFastPtr[k][s], mWidth, mFastBits, 8, dst, mRowBytes, nullptr));
        }
    }
    if (mLevel > 0 && mFastOut[0]) {           // already evaluated: the output wires too
        if (mFastRecv.valid()) mFastRecv.get();
        for (int s = 0; s < 2; ++s) {
            u8* dst = (u8*)mMem[s].ptr() + (u64)mCir->mOutputs[0].mWires.front() * mRowBytes;
            gpu::check(aby3cu_bit_transpose(mCtx->h(), mFastOut[s].ptr(), mWidth, mFastBits, 8, dst, mRowBytes, nullptr));
        }
        mLevel = mCir->mLevelCounts.size() + 1;
    }
}

// level 0 of the row-major path: the whole circuit + the reshare (:1161-1171; the message is the row-major plane)
void Sh3BinaryEvaluator::fastRound(CommPkg& comm) {
    if (!mFastPtr[0][0] || !mFastPtr[1][0]) throw std::runtime_error("binary engine: inputs missing " LOCATION);
    if (mShareIdx != 0) throw RTE_LOC;
    const u64 bytes = mWidth * 8;
    for (auto& b : mFastOut) b.reset(mCtx, bytes);
    gpu::check(aby3cu_bin_bitwise_rowmajor(mCtx->h(), mFastType, mFastPtr[0][0], mFastPtr[0][1], mFastPtr[1][0], mFastPtr[1][1],
                                           (i64*)mFastOut[0].ptr(), nullptr, mWidth, mFastBits, mRowBytes,
                                           mShareAES[0].key().data(), mShareAES[1].key().data(), mShareIdx));
    mShareIdx += mFastBits;
    comm.mNext.asyncSendDevice(mFastOut[0].ptr(), bytes);
    mFastRecv = comm.mPrev.asyncRecvDevice(mFastOut[1].ptr(), bytes);
    for (auto& in : mFastIn) for (auto& b : in) b.free();       // stream-ordered: the kernel above has been enqueued
}

void Sh3BinaryEvaluator::setInput(const oc::BetaBundle& inWires, const sbMatrix& in) {
    if (!mCir) throw std::runtime_error(LOCATION);
    if (mFast) {
        const bool whole = (inWires.mWires == mCir->mInputs[0].mWires) || (inWires.mWires == mCir->mInputs[1].mWires);
        if (whole) { fastSetInput(inWires, in, true); return; }
        materialize();
    }
    mLevel = 0;
    if (in.bitCount() != inWires.size()) throw std::invalid_argument("input data wrong size");
    if (in.rows() != mWidth) throw std::invalid_argument("incorrect number of rows");
    for (u64 k = 0; k + 1 < inWires.size(); ++k)
        if (inWires[k] + 1 != inWires[k + 1]) throw std::runtime_error("expecting contiguous input wires. " LOCATION);
    for (int s = 0; s < (mShare ? 1 : 2); ++s) {          // shared planes: plane 1 is the previous party's plane 0
        u8* dst = (s == 0 ? mem0() : (u8*)mMem[1].ptr()) + (u64)inWires.front() * mRowBytes;
        gpu::check(aby3cu_bit_transpose(mCtx->h(), in.mShares[s].dev(), mWidth, in.bitCount(), in.i64Cols() * 8,
                                        dst, mRowBytes, nullptr));
    }
}

// already bit-sliced input: copy each row into its wire (.cpp:279-309)
void Sh3BinaryEvaluator::setInput(u64 idx, const sPackedBin& in) {
    if (!mCir) throw std::runtime_error(LOCATION);
    if (idx >= mCir->mInputs.size()) throw std::invalid_argument("input index out of bounds");
    if (in.shareCount() != mWidth) throw std::runtime_error(LOCATION);
    materialize();
    const auto& wires = mCir->mInputs[idx].mWires;
    if (in.bitCount() != wires.size()) throw std::runtime_error(LOCATION);
    mLevel = 0;
    std::vector<u32> idxs(wires.begin(), wires.end());
    gpu::Buffer dIdx(mCtx, std::max<size_t>(idxs.size() * 4, 16));
    gpu::check(aby3cu_h2d(mCtx->h(), dIdx.ptr(), idxs.data(), idxs.size() * 4));
    for (int s = 0; s < (mShare ? 1 : 2); ++s)
        gpu::check(aby3cu_bin_scatter_rows(mCtx->h(), s == 0 ? (void*)mem0() : mMem[1].ptr(), mRowBytes, (const u32*)dIdx.ptr(), (u32)idxs.size(),
                                           in.simdWidth() * 8, in.mShares[s].dev()));
}

Sh3Task Sh3BinaryEvaluator::asyncEvaluate(Sh3Task dependency) {
    return dependency.then([this](CommPkg& comm, Sh3Task& self) { roundCallback(comm, self); }, "bin-eval-closure")
        .getClosure();
}

Sh3Task Sh3BinaryEvaluator::asyncEvaluate(Sh3Task dependency, oc::BetaCircuit* cir, Sh3ShareGen& gen,
                                          std::vector<const sbMatrix*> inputs, std::vector<sbMatrix*> outputs) {
    if (cir->mInputs.size() != inputs.size()) throw std::runtime_error(LOCATION);
    if (cir->mOutputs.size() != outputs.size()) throw std::runtime_error(LOCATION);
    return dependency.then([this, cir, &gen, inputs = std::move(inputs)](CommPkg&, Sh3Task& self) {
        const u64 width = inputs[0]->rows();
        setCir(cir, width, gen);
        for (u64 i = 0; i < inputs.size(); ++i) {
            if (inputs[i]->rows() != width) throw std::runtime_error(LOCATION);
            setInput(i, *inputs[i]);
        }
        self.then([this](CommPkg& comm, Sh3Task& self) { roundCallback(comm, self); });
    }).getClosure().then([this, outputs = std::move(outputs)](Sh3Task&) {
        for (u64 i = 0; i < outputs.size(); ++i) getOutput(i, *outputs[i]);
    });
}

// One AND-depth level per call (.cpp:539-1196): finish the previous level's
// receive, run this level's gates, reshare its AND outputs, re-schedule.
void Sh3BinaryEvaluator::roundCallback(CommPkg& comm, Sh3Task task) {
    const u64 levels = mCir->mLevelCounts.size();
    if (mLevel > levels) throw std::runtime_error("evaluateRound() was called but no rounds remain... " LOCATION);
    if (mFast && mDebug) materialize();
    if (mFast) {
        if (mLevel == 0) fastRound(comm);
        else if (mFastRecv.valid()) mFastRecv.get();
        mLevel++;
        if (hasMoreRounds()) {
            auto t = task.then([this](CommPkg& comm, Sh3Task& task) { roundCallback(comm, task); });
            t.name() = "callback";
        }
        return;
    }
    // each AND output row travels as ceil(width/8) bytes (.cpp:795-796), padded to 16 so that the
    // pack / scatter kernels move whole 128-bit words (the pad carries bits beyond `width` only)
    const u64 sendBytes = (((mWidth + 7) / 8) + 15) & ~15ull;

    if (mShare) {
        if (mLevel < levels) {
            const u64 g0 = mLevelGateOff[mLevel];
            const u64 a0 = mLevelAndOff[mLevel], nAnd = mLevelAndOff[mLevel + 1] - a0;
            if (a0 != mShareIdx) throw RTE_LOC;
            const u64 nLin = (u64)mLevelLinear[mLevel];
            if (nLin) gpu::check(aby3cu_bin_linear_plane0(mCtx->h(), (const u32*)mGatesDev.ptr() + 4 * g0, (u32)nLin, (const u8*)mLinBatchDev.ptr() + g0,
                                                          mem0(), mRowBytes));
            if (nAnd) {
                exchangeReady(comm, nAnd * sendBytes);
                gpu::check(aby3cu_bin_and_layer(mCtx->h(), (const u32*)mGatesDev.ptr() + 4 * (g0 + nLin), (u32)nAnd, mem0(), mem1(), mRowBytes,
                                                mShareAES[0].key().data(), mShareAES[1].key().data(), mShareIdx));
                mShareIdx += nAnd;
            }
        } else {
            exchangeReady(comm, 0);        // the neighbour's last level (and its trailing linear gates) precede getOutput
        }
        mLevel++;
        if (hasMoreRounds()) {
            auto t = task.then([this](CommPkg& comm, Sh3Task& task) { roundCallback(comm, task); });
            t.name() = "callback";
        }
        return;
    }

    if (mLevel) {                                                     // :555-573
        const u64 prev = mLevel - 1;
        const u64 nAnd = mLevelAndOff[prev + 1] - mLevelAndOff[prev];
        if (nAnd) {
            for (auto& fu : mRecvFutr) fu.get();
            mRecvFutr.clear();
            gpu::check(aby3cu_bin_scatter_rows(mCtx->h(), mMem[1].ptr(), mRowBytes,
                                               (const u32*)mAndLocsDev.ptr() + mLevelAndOff[prev], (u32)nAnd, sendBytes,
                                               mRecvBuf.ptr()));
        }
    }

    if (mLevel < levels) {
        const u64 g0 = mLevelGateOff[mLevel], ng = mLevelGateOff[mLevel + 1] - g0;
        const u64 a0 = mLevelAndOff[mLevel], nAnd = mLevelAndOff[mLevel + 1] - a0;
        if (a0 != mShareIdx) throw RTE_LOC;
        const i64 nLin = mLevelLinear[mLevel];
        if (nLin >= 0 && nAnd) {
            // linear gates column-parallel (in order), then the level's nonlinear gates gate x column parallel
            gpu::check(aby3cu_bin_level(mCtx->h(), (const u32*)mGatesDev.ptr() + 4 * g0, (u32)nLin, mMem[0].ptr(), mMem[1].ptr(),
                                        mRowBytes, nullptr, nullptr, mShareIdx));
            gpu::check(aby3cu_bin_and_layer(mCtx->h(), (const u32*)mGatesDev.ptr() + 4 * (g0 + nLin), (u32)nAnd, mMem[0].ptr(),
                                            mMem[1].ptr(), mRowBytes, mShareAES[0].key().data(), mShareAES[1].key().data(), mShareIdx));
        } else {
            gpu::check(aby3cu_bin_level(mCtx->h(), (const u32*)mGatesDev.ptr() + 4 * g0, (u32)ng, mMem[0].ptr(),
                                        mMem[1].ptr(), mRowBytes, nAnd ? mShareAES[0].key().data() : nullptr,
                                        nAnd ? mShareAES[1].key().data() : nullptr, mShareIdx));
        }
        mShareIdx += nAnd;
        if (nAnd) {                                                   // :795-796, :1161-1171
            const size_t bytes = nAnd * sendBytes;
            gpu::Buffer send(mCtx, std::max<size_t>(bytes, 16));
            gpu::check(aby3cu_bin_pack_rows(mCtx->h(), mMem[0].ptr(), mRowBytes, (const u32*)mAndLocsDev.ptr() + a0,
                                            (u32)nAnd, sendBytes, send.ptr(), nullptr));
            comm.mNext.asyncSendDevice(send.ptr(), bytes);
            mRecvBuf.reset(mCtx, std::max<size_t>(bytes, 16));
            mRecvFutr.emplace_back(comm.mPrev.asyncRecvDevice(mRecvBuf.ptr(), bytes));
        }
    }

    mLevel++;
    if (hasMoreRounds()) {
        auto t = task.then([this](CommPkg& comm, Sh3Task& task) { roundCallback(comm, task); });
        t.name() = "callback";
    } else if (mDebug) {
        validateMemory();
    }
}

// Shadow evaluation after the last round: fetch x_{i+1} (the next party's own plane) over the debug channels,
// reconstruct every wire and re-evaluate every gate on the device.
void Sh3BinaryEvaluator::validateMemory() {
    if (!mDebug) return;
    materialize();
    if (!mDebugPrev.isConnected() || !mDebugNext.isConnected()) throw std::runtime_error("enableDebug needs both debug channels " LOCATION);
    const u64 planeBytes = (u64)mCir->mWireCount * mRowBytes;
    gpu::Buffer third(mCtx, std::max<u64>(planeBytes, 16)), res(mCtx, 16);
    mDebugPrev.asyncSendDevice(mem0(), planeBytes);                     // our x_i is the previous party's x_{(i-1)+1}
    mDebugNext.recvDevice(third.ptr(), planeBytes);
    // a wire written by more than one gate holds only its last value: gates touching such wires are exempt
    std::vector<u32> writes(mCir->mWireCount, 0);
    for (auto& G : mCir->mGates) ++writes[G.mOutput];
    std::vector<u8> skip(mCir->mGates.size(), 0);
    bool anySkip = false;
    for (u64 g = 0; g < mCir->mGates.size(); ++g) {
        const auto& G = mCir->mGates[g];
        if (writes[G.mOutput] > 1 || writes[G.mInput[0]] > 1 || writes[G.mInput[1]] > 1) { skip[g] = 1; anySkip = true; }
        // an input bundle wire that is also written by a gate is overwritten as well
    }
    for (auto& in : mCir->mInputs)
        for (auto w : in.mWires)
            if (writes[w]) {
                for (u64 g = 0; g < mCir->mGates.size(); ++g) {
                    const auto& G = mCir->mGates[g];
                    if (G.mOutput == w || G.mInput[0] == w || G.mInput[1] == w) { skip[g] = 1; anySkip = true; }
                }
            }
    gpu::Buffer dSkip(mCtx, std::max<size_t>(skip.size(), 16));
    if (anySkip) gpu::check(aby3cu_h2d(mCtx->h(), dSkip.ptr(), skip.data(), skip.size()));
    gpu::check(aby3cu_bin_check_gates(mCtx->h(), (const u32*)mGatesDev.ptr(), anySkip ? (const u8*)dSkip.ptr() : nullptr,
                                      (u32)mCir->mGates.size(), mem0(), mem1(), third.ptr(), mRowBytes, mWidth,
                                      (u64*)res.ptr(), (u32*)((u8*)res.ptr() + 8)));
    u64 host[2] = {0, 0};
    gpu::check(aby3cu_d2h(mCtx->h(), host, res.ptr(), 16));
    mCtx->sync();
    mDebugMismatches = host[0];
    if (host[0]) {
        const u32 g = (u32)host[1];
        throw std::runtime_error("binary engine shadow check failed: " + std::to_string(host[0]) + " instance-gates disagree, first at gate " +
                                 std::to_string(g) + " (type " + std::to_string((u32)mCir->mGates[g].mType) + ") " LOCATION);
    }
}

void Sh3BinaryEvaluator::getOutput(u64 i, sbMatrix& out, bool allowUninitialized) {
    if (mCir->mOutputs.size() <= i) throw std::runtime_error(LOCATION);
    getOutput(mCir->mOutputs[i].mWires, out, allowUninitialized);
}

// gather the output wires (complementing inverted ones) and transpose back (.cpp:1285-1404)
void Sh3BinaryEvaluator::getOutput(const std::vector<oc::BetaWire>& outWires, sbMatrix& out, bool) {
    if (outWires.size() != out.bitCount()) throw std::runtime_error(LOCATION);
    if (mFast) {
        if (outWires == mCir->mOutputs[0].mWires && !mFastTaken && mLevel > mCir->mLevelCounts.size() && mFastOut[0]) {
            // the result planes become the output matrix: no copy
            if (mFastRecv.valid()) mFastRecv.get();
            out.resize(mWidth, out.bitCount());
            for (int s = 0; s < 2; ++s) out.mShares[s].adoptDevice(std::move(mFastOut[s]));
            mFastTaken = true;
            return;
        }
        if (mFastTaken) throw std::runtime_error("binary engine: the output of this evaluation has already been taken " LOCATION);
        materialize();
    }
    if (out.rows() != mWidth) out.resize(mWidth, out.bitCount());
    const u64 bits = outWires.size();
    std::vector<u32> idx(outWires.begin(), outWires.end());
    std::vector<u8> inv(bits, 0);
    bool anyInv = false;
    for (u64 b = 0; b < bits; ++b) {
        inv[b] = mCir->isInvert(outWires[b]) ? 1 : 0;
        anyInv |= inv[b] != 0;
    }
    gpu::Buffer dIdx(mCtx, std::max<size_t>(bits * 4, 16)), dInv(mCtx, std::max<size_t>(bits, 16));
    gpu::check(aby3cu_h2d(mCtx->h(), dIdx.ptr(), idx.data(), bits * 4));
    if (anyInv) gpu::check(aby3cu_h2d(mCtx->h(), dInv.ptr(), inv.data(), bits));
    for (int s = 0; s < 2; ++s) {
        i64* dst = out.mShares[s].devOut();
        gpu::check(aby3cu_memset(mCtx->h(), dst, 0, out.i64Size() * 8));
        gpu::check(aby3cu_bit_transpose_gather(mCtx->h(), s == 0 ? (const void*)mem0() : (const void*)mem1(), (const u32*)dIdx.ptr(), bits, mWidth, mRowBytes,
                                               dst, out.i64Cols() * 8, anyInv ? (const u8*)dInv.ptr() : nullptr));
    }
}

void Sh3BinaryEvaluator::getOutput(u64 i, sPackedBin& out, bool allowUninitialized) {
    if (mCir->mOutputs.size() <= i) throw std::runtime_error(LOCATION);
    getOutput(mCir->mOutputs[i].mWires, out, allowUninitialized);
}

// bit-sliced output: gather the wire rows, complementing inverted wires (.cpp:1213-1283)
void Sh3BinaryEvaluator::getOutput(const std::vector<oc::BetaWire>& outWires, sPackedBin& out, bool) {
    if (mFast && mFastTaken) throw std::runtime_error("binary engine: the output of this evaluation has already been taken " LOCATION);
    materialize();
    out.reset(mWidth, outWires.size());
    const u64 n = outWires.size();
    std::vector<u32> idx(outWires.begin(), outWires.end());
    std::vector<u8> inv(n, 0);
    bool anyInv = false;
    for (u64 b = 0; b < n; ++b) {
        if (outWires[b] >= mCir->mWireCount) throw RTE_LOC;
        inv[b] = mCir->isInvert(outWires[b]) ? 1 : 0;
        anyInv |= inv[b] != 0;
    }
    gpu::Buffer dIdx(mCtx, std::max<size_t>(n * 4, 16)), dInv(mCtx, std::max<size_t>(n, 16));
    gpu::check(aby3cu_h2d(mCtx->h(), dIdx.ptr(), idx.data(), n * 4));
    if (anyInv) gpu::check(aby3cu_h2d(mCtx->h(), dInv.ptr(), inv.data(), n));
    for (int s = 0; s < 2; ++s)
        gpu::check(aby3cu_bin_pack_rows(mCtx->h(), s == 0 ? (const void*)mem0() : (const void*)mem1(), mRowBytes, (const u32*)dIdx.ptr(), (u32)n, out.simdWidth() * 8,
                                        out.mShares[s].devOut(), anyInv ? (const u8*)dInv.ptr() : nullptr));
}

}  // namespace aby3
