// Sh3BinaryEvaluator.h -- bit-sliced evaluation of a Boolean circuit over
// replicated binary shares (aby3/sh3/Sh3BinaryEvaluator.h:16-146,
// Sh3BinaryEvaluator.cpp:66-103, 200-253, 477-491, 539-1196, 1285-1442).
// Same call sequence as the reference: setCir -> setInput* -> asyncEvaluate ->
// getOutput.  The wire memory (one row per wire, one bit per instance, two share
// planes) lives in HBM; each AND-depth level is one gate-interpreter kernel with
// in-register AES-CTR zero shares, followed by the reshare of the level's AND
// outputs to the next party.  The BINARY_ENGINE_DEBUG shadow evaluator of the
// reference is a run-time switch here (enableDebug): a device-side gate checker.
#pragma once
#include <memory>

#include "BetaCircuit.h"
#include "Sh3Runtime.h"
#include "Sh3ShareGen.h"

namespace aby3 {

class Sh3BinaryEvaluator {
public:
    oc::BetaCircuit* mCir = nullptr;
    u64 mLevel = 0;

    void setCir(oc::BetaCircuit* cir, u64 width, Sh3ShareGen& gen) {
        block p = gen.mPrevCommon.get<block>();          // Sh3BinaryEvaluator.h:98-101
        block n = gen.mNextCommon.get<block>();
        setCir(cir, width, p, n);
    }
    void setCir(oc::BetaCircuit* cir, u64 width, block prevSeed, block nextSeed);

    void setInput(u64 i, const sbMatrix& in);
    void setInput(const oc::BetaBundle& wires, const sbMatrix& in);
    // as setInput(i, in), but `in` is only READ when the evaluation is enqueued: the caller keeps it alive and unmodified
    // until asyncEvaluate(...).get() / roundCallback has run (saves the copy on the row-major path below)
    void setInputRef(u64 i, const sbMatrix& in);
    void setInput(u64 i, const sPackedBin& in);
    void setReplicatedInput(u64 i, const sbMatrix& in) { setInput(i, in); }

    Sh3Task asyncEvaluate(Sh3Task dependency);
    Sh3Task asyncEvaluate(Sh3Task dependency, oc::BetaCircuit* cir, Sh3ShareGen& gen,
                          std::vector<const sbMatrix*> inputs, std::vector<sbMatrix*> outputs);
    void roundCallback(CommPkg& comm, Sh3Task task);

    void getOutput(u64 i, sbMatrix& out, bool allowUninitialized = false);
    void getOutput(const std::vector<oc::BetaWire>& wires, sbMatrix& out, bool allowUninitialized = false);
    void getOutput(u64 i, sPackedBin& out, bool allowUninitialized = false);
    void getOutput(const std::vector<oc::BetaWire>& wires, sPackedBin& out, bool allowUninitialized = false);

    bool hasMoreRounds() const { return mLevel <= mCir->mLevelCounts.size(); }

    // The reference's BINARY_ENGINE_DEBUG shadow evaluator as a run-time switch (Sh3BinaryEvaluator.h:29-47): with two
    // extra channels to the neighbours, the parties exchange their third share plane after the last round and every gate
    // is re-evaluated on the reconstructed wire values on the device; a mismatch throws.  (Gates touching a wire that is
    // written more than once cannot be checked after the fact and are skipped.)
    bool mDebug = false;
    u64 mDebugPartyIdx = (u64)-1;
    oc::Channel mDebugPrev, mDebugNext;
    void enableDebug(u64 partyIdx, oc::Channel debugPrev, oc::Channel debugNext) {
        mDebug = true;
        mDebugPartyIdx = partyIdx;
        mDebugPrev = std::move(debugPrev);
        mDebugNext = std::move(debugNext);
    }
    // number of instance-gates that failed the last check (0 after a clean run)
    u64 mDebugMismatches = 0;
    void validateMemory();
    // the reference's debug-input exchange (Sh3BinaryEvaluator.cpp:381-425); the device-side checker exchanges whole share
    // planes after the last round instead, so there is nothing to do up front (callers: aby3-Basic/BuildingBlocks.cpp:656)
    void distributeInputs() {}

    u64 shareCount() const { return mWidth; }
    u64 rowBytes() const { return mRowBytes; }
    // raw wire memory of one share plane (tests): wires x rowBytes
    const void* planeDevice(int s) const { return s == 0 ? (const void*)mem0() : (const void*)mem1(); }

    std::array<oc::AES, 2> mShareAES;   // [0] prev key, [1] next key
    u64 mShareIdx = 0;                  // nonlinear gates evaluated so far (z counter = mShareIdx * rowBytes/16)

    // One-level bitwise circuits (int_int_bitwiseAnd / bitwiseOr) are evaluated on the row-major share words
    // (aby3cu_bin_bitwise_rowmajor): no wire memory, no transposes; identical shares.  ABY3_BIN_ROWMAJOR=0 turns it off.
    bool rowMajorPath() const { return mFast; }

    // "Shared planes" (round 2): in a replicated sharing the second plane of every wire IS the previous party's first plane
    // (linear gates keep that invariant, and an AND output's second plane is by definition the neighbour's first,
    // Sh3BinaryEvaluator.cpp:1161-1171).  When both neighbours sit on the same GPU the party therefore keeps plane 0 only:
    // linear gates are evaluated on one plane, an AND level reads its second-plane operands in the previous party's wire
    // memory (ordered by one event per level, which is all the reshare message carries then), setInput transposes one plane,
    // getOutput reads the second plane of the output wires next door.  No pack / copy / scatter of the AND rows, half the
    // linear work, half the wire memory.  Identical share words.  Opt-in per evaluation (call before setCir): the INPUT
    // sharings must be consistent (mShares[1] == previous party's mShares[0]) -- true for every sharing the protocols
    // produce; a caller that feeds unrelated planes (a kernel test) must not ask for it.  ABY3_BIN_SHARED_PLANES=0 turns it off.
    void sharePlanes(CommPkg& comm);
    bool sharedPlanes() const { return mShare; }
    // hand the neighbour's wire memory back (reader events recorded at this point of the party's stream): evaluators that
    // outlive an evaluation (Sh3Piecewise::binEng) call it after their last getOutput
    void releaseSharedPlanes() { releaseBorrows(); }
    ~Sh3BinaryEvaluator() { try { releaseBorrows(); } catch (...) {} }
    Sh3BinaryEvaluator() = default;
    Sh3BinaryEvaluator(const Sh3BinaryEvaluator&) = delete;
    Sh3BinaryEvaluator& operator=(const Sh3BinaryEvaluator&) = delete;

private:
    // ---- the row-major path ------------------------------------------------------------------------------------
    bool mFast = false, mFastTaken = false;
    u32 mFastType = 0, mFastBits = 0;
    std::array<std::array<gpu::Buffer, 2>, 2> mFastIn;          // [input][plane] copies (setInput)
    std::array<std::array<const i64*, 2>, 2> mFastPtr{};        // [input][plane] what the kernel reads
    std::array<gpu::Buffer, 2> mFastOut;                        // result planes, handed to the output matrix
    std::future<void> mFastRecv;
    static bool fastEligible(const oc::BetaCircuit& cir, u32& type, u32& bits);
    void fastSetInput(const oc::BetaBundle& wires, const sbMatrix& in, bool copy);
    void fastRound(CommPkg& comm);
    void materialize();                                          // leave the row-major path: build the wire memory
    void allocWireMemory();

    // ---- shared planes -----------------------------------------------------------------------------------------
    bool mShareWanted = false, mShare = false;
    std::shared_ptr<gpu::SharedBuffer> mMem0Shared;             // plane 0 of the wire memory, read in place by the next party
    std::vector<oc::Borrowed> mBorrows;                          // the previous party's plane 0 (one handle per level message)
    const u8* mPrevMem = nullptr;
    void exchangeReady(CommPkg& comm, u64 logicalBytes);
    void releaseBorrows();
    u8* mem0() const { return (u8*)(mShare ? mMem0Shared->ptr() : mMem[0].ptr()); }
    const u8* mem1() const {
        if (!mShare) return (const u8*)mMem[1].ptr();
        if (!mPrevMem) throw std::runtime_error("binary engine (shared planes): the second plane is read before the first exchange " LOCATION);
        return mPrevMem;
    }

    gpu::Context* mCtx = nullptr;
    u64 mWidth = 0, mRowBytes = 0;
    std::array<gpu::Buffer, 2> mMem;
    gpu::Buffer mGatesDev;              // [gates][4] u32
    gpu::Buffer mAndLocsDev;            // output wires of the nonlinear gates, in gate order
    gpu::Buffer mLinBatchDev;           // per gate: 1 = first gate of a batch of mutually independent linear gates (shared planes)
    std::vector<u64> mLevelGateOff, mLevelAndOff;
    std::vector<i64> mLevelLinear;      // leading linear gates of the level when the rest is all nonlinear, else -1
    gpu::Buffer mRecvBuf;
    std::vector<std::future<void>> mRecvFutr;
    u64 mRecvLevel = 0;
};

}  // namespace aby3
