// Sh3ShareGen.h -- zero-share generator (aby3/sh3/Sh3ShareGen.h:7-129).
// State is the same as the reference's: three common PRNGs and two AES-CTR
// keystreams keyed by the first block of mPrevCommon / mNextCommon.  The
// reference walks a 256-block buffer one i64 at a time; since both streams are
// plain CTR from counter 0, "the j-th share" is fully described by the element
// cursor j, which is what the device kernels take (key, first element).
#pragma once
#include "Sh3Types.h"

namespace aby3 {

struct Sh3ShareGen {
    void init(block prevSeed, block nextSeed, u64 /*buffSize*/ = 256) {
        mCommon.SetSeed(oc::toBlock(3488535245ull, 2454523ull));
        mNextCommon.SetSeed(nextSeed);
        mPrevCommon.SetSeed(prevSeed);
        mShareGen[0].setKey(mPrevCommon.get<block>());
        mShareGen[1].setKey(mNextCommon.get<block>());
        mShareElemIdx = 0;
    }
    void init(CommPkg& comm, block& seed, u64 buffSize = 256) {
        comm.mNext.asyncSendCopy(seed);
        block prevSeed;
        comm.mPrev.recv(prevSeed);
        init(prevSeed, seed, buffSize);
    }
    void validate(CommPkg& comm) {
        auto next = mNextCommon.get<block>();
        auto prev = mPrevCommon.get<block>();
        comm.mNext.send(next);
        block pp;
        comm.mPrev.recv(pp);
        if (pp != prev) throw RTE_LOC;
    }

    // index of the next zero-share element (the reference's mShareGenIdx/mShareIdx
    // pair collapsed into one running count)
    u64 mShareElemIdx = 0;
    oc::PRNG mNextCommon, mPrevCommon, mCommon;
    std::array<oc::AES, 2> mShareGen;       // [0] prev key, [1] next key

    // ---- scalar forms (host, one AES block per stream) -- :60-109 ----------
    i64 getShare() { u64 a, b; fetch(a, b); return (i64)(a - b); }
    i64 getBinaryShare() { u64 a, b; fetch(a, b); return (i64)(a ^ b); }
    si64 getRandIntShare() { u64 a, b; fetch(a, b); si64 r; r[0] = (i64)b; r[1] = (i64)a; return r; }
    sb64 getRandBinaryShare() { auto i = getRandIntShare(); return sb64(std::array<i64, 2>{{i[0], i[1]}}); }

    // ---- bulk forms (device): out[i] = addend[i] (+|^) z_(cursor+i), advances the cursor
    void getShares(gpu::Context* ctx, const i64* d_addend, i64* d_out, u64 n, bool binary) {
        gpu::check(aby3cu_zero_share(ctx->h(), mShareGen[0].key().data(), mShareGen[1].key().data(),
                                     mShareElemIdx, d_addend, d_out, n, binary ? 1 : 0));
        mShareElemIdx += n;
    }

private:
    void fetch(u64& a, u64& b) {
        u8 ka[8], kb[8];
        gpu::check(aby3cu_host_keystream(mShareGen[0].key().data(), 8 * mShareElemIdx, 8, ka));
        gpu::check(aby3cu_host_keystream(mShareGen[1].key().data(), 8 * mShareElemIdx, 8, kb));
        memcpy(&a, ka, 8); memcpy(&b, kb, 8);
        ++mShareElemIdx;
    }
};

}  // namespace aby3
