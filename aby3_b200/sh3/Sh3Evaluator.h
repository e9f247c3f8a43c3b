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
#include <cstdlib>
#include <memory>

#include "Colocated.h"
#include "Sh3FixedPoint.h"
#include "Sh3Runtime.h"
#include "Sh3ShareGen.h"

namespace aby3 {

struct TruncationPair {
    i64Matrix mR;          // additive share of r, subtracted before the value is opened
    si64Matrix mRTrunc;    // replicated share of ~ r >> d, added back after truncation
};

// 3-party "shared OT" (aby3/OT/SharedOT.{h,cpp}): sender and helper hold the same AES key
// and block counter; the receiver learns m[choice].  All buffers are device buffers; the
// masks are generated inside the kernels (aby3cu_ot_send / _help / _recv).
class SharedOT {
public:
    void setSeed(const block& seed, u64 seedIdx = 0) { mKey = seed; mIdx = seedIdx; }
    // an OT whose key was never set would pad every message with a default key: a silent loss of privacy (SharedOT.cpp:10-11)
    void requireSeed() const { if (mIdx == ~0ull) throw RTE_LOC; }
    // masks n message pairs (d_msgs: n x 2 int64) and sends them to `recver`
    void send(oc::Channel& recver, const i64* d_msgs, u64 n) {
        requireSeed();
        gpu::Context* ctx = gpu::current();
        gpu::Buffer masked(ctx, std::max<size_t>(16 * n, 16));
        gpu::check(aby3cu_ot_send(ctx->h(), mKey.data(), mIdx, d_msgs, (i64*)masked.ptr(), n));
        mIdx += n;
        recver.asyncSendDevice(masked.ptr(), 16 * n);
    }
    // sends pad_i[choice_i] for the receiver's choice bits (bit 0 of each word of d_choice)
    void help(oc::Channel& recver, const i64* d_choice, u64 n) {
        requireSeed();
        gpu::Context* ctx = gpu::current();
        gpu::Buffer mc(ctx, std::max<size_t>(8 * n, 16));
        gpu::check(aby3cu_ot_help(ctx->h(), mKey.data(), mIdx, d_choice, (i64*)mc.ptr(), n));
        mIdx += n;
        recver.asyncSendDevice(mc.ptr(), 8 * n);
    }
    // receiver side: post both receives now, combine later with finish()
    struct AsyncRecv {
        std::shared_future<void> fMsgs, fHelp;
        std::shared_ptr<gpu::Buffer> msgs, help;
        u64 n = 0;
        // out[i] (+)= msgs[i][choice_i] ^ help[i]
        void finish(const i64* d_choice, i64* d_out, bool accumulate) const {
            fMsgs.get(); fHelp.get();
            gpu::Context* ctx = gpu::current();
            gpu::check(aby3cu_ot_recv(ctx->h(), (const i64*)msgs->ptr(), (const i64*)help->ptr(), d_choice, d_out, n, accumulate ? 1 : 0));
        }
    };
    static AsyncRecv asyncRecv(oc::Channel& sender, oc::Channel& helper, u64 n) {
        gpu::Context* ctx = gpu::current();
        AsyncRecv r;
        r.n = n;
        r.msgs = std::make_shared<gpu::Buffer>(ctx, std::max<size_t>(16 * n, 16));
        r.help = std::make_shared<gpu::Buffer>(ctx, std::max<size_t>(8 * n, 16));
        r.fMsgs = sender.asyncRecvDevice(r.msgs->ptr(), 16 * n).share();
        r.fHelp = helper.asyncRecvDevice(r.help->ptr(), 8 * n).share();
        return r;
    }
    // ---- the reference's HOST signatures (aby3/OT/SharedOT.h:13-46; callers: aby3-Basic/BuildingBlocks.cpp:335-380) ----
    // Host vectors in, host spans out; the pads are still drawn by the device kernels above.
    void send(oc::Channel& recver, span<std::array<i64, 2>> msgs) {
        gpu::Context* ctx = gpu::current();
        const u64 n = msgs.size();
        gpu::Buffer d(ctx, std::max<size_t>(16 * n, 16));
        if (n) gpu::check(aby3cu_h2d(ctx->h(), d.ptr(), msgs.data(), 16 * n));
        send(recver, (const i64*)d.ptr(), n);
    }
    void help(oc::Channel& recver, const oc::BitVector& choices) {
        gpu::Buffer d = uploadChoices(choices);
        help(recver, (const i64*)d.ptr(), choices.size());
    }
    struct HostAsyncRecv {
        AsyncRecv dev;
        std::shared_ptr<gpu::Buffer> choice;
        span<i64> out;
        void get() const {
            gpu::Context* ctx = gpu::current();
            gpu::Buffer o(ctx, std::max<size_t>(8 * dev.n, 16));
            dev.finish((const i64*)choice->ptr(), (i64*)o.ptr(), false);
            if (dev.n) gpu::check(aby3cu_d2h(ctx->h(), out.data(), o.ptr(), 8 * dev.n));
            ctx->sync();
        }
    };
    static HostAsyncRecv asyncRecv(oc::Channel& sender, oc::Channel& helper, oc::BitVector&& choices, span<i64> recvMsgs) {
        if (recvMsgs.size() != choices.size()) throw RTE_LOC;
        HostAsyncRecv r;
        r.choice = std::make_shared<gpu::Buffer>(uploadChoices(choices));
        r.out = recvMsgs;
        r.dev = asyncRecv(sender, helper, choices.size());
        return r;
    }
    static void recv(oc::Channel& sender, oc::Channel& helper, const oc::BitVector& choices, span<i64> recvMsgs) {
        oc::BitVector c = choices;
        asyncRecv(sender, helper, std::move(c), recvMsgs).get();
    }
    block mKey;
    u64 mIdx = (u64)-1;

private:
    // one int64 word per choice bit (the layout aby3cu_ot_help / _recv read)
    static gpu::Buffer uploadChoices(const oc::BitVector& choices) {
        gpu::Context* ctx = gpu::current();
        const u64 n = choices.size();
        std::vector<i64> w(n);
        for (u64 i = 0; i < n; ++i) w[i] = choices[i];
        gpu::Buffer d(ctx, std::max<size_t>(8 * n, 16));
        if (n) {
            gpu::check(aby3cu_h2d(ctx->h(), d.ptr(), w.data(), 8 * n));
            ctx->sync();                    // w dies with this scope
        }
        return d;
    }
};

class Sh3Evaluator {
public:
    void init(u64 partyIdx, block prevSeed, block nextSeed, u64 buffSize = 256);
    void init(u64 partyIdx, CommPkg& comm, block seed, u64 buffSize = 256);

    bool DEBUG_disable_randomization = false;
    // force a GEMM algorithm (ABY3CU_GEMM_*); AUTO picks tcgen05 for dense shapes
    int mGemmAlgo = ABY3CU_GEMM_AUTO;
    // truncation pairs of large matrix products are produced ahead on the party's second stream (Sh3Evaluator.cpp,
    // truncating asyncMul); ABY3_EARLY_TRUNCATION=0 in the environment keeps them on the party's own stream
    bool mEarlyTruncation = earlyTruncationDefault();
    static bool earlyTruncationDefault() {
        const char* e = std::getenv("ABY3_EARLY_TRUNCATION");
        return !(e && e[0] == '0');
    }

    // Parties on DIFFERENT GPUs: the opened xy - r of a truncating matrix product (Sh3Evaluator.cpp:681-684) leaves in
    // mOpenBlocks row blocks, each on the party's communication stream as soon as its rows are final, while the later
    // blocks still multiply.  A protocol parameter: every party must use the same value (1 = one message, the default).
    u64 mOpenBlocks = 1;

    // The three parties of this evaluator's protocol instance run on ONE GPU (set by whoever placed them, e.g. the harness
    // Session): GEMV-shaped truncating products (N = 1, at least kRingMin elements of A) meet in the group and run as one
    // launch that reads every share plane of A once (sh3/Colocated.h).  ABY3_RING_GEMV=0 turns it off.
    std::shared_ptr<gpu::ColocatedGroup> mColocated;
    static constexpr u64 kRingMin = u64(1) << 22;

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

    // bit x arithmetic products over SharedOT (Sh3Evaluator.cpp:119-263, 418-501); B holds one bit per row
    Sh3Task asyncMul(Sh3Task dep, const si64Matrix& A, const sbMatrix& B, si64Matrix& C);
    Sh3Task asyncMul(Sh3Task dep, const i64& a, const sbMatrix& B, si64Matrix& C);

    TruncationPair getTruncationTuple(u64 xSize, u64 ySize, u64 d);

    u64 mPartyIdx = (u64)-1, mTruncationIdx = 0;
    Sh3ShareGen mShareGen;
    SharedOT mOtPrevRecver;   // seed shared with the next party
    SharedOT mOtNextRecver;   // seed shared with the previous party

    // index of the next 8-byte element of a common PRNG's keystream (what the fused device kernels are addressed by)
    static u64 streamElem(const oc::PRNG& p);

private:
    enum class MulMode { Matmul, Hadamard };
    static MulMode mulMode(const si64Matrix& A, const si64Matrix& B);
};

}  // namespace aby3
