// Sh3Converter.cpp -- see Sh3Converter.h.
#include "Sh3Converter.h"

namespace aby3 {

// row-major rows x bitCount  ->  bit-sliced bitCount x ceil(rows/64) words (Sh3Converter.cpp:12-34)
void Sh3Converter::toPackedBin(const sbMatrix& in, sPackedBin& dest) {
    gpu::Context* ctx = gpu::current();
    dest.reset(in.rows(), in.bitCount());
    if (!dest.size()) return;
    for (int s = 0; s < 2; ++s) {
        i64* d = dest.mShares[s].devOut();
        gpu::check(aby3cu_memset(ctx->h(), d, 0, dest.size() * 8));
        gpu::check(aby3cu_bit_transpose(ctx->h(), in.mShares[s].dev(), in.rows(), in.bitCount(), in.i64Cols() * 8,
                                        d, dest.simdWidth() * 8, nullptr));
    }
}

// bit-sliced -> row-major (Sh3Converter.cpp:36-61)
void Sh3Converter::toBinaryMatrix(const sPackedBin& in, sbMatrix& dest) {
    gpu::Context* ctx = gpu::current();
    dest.resize(in.shareCount(), in.bitCount());
    if (!(in.bitCount() * in.shareCount())) return;
    for (int s = 0; s < 2; ++s) {
        i64* d = dest.mShares[s].devOut();
        gpu::check(aby3cu_memset(ctx->h(), d, 0, dest.i64Size() * 8));
        gpu::check(aby3cu_bit_transpose(ctx->h(), in.mShares[s].dev(), in.bitCount(), in.shareCount(), in.simdWidth() * 8,
                                        d, dest.i64Cols() * 8, nullptr));
    }
}

namespace {
// keep the low bitCount % 64 bits of the last word of every row (Sh3Converter.cpp:97-106)
void maskLastWord(gpu::Context* ctx, eMatrix<i64>& m, u64 bitCount) {
    if (bitCount % 64 == 0 || !m.size()) return;
    gpu::check(aby3cu_mask_last_word(ctx->h(), m.devMut(), m.rows(), m.cols(), (1ull << (bitCount % 64)) - 1));
}
}  // namespace

// Sh3Converter.cpp:63-209.  x0 = binary sharing of (in.0 + in.2): party 0 forms the sum and masks it with a word
// of the stream it shares with party 2; x1 = binary sharing of in.1, held in the clear by parties 1 and 2.
// dest = x0 + x1 through the depth-optimised adder circuit.
Sh3Task Sh3Converter::toBinaryMatrix(Sh3Task dep, const si64Matrix& in, sbMatrix& dest) {
    struct State { sbMatrix x0, x1; };
    auto state = std::make_shared<State>();
    return dep.then([&in, &dest, this, state](CommPkg& comm, Sh3Task self) {
        gpu::Context* ctx = gpu::current();
        if (dest.rows() == 0) dest.resize(in.rows(), 64 * in.cols());
        sbMatrix& x0 = state->x0;
        sbMatrix& x1 = state->x1;
        x0.resize(in.rows(), dest.bitCount());
        x1.resize(in.rows(), dest.bitCount());
        const u64 n = x0.i64Size();
        if (n != in.size()) throw std::runtime_error("toBinaryMatrix: dest.bitCount() must cover 64 bits per input word " LOCATION);
        auto zero = [&](eMatrix<i64>& m) { gpu::check(aby3cu_memset(ctx->h(), m.devOut(), 0, std::max<u64>(n, 1) * 8)); };
        switch (self.getRuntime().mPartyIdx) {
        case 0: {
            // x0[1] = next words of mPrevCommon; x0[0] = (in[0] + in[1]) ^ x0[1]      (:88-95)
            mRandGen->mPrevCommon.getDevice(ctx, x0.mShares[1].devOut(), 8 * n);
            eMatrix<i64> sum = in.mShares[0] + in.mShares[1];
            gpu::check(aby3cu_share_op(ctx->h(), ABY3CU_OP_XOR, sum.dev(), x0.mShares[1].dev(), x0.mShares[0].devOut(), n));
            maskLastWord(ctx, x0.mShares[0], dest.bitCount());
            maskLastWord(ctx, x0.mShares[1], dest.bitCount());
            zero(x1.mShares[0]); zero(x1.mShares[1]);
            comm.mNext.asyncSendDevice(x0.mShares[0].dev(), 8 * n);                  // :109
            mCir = getArithToBinCircuit(64, dest.bitCount());
            mBin.asyncEvaluate(self, &mCir, *mRandGen, {&x0, &x1}, {&dest}).then([state](Sh3Task) {});
            break;
        }
        case 1: {
            zero(x0.mShares[0]);
            x1.mShares[0] = in.mShares[0];                                            // :137-140
            maskLastWord(ctx, x1.mShares[0], dest.bitCount());
            zero(x1.mShares[1]);
            auto f = comm.mPrev.asyncRecvDevice(x0.mShares[1].devOut(), 8 * n);
            mCir = getArithToBinCircuit(64, dest.bitCount());
            self.then([f = std::move(f)](CommPkg&, Sh3Task) mutable { f.get(); });
            mBin.asyncEvaluate(self, &mCir, *mRandGen, {&x0, &x1}, {&dest}).then([state](Sh3Task) {});
            break;
        }
        case 2: {
            zero(x0.mShares[1]);
            mRandGen->mNextCommon.getDevice(ctx, x0.mShares[0].devOut(), 8 * n);    // :180-184
            x1.mShares[1] = in.mShares[1];
            maskLastWord(ctx, x0.mShares[0], dest.bitCount());
            maskLastWord(ctx, x1.mShares[1], dest.bitCount());
            zero(x1.mShares[0]);
            mCir = getArithToBinCircuit(64, dest.bitCount());
            mBin.asyncEvaluate(self, &mCir, *mRandGen, {&x0, &x1}, {&dest}).then([state](Sh3Task) {});
            break;
        }
        default: throw std::runtime_error("logic error. " LOCATION);
        }
    }).getClosure();
}

// Sh3Converter.cpp:211-370.  One arithmetic output element per input BIT (dest is rows x bitCount).
// Party 2 knows b = x1 ^ x2 of every bit and sends the pair (m + (0^b), m + (1^b)) with m = -d0 - d1 through both
// shared OTs; parties 0 and 1 both choose with x0 and so both learn m + (x0 ^ x1 ^ x2).
Sh3Task Sh3Converter::bitInjection(Sh3Task dep, const sbMatrix& in, si64Matrix& dest, bool twoRounds) {
    return dep.then([this, &in, &dest, twoRounds](CommPkg& comm, Sh3Task self) {
        if (!mRandGen) throw std::runtime_error("init was not called. " LOCATION);
        gpu::Context* ctx = gpu::current();
        dest.resize(in.rows(), in.bitCount());
        const u64 n = in.rows() * in.bitCount();
        const u64 words = in.i64Cols();
        auto choicesOf = [&](const eMatrix<i64>& plane) {
            auto c = std::make_shared<gpu::Buffer>(ctx, std::max<size_t>(8 * n, 16));
            gpu::check(aby3cu_bits_expand(ctx->h(), plane.dev(), in.rows(), words, in.bitCount(), (i64*)c->ptr()));
            return c;
        };
        switch (self.getRuntime().mPartyIdx) {
        case 0: {
            // receiver 0: chooses with its own share x0; sender P2 (prev), helper P1 (next)
            auto choices = choicesOf(in.mShares[0]);
            auto r = SharedOT::asyncRecv(comm.mPrev, comm.mNext, n);
            self.then([&dest, twoRounds, choices, r, n](CommPkg& comm, Sh3Task) {
                r.finish((const i64*)choices->ptr(), dest.mShares[0].devOut(), false);
                if (twoRounds) comm.mNext.asyncSendDevice(dest.mShares[0].dev(), 8 * n);
            });
            if (!twoRounds) mOT02.help(comm.mNext, (const i64*)choices->ptr(), n);   // helper for receiver 1
            mRandGen->mPrevCommon.getDevice(ctx, dest.mShares[1].devOut(), 8 * n);   // = party 2's d0
            break;
        }
        case 1: {
            // helper for receiver 0 with x0 (its prev share)
            auto choices = choicesOf(in.mShares[1]);
            mOT12.help(comm.mPrev, (const i64*)choices->ptr(), n);
            mRandGen->mNextCommon.getDevice(ctx, dest.mShares[0].devOut(), 8 * n);   // = party 2's d1
            if (!twoRounds) {
                auto r = SharedOT::asyncRecv(comm.mNext, comm.mPrev, n);             // sender P2 (next), helper P0 (prev)
                self.then([&dest, choices, r](CommPkg&, Sh3Task) { r.finish((const i64*)choices->ptr(), dest.mShares[1].devOut(), false); });
            } else {
                auto f = comm.mPrev.asyncRecvDevice(dest.mShares[1].devOut(), 8 * n).share();
                self.then([f](CommPkg&, Sh3Task) { f.get(); });
            }
            break;
        }
        case 2: {
            auto& g = *mRandGen;
            gpu::Buffer msgs(ctx, std::max<size_t>(16 * n, 16));
            gpu::check(aby3cu_bitinj_msgs(ctx->h(), in.mShares[0].dev(), in.mShares[1].dev(), in.rows(), words, in.bitCount(),
                                          g.mNextCommon.getSeed().data(), Sh3Evaluator::streamElem(g.mNextCommon),
                                          g.mPrevCommon.getSeed().data(), Sh3Evaluator::streamElem(g.mPrevCommon),
                                          dest.mShares[0].devOut(), dest.mShares[1].devOut(), (i64*)msgs.ptr()));
            g.mNextCommon.skip(8 * n);
            g.mPrevCommon.skip(8 * n);
            mOT12.send(comm.mNext, (const i64*)msgs.ptr(), n);                       // to receiver 0
            if (!twoRounds) mOT02.send(comm.mPrev, (const i64*)msgs.ptr(), n);       // to receiver 1
            break;
        }
        default: throw std::runtime_error("logic error");
        }
    }).getClosure();
}

// one depth-optimised adder per `base`-bit word of the inputs (Sh3Converter.cpp:372-411)
oc::BetaCircuit Sh3Converter::getArithToBinCircuit(u64 base, u64 bitCount) {
    oc::BetaCircuit cir;
    const u64 numWords = (base + bitCount - 1) / base;
    oc::BetaBundle in0(bitCount), in1(bitCount), out(bitCount), temp;
    cir.addInputBundle(in0);
    cir.addInputBundle(in1);
    cir.addOutputBundle(out);
    for (u64 i = 0; i < numWords; ++i) {
        const u64 begin = i * base, end = std::min<u64>(begin + base, bitCount);
        oc::BetaBundle w0, w1, o;
        w0.mWires.assign(in0.mWires.begin() + begin, in0.mWires.begin() + end);
        w1.mWires.assign(in1.mWires.begin() + begin, in1.mWires.begin() + end);
        o.mWires.assign(out.mWires.begin() + begin, out.mWires.begin() + end);
        oc::BetaLibrary::add_build(cir, w0, w1, o, temp, oc::BetaLibrary::IntType::TwosComplement, oc::BetaLibrary::Optimized::Depth);
    }
    return cir;
}

}  // namespace aby3
