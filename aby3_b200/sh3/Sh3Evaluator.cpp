// Sh3Evaluator.cpp -- see Sh3Evaluator.h.
#include "Sh3Evaluator.h"

namespace aby3 {

void Sh3Evaluator::init(u64 partyIdx, block prevSeed, block nextSeed, u64 buffSize) {
    mShareGen.init(prevSeed, nextSeed, buffSize);
    mPartyIdx = partyIdx;
    mOtPrevRecver.setSeed(mShareGen.mNextCommon.get<block>());     // Sh3Evaluator.cpp:13
    mOtNextRecver.setSeed(mShareGen.mPrevCommon.get<block>());     // :14
}

void Sh3Evaluator::init(u64 partyIdx, CommPkg& comm, block seed, u64 buffSize) {
    mShareGen.init(comm, seed, buffSize);
    mPartyIdx = partyIdx;
    mOtPrevRecver.setSeed(mShareGen.mNextCommon.get<block>());
    mOtNextRecver.setSeed(mShareGen.mPrevCommon.get<block>());
}

Sh3Evaluator::MulMode Sh3Evaluator::mulMode(const si64Matrix& A, const si64Matrix& B) {
    if (A.cols() == B.rows()) return MulMode::Matmul;
    if (A.rows() == B.rows() && A.cols() == B.cols()) return MulMode::Hadamard;
    throw std::runtime_error("asyncMul: operand shapes allow neither a matrix nor an element-wise product " LOCATION);
}

u64 Sh3Evaluator::streamElem(const oc::PRNG& p) {
    if (p.byteCursor() % 8) throw std::runtime_error("common PRNG cursor is not 8-byte aligned " LOCATION);
    return p.byteCursor() / 8;
}

// ---- scalar, no truncation -- Sh3Evaluator.cpp:71-89 ----------------------------
Sh3Task Sh3Evaluator::asyncMul(Sh3Task dependency, const si64& A, const si64& B, si64& C) {
    return dependency.then([&](CommPkg& comm, Sh3Task self) {
        C[0] = (i64)((u64)A[0] * (u64)B[0] + (u64)A[0] * (u64)B[1] + (u64)A[1] * (u64)B[0] + (u64)mShareGen.getShare());
        comm.mNext.asyncSendCopy(C[0]);
        auto fu = comm.mPrev.asyncRecv(C[1]);
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}

// ---- matrix, no truncation -- Sh3Evaluator.cpp:92-116 ---------------------------
Sh3Task Sh3Evaluator::asyncMul(Sh3Task dependency, const si64Matrix& A, const si64Matrix& B, si64Matrix& C) {
    return dependency.then([&](CommPkg& comm, Sh3Task self) {
        gpu::Context* ctx = gpu::current();
        const MulMode mode = mulMode(A, B);
        // the product is built in a fresh matrix so that C may alias A or B
        eMatrix<i64> c0;
        if (mode == MulMode::Matmul) {
            const u64 M = A.rows(), K = A.cols(), N = B.cols();
            c0.resize(M, N);
            // z first, then the contraction accumulates on top of it
            mShareGen.getShares(ctx, nullptr, c0.devOut(), M * N, false);
            gpu::check(aby3cu_gemm_cross(ctx->h(), mGemmAlgo, A.mShares[0].dev(), A.mShares[1].dev(),
                                         B.mShares[0].dev(), B.mShares[1].dev(), M, K, N, c0.devMut(), 1));
        } else {
            c0.resize(A.rows(), A.cols());
            const u64 n = c0.size();
            gpu::check(aby3cu_mul_hadamard(ctx->h(), A.mShares[0].dev(), A.mShares[1].dev(), B.mShares[0].dev(),
                                           B.mShares[1].dev(), mShareGen.mShareGen[0].key().data(),
                                           mShareGen.mShareGen[1].key().data(), mShareGen.mShareElemIdx, c0.devOut(), n));
            mShareGen.mShareElemIdx += n;
        }
        auto& C0 = C.mShares[0];
        C0 = std::move(c0);
        C.mShares[1].resizeLike(C0);
        const size_t bytes = C0.size() * sizeof(i64);
        comm.mNext.asyncSendDevice(C0.dev(), bytes);                               // :109
        auto fu = comm.mPrev.asyncRecvDevice(C.mShares[1].devOut(), bytes);        // :110
        self.then([fu = std::move(fu)](CommPkg&, Sh3Task&) mutable { fu.get(); });
    }).getClosure();
}

// ---- truncation pair -- Sh3Evaluator.cpp:503-566 ---------------------------------
TruncationPair Sh3Evaluator::getTruncationTuple(u64 xSize, u64 ySize, u64 d) {
    gpu::Context* ctx = gpu::current();
    TruncationPair pair;
    pair.mR.resize(xSize, ySize);
    pair.mRTrunc.resize(xSize, ySize);
    const u64 n = xSize * ySize;
    if (DEBUG_disable_randomization) {
        gpu::check(aby3cu_trunc_tuple(ctx->h(), nullptr, 0, nullptr, 0, d, pair.mR.devOut(), nullptr,
                                      pair.mRTrunc.mShares[0].devOut(), pair.mRTrunc.mShares[1].devOut(), n));
    } else {
        auto& g = mShareGen;
        gpu::check(aby3cu_trunc_tuple(ctx->h(), g.mNextCommon.getSeed().data(), streamElem(g.mNextCommon),
                                      g.mPrevCommon.getSeed().data(), streamElem(g.mPrevCommon), d, pair.mR.devOut(),
                                      nullptr, pair.mRTrunc.mShares[0].devOut(), pair.mRTrunc.mShares[1].devOut(), n));
        g.mNextCommon.skip(8 * n);     // :526
        g.mPrevCommon.skip(8 * n);     // :527
    }
    return pair;
}

// ---- scalar with truncation -- Sh3Evaluator.cpp:568-648 --------------------------
Sh3Task Sh3Evaluator::asyncMul(Sh3Task dependency, const si64& A, const si64& B, si64& C, u64 shift) {
    return dependency.then([&, shift](CommPkg& comm, Sh3Task& self) -> void {
        i64 r = 0, t0 = 0, t1 = 0;
        if (!DEBUG_disable_randomization) {
            t0 = mShareGen.mNextCommon.get<i64>();
            t1 = mShareGen.mPrevCommon.get<i64>();
            r = t0 >> 2; t0 >>= (shift + 2); t1 >>= (shift + 2);
        }
        i64 abMinusR = (i64)((u64)A[0] * (u64)B[0] + (u64)A[0] * (u64)B[1] + (u64)A[1] * (u64)B[0] - (u64)r);
        C[0] = t0; C[1] = t1;
        auto& rt = self.getRuntime();
        const u64 next = (rt.mPartyIdx + 1) % 3, prev = (rt.mPartyIdx + 2) % 3;
        if (next < 2) comm.mNext.asyncSendCopy(abMinusR);
        if (prev < 2) comm.mPrev.asyncSendCopy(abMinusR);
        if (rt.mPartyIdx < 2) {
            auto shares = std::make_shared<std::array<i64, 3>>();
            auto fu0 = comm.mNext.asyncRecv((*shares)[0]).share();
            auto fu1 = comm.mPrev.asyncRecv((*shares)[1]).share();
            (*shares)[2] = abMinusR;
            self.then([fu0, fu1, shares, &C, shift, this](CommPkg&, Sh3Task&) mutable {
                fu0.get(); fu1.get();
                const i64 sum = (i64)((u64)(*shares)[0] + (u64)(*shares)[1] + (u64)(*shares)[2]);
                C.mData[mPartyIdx] = (i64)((u64)C.mData[mPartyIdx] + (u64)(sum >> shift));
            });
        }
    }).getClosure();
}

// ---- matrix with truncation -- Sh3Evaluator.cpp:651-730 --------------------------
Sh3Task Sh3Evaluator::asyncMul(Sh3Task dependency, const si64Matrix& A, const si64Matrix& B, si64Matrix& C, u64 shift) {
    return dependency.then([&, shift](CommPkg& comm, Sh3Task& self) -> void {
        gpu::Context* ctx = gpu::current();
        const MulMode mode = mulMode(A, B);
        const u64 M = A.rows(), N = (mode == MulMode::Matmul) ? B.cols() : A.cols(), n = M * N;
        const bool rnd = !DEBUG_disable_randomization;
        auto& g = mShareGen;
        const block seedNext = g.mNextCommon.getSeed(), seedPrev = g.mPrevCommon.getSeed();
        const u8* kn = rnd ? seedNext.data() : nullptr;
        const u8* kp = rnd ? seedPrev.data() : nullptr;
        const u64 en = rnd ? streamElem(g.mNextCommon) : 0, ep = rnd ? streamElem(g.mPrevCommon) : 0;

        // abMinusR is written once and then only read -- by this party's continuation and by the parties it is
        // opened to -- so it is sent without a staging copy and received without one (SharedBuffer / Borrowed)
        struct Scratch { std::shared_ptr<gpu::SharedBuffer> v; oc::Borrowed s0, s1; };
        auto sc = std::make_shared<Scratch>();
        const size_t bytes = std::max<size_t>(n * sizeof(i64), 16);
        // The truncation pair depends on the common keystreams only, not on A or B.  For a matrix product big enough to
        // matter it is produced on the party's second stream into blocks that are free EARLY (gpu::Context::allocEarly),
        // so it runs under whatever this party and its neighbours still have in flight -- typically the contraction of
        // the previous product -- and the party's own stream meets it just before the contraction (joinAux).
        const bool early = mode == MulMode::Matmul && bytes >= gpu::Context::kEarlyMin && mEarlyTruncation;
        sc->v = early ? std::make_shared<gpu::SharedBuffer>(ctx, bytes, gpu::Early{})
                      : std::make_shared<gpu::SharedBuffer>(ctx, bytes);
        i64* V = (i64*)sc->v->ptr();

        // RTrunc is produced into fresh matrices and moved into C only after the
        // cross term has been enqueued, so C may alias A or B (as in the reference,
        // where C.mShares = move(RTrunc.mShares) follows the product, :673)
        eMatrix<i64> rt0(M, N), rt1(M, N);
        if (early) {
            rt0.adoptDevice(gpu::Buffer(ctx, bytes, gpu::Early{}));
            rt1.adoptDevice(gpu::Buffer(ctx, bytes, gpu::Early{}));
        }
        i64* RT0 = rt0.devOut();
        i64* RT1 = rt1.devOut();
        // block-wise open (parties on different GPUs): same for every party, a function of the shape only
        u64 blockRows = 0, nBlocks = 1;
        if (mode == MulMode::Matmul && mOpenBlocks > 1 && M >= 256) {
            blockRows = ((M + mOpenBlocks - 1) / mOpenBlocks + 127) / 128 * 128;
            nBlocks = (M + blockRows - 1) / blockRows;
        }
        std::vector<void*> blockDone;
        if (mode == MulMode::Matmul) {
            // V = -r (pre-load), then V += A0*B0 + A0*B1 + A1*B0   (:662-665, :672)
            gpu::check(aby3cu_trunc_tuple(early ? ctx->aux()->h() : ctx->h(), kn, en, kp, ep, shift, nullptr, V, RT0, RT1, n));
            // the limb pre-pass of the operands does not wait for the pair, the contraction (which accumulates onto -r) does
            const i64 *a0 = A.mShares[0].dev(), *a1 = A.mShares[1].dev(), *b0 = B.mShares[0].dev(), *b1 = B.mShares[1].dev();
            void* pairDone = early ? ctx->aux()->recordEvent() : nullptr;
            int rc;
            static const bool ringOn = [] { const char* e = std::getenv("ABY3_RING_GEMV"); return !(e && e[0] == '0'); }();
            if (ringOn && mColocated && N == 1 && nBlocks == 1 && M * A.cols() >= kRingMin) {
                // co-located parties: one launch for the three cross terms, every plane of A read once (a1 is the previous
                // party's a0 and is read there)
                gpu::ColocatedGroup::GemvJob job;
                job.ctx = ctx; job.a0 = a0; job.b0 = b0; job.b1 = b1; job.c = V; job.M = M; job.K = A.cols();
                job.ready = ctx->recordEvent();
                job.ready2 = pairDone;
                rc = 0;
                try { mColocated->ringGemv((int)self.getRuntime().mPartyIdx, job); }
                catch (...) { ctx->recycleEvent(job.ready); if (pairDone) ctx->aux()->recycleEvent(pairDone); throw; }
                ctx->recycleEvent(job.ready);
            } else if (nBlocks > 1) {
                for (u64 b = 0; b < nBlocks; ++b) blockDone.push_back(ctx->newEvent());
                rc = aby3cu_gemm_cross_blocks(ctx->h(), mGemmAlgo, a0, a1, b0, b1, M, A.cols(), N, V, 1, pairDone, blockRows,
                                              blockDone.data(), (u32)nBlocks);
            } else {
                rc = aby3cu_gemm_cross_after(ctx->h(), mGemmAlgo, a0, a1, b0, b1, M, A.cols(), N, V, 1, pairDone);
            }
            if (pairDone) ctx->aux()->recycleEvent(pairDone);
            if (rc) for (void* e : blockDone) ctx->recycleEvent(e);
            gpu::check(rc);
        } else {
            gpu::check(aby3cu_mul_hadamard_trunc(ctx->h(), A.mShares[0].dev(), A.mShares[1].dev(), B.mShares[0].dev(),
                                                 B.mShares[1].dev(), kn, en, kp, ep, shift, V, RT0, RT1, n));
        }
        if (rnd) { g.mNextCommon.skip(8 * n); g.mPrevCommon.skip(8 * n); }
        C.mShares[0] = std::move(rt0);
        C.mShares[1] = std::move(rt1);

        // open xy - r to parties 0 and 1 (:676-684)
        auto& rt = self.getRuntime();
        const u64 next = (rt.mPartyIdx + 1) % 3, prev = (rt.mPartyIdx + 2) % 3;
        if (nBlocks > 1) {
            // block b travels on the communication stream once its rows are final; later blocks are still multiplying
            gpu::Context* cm = ctx->comm();
            struct Recv { gpu::Buffer s0, s1; std::vector<std::shared_future<void>> fu; };
            auto rc = std::make_shared<Recv>();
            if (rt.mPartyIdx < 2) { rc->s0.reset(ctx, bytes); rc->s1.reset(ctx, bytes); }
            {   // the receive buffers come from the party's pool: the communication stream starts behind their last use
                void* e = ctx->recordEvent();
                gpu::check(aby3cu_event_wait(cm->h(), e));
                ctx->recycleEvent(e);
            }
            for (u64 b = 0; b < nBlocks; ++b) {
                const u64 r0 = b * blockRows, rows = std::min(blockRows, M - r0);
                const size_t off = r0 * N * sizeof(i64), len = rows * N * sizeof(i64);
                gpu::check(aby3cu_event_wait(cm->h(), blockDone[b]));
                ctx->recycleEvent(blockDone[b]);
                if (next < 2) comm.mNext.asyncSendDeviceSharedOn(cm, sc->v, off, len);
                if (prev < 2) comm.mPrev.asyncSendDeviceSharedOn(cm, sc->v, off, len);
                if (rt.mPartyIdx < 2) {
                    rc->fu.push_back(comm.mNext.asyncRecvDeviceOn(cm, (u8*)rc->s0.ptr() + off, len).share());
                    rc->fu.push_back(comm.mPrev.asyncRecvDeviceOn(cm, (u8*)rc->s1.ptr() + off, len).share());
                }
                comm.mNext.flushOn(cm);          // NCCL: the block's sends and receives are ONE group (both channels share the endpoint)
                comm.mPrev.flushOn(cm);
            }
            if (rt.mPartyIdx < 2) {
                self.then([rc, sc, &C, shift, n, ctx, this](CommPkg&, Sh3Task&) mutable {
                    for (auto& f : rc->fu) f.get();
                    ctx->joinComm();
                    gpu::check(aby3cu_trunc_finish(ctx->h(), (const i64*)rc->s0.ptr(), (const i64*)rc->s1.ptr(),
                                                   (const i64*)sc->v->ptr(), C.mShares[mPartyIdx].devMut(), n, shift));
                });
            }
            return;
        }
        if (next < 2) comm.mNext.asyncSendDeviceShared(sc->v, n * sizeof(i64));
        if (prev < 2) comm.mPrev.asyncSendDeviceShared(sc->v, n * sizeof(i64));
        if (rt.mPartyIdx < 2) {
            // two senders, possibly on two other GPUs: the second message lands through the communication stream so that
            // the two NVLink copies run side by side
            const bool big = bytes >= (size_t(8) << 20);
            auto fu0 = comm.mNext.asyncRecvDeviceBorrow(n * sizeof(i64), &sc->s0).share();
            auto fu1 = comm.mPrev.asyncRecvDeviceBorrow(n * sizeof(i64), &sc->s1, big ? ctx->comm() : nullptr).share();
            self.then([fu0, fu1, sc, &C, shift, n, ctx, big, this](CommPkg&, Sh3Task&) mutable {
                fu0.get(); fu1.get();
                if (big) ctx->joinComm();
                // C[mPartyIdx] += (s0 + s1 + v) >> shift   (:712-718)
                gpu::check(aby3cu_trunc_finish(ctx->h(), (const i64*)sc->s0.ptr, (const i64*)sc->s1.ptr,
                                               (const i64*)sc->v->ptr(), C.mShares[mPartyIdx].devMut(), n, shift));
                sc->s0.release(ctx);
                sc->s1.release(ctx);
            });
        }
    }).getClosure();
}

// ---- arithmetic x one shared bit -- Sh3Evaluator.cpp:119-263 ----------------------
// c = b * a.  P1 is the OT receiver twice: from P0 it learns b*(a0+a2) - c0 - c2 - z (helper P2),
// from P2 it learns b*a1 + z (helper P0).  Draw order from the common PRNGs is per element and
// interleaved exactly as in the reference (it decides every party's share values).
Sh3Task Sh3Evaluator::asyncMul(Sh3Task dep, const si64Matrix& A, const sbMatrix& B, si64Matrix& c) {
    return dep.then([&](CommPkg& comm, Sh3Task self) {
        if (B.rows() != A.rows() || A.cols() != 1 || B.bitCount() != 1) throw std::runtime_error(LOCATION);
        gpu::Context* ctx = gpu::current();
        const u64 n = A.rows();
        auto& g = mShareGen;
        // results are built aside so that c may alias A (Sh3Piecewise passes the same matrix)
        eMatrix<i64> c0(n, 1), c1(n, 1);
        const block seedPrev = g.mPrevCommon.getSeed(), seedNext = g.mNextCommon.getSeed();
        switch (mPartyIdx) {
        case 0: {
            gpu::Buffer msgs(ctx, std::max<size_t>(16 * n, 16));
            gpu::check(aby3cu_bitmul_msgs_p0(ctx->h(), A.mShares[0].dev(), A.mShares[1].dev(), B.mShares[0].dev(), B.mShares[1].dev(),
                                             seedPrev.data(), streamElem(g.mPrevCommon), seedNext.data(), streamElem(g.mNextCommon),
                                             c0.devOut(), c1.devOut(), (i64*)msgs.ptr(), n));
            g.mPrevCommon.skip(16 * n);          // z and c[1] per element
            g.mNextCommon.skip(8 * n);           // c[0]
            mOtNextRecver.send(comm.mNext, (const i64*)msgs.ptr(), n);        // sender, receiver P1, helper P2
            mOtNextRecver.help(comm.mNext, B.mShares[0].dev(), n);           // helper for P2 -> P1, choice b0
            c.mShares[0] = std::move(c0);
            c.mShares[1] = std::move(c1);
            break;
        }
        case 1: {
            g.mPrevCommon.getDevice(ctx, c1.devOut(), 8 * n);                // c[1]
            auto f0 = SharedOT::asyncRecv(comm.mPrev, comm.mNext, n);        // sender P0, helper P2
            auto f1 = SharedOT::asyncRecv(comm.mNext, comm.mPrev, n);        // sender P2, helper P0
            c.mShares[1] = std::move(c1);
            self.then([&, f0, f1, n](CommPkg& comm, Sh3Task) {
                eMatrix<i64> r(n, 1);
                f1.finish(B.mShares[1].dev(), r.devOut(), false);             // b*a1 + z
                f0.finish(B.mShares[0].dev(), r.devMut(), true);              // + b*(a0+a2) - c0 - c2 - z
                c.mShares[0] = std::move(r);
                comm.mNext.asyncSendDevice(c.mShares[0].dev(), 8 * n);
            });
            break;
        }
        case 2: {
            gpu::Buffer msgs(ctx, std::max<size_t>(16 * n, 16));
            gpu::check(aby3cu_bitmul_msgs_p2(ctx->h(), A.mShares[1].dev(), B.mShares[0].dev(), B.mShares[1].dev(), seedNext.data(),
                                             streamElem(g.mNextCommon), c0.devOut(), (i64*)msgs.ptr(), n));
            g.mNextCommon.skip(16 * n);          // z and c[0] per element
            mOtPrevRecver.help(comm.mPrev, B.mShares[1].dev(), n);           // helper for P0 -> P1, choice b1
            mOtPrevRecver.send(comm.mPrev, (const i64*)msgs.ptr(), n);        // sender, receiver P1, helper P0
            c.mShares[0] = std::move(c0);
            c.mShares[1].resize(n, 1);
            self.then([&, n](CommPkg& comm, Sh3Task self) {
                auto f = comm.mPrev.asyncRecvDevice(c.mShares[1].devOut(), 8 * n).share();
                self.then([f](CommPkg&, Sh3Task) { f.get(); });
            });
            break;
        }
        default: throw RTE_LOC;
        }
    }).getClosure();
}

// ---- public constant x one shared bit -- Sh3Evaluator.cpp:418-501 -------------------
Sh3Task Sh3Evaluator::asyncMul(Sh3Task dep, const i64& a, const sbMatrix& b, si64Matrix& c) {
    return dep.then([&, a](CommPkg& comm, Sh3Task self) {
        if (b.bitCount() != 1) throw RTE_LOC;
        gpu::Context* ctx = gpu::current();
        const u64 n = b.rows();
        if (c.rows() != n || c.cols() != 1) c.resize(n, 1);
        auto& g = mShareGen;
        switch (mPartyIdx) {
        case 0: {
            gpu::Buffer msgs(ctx, std::max<size_t>(16 * n, 16));
            gpu::check(aby3cu_bitmul_pub_msgs(ctx->h(), a, b.mShares[0].dev(), b.mShares[1].dev(), g.mShareGen[0].key().data(),
                                              g.mShareGen[1].key().data(), g.mShareElemIdx, (i64*)msgs.ptr(), n));
            g.mShareElemIdx += n;
            mOtNextRecver.send(comm.mNext, (const i64*)msgs.ptr(), n);
            mOtPrevRecver.send(comm.mPrev, (const i64*)msgs.ptr(), n);
            auto fu1 = comm.mNext.asyncRecvDevice(c.mShares[0].devOut(), 8 * n).share();
            auto fu2 = comm.mPrev.asyncRecvDevice(c.mShares[1].devOut(), 8 * n).share();
            self.then([fu1, fu2](CommPkg&, Sh3Task) { fu1.get(); fu2.get(); });
            break;
        }
        case 1: {
            g.getShares(ctx, nullptr, c.mShares[1].devOut(), n, false);
            mOtNextRecver.help(comm.mNext, b.mShares[0].dev(), n);
            comm.mPrev.asyncSendDevice(c.mShares[1].dev(), 8 * n);
            auto f = SharedOT::asyncRecv(comm.mPrev, comm.mNext, n);
            self.then([&, f](CommPkg&, Sh3Task) { f.finish(b.mShares[0].dev(), c.mShares[0].devOut(), false); });
            break;
        }
        case 2: {
            g.getShares(ctx, nullptr, c.mShares[0].devOut(), n, false);
            mOtPrevRecver.help(comm.mPrev, b.mShares[1].dev(), n);
            comm.mNext.asyncSendDevice(c.mShares[0].dev(), 8 * n);
            auto f = SharedOT::asyncRecv(comm.mNext, comm.mPrev, n);
            self.then([&, f](CommPkg&, Sh3Task) { f.finish(b.mShares[1].dev(), c.mShares[1].devOut(), false); });
            break;
        }
        default: throw std::runtime_error(LOCATION);
        }
    }).getClosure();
}

}  // namespace aby3
