// SgdGraph.h -- SGD_Linear (aby3-ML/Regression.h:112-184) for three parties that share ONE GPU, replayed as a
// CUDA graph.  An iteration of the reference loop is ~28 small kernels spread over three party threads; issued
// one by one it is bound by driver calls, not by the device.  Here ONE thread issues an iteration of all three parties
// -- every protocol step as one launch over the three parties' pointers (aby3cu_*_batch; the reshare between co-located
// parties is a pointer: the opened xy - r is read in place) -- captures that once, and replays it with a single launch
// per iteration: 10 kernels instead of 28.
// Everything that changes from one iteration to the next -- the mini-batch rows and the offsets into the
// common-PRNG keystreams -- is read from a device-resident iteration counter (aby3cu_*_at entry points).
//
// Same kernels, same keystream offsets, same arithmetic as SGD_Linear over the sh3 facade: the resulting shares of w
// and the PRNG cursors are bit-identical (tests/test_gpu_sh3.py::test_graph_sgd_matches_facade_and_oracle).
#pragma once
#include <array>
#include <cmath>
#include <memory>

#include "Regression.h"

namespace aby3 {

template <Decimal D>
class ColocatedSgdLinear {
public:
    struct PartyRef {
        gpu::Context* ctx;
        Sh3Evaluator* eval;
        sf64Matrix<D>* X;
        sf64Matrix<D>* Y;
        sf64Matrix<D>* w;
    };

    // runs params.mIterations iterations; returns the number of kernels per iteration (for launch accounting)
    static u64 run(std::array<PartyRef, 3> P, const RegressionParam& params, const std::vector<u64>& batchIndices) {
        const u64 B = params.mBatchSize, iters = params.mIterations;
        const u64 F = P[0].X->cols(), rows = P[0].X->rows();
        if (!iters) return 0;
        if (batchIndices.size() != iters * B) throw std::runtime_error(LOCATION);
        for (u64 i : batchIndices) if (i >= rows) throw std::runtime_error("ColocatedSgdLinear: batch index out of range " LOCATION);
        for (auto& p : P) {
            if (p.ctx->device() != P[0].ctx->device()) throw std::runtime_error("ColocatedSgdLinear: the parties must share one GPU " LOCATION);
            if (p.X->rows() != rows || p.X->cols() != F || p.Y->rows() != rows || p.Y->cols() != 1 || p.w->rows() != F || p.w->cols() != 1)
                throw std::runtime_error(LOCATION);
            if (p.eval->DEBUG_disable_randomization) throw std::runtime_error("ColocatedSgdLinear: randomisation must be on " LOCATION);
        }
        const u64 aB = (u64)std::log2(1 / (params.mLearningRate / B));      // Regression.h:139
        const u64 S = B + F;                                                // keystream elements per iteration and stream

        gpu::Context* c0 = P[0].ctx;
        gpu::Buffer dIdx(c0, std::max<size_t>(batchIndices.size() * 8, 16)), dIter(c0, 16);
        gpu::check(aby3cu_h2d(c0->h(), dIdx.ptr(), batchIndices.data(), batchIndices.size() * 8));
        gpu::check(aby3cu_memset(c0->h(), dIter.ptr(), 0, 16));
        c0->sync();                                                         // batchIndices may be pageable

        struct Work {
            gpu::Buffer XX[2], YY[2], V1, E[2], XT[2], V2, U[2];
            u64 en = 0, ep = 0;
            block kn, kp;
            void* evV1 = nullptr;
            void* evV2 = nullptr;
            void* evDone = nullptr;
            void* evT1 = nullptr;
            void* evT2 = nullptr;
            void* evG = nullptr;
            std::unique_ptr<gpu::Context> side;      // second stream: the truncation pairs do not depend on the batch
        } W[3];
        for (int p = 0; p < 3; ++p) {
            gpu::Context* c = P[p].ctx;
            for (int s = 0; s < 2; ++s) {
                W[p].XX[s].reset(c, B * F * 8); W[p].YY[s].reset(c, B * 8); W[p].E[s].reset(c, B * 8);
                W[p].XT[s].reset(c, B * F * 8); W[p].U[s].reset(c, F * 8);
            }
            W[p].V1.reset(c, B * 8);
            W[p].V2.reset(c, F * 8);
            auto& g = P[p].eval->mShareGen;
            W[p].en = Sh3Evaluator::streamElem(g.mNextCommon);
            W[p].ep = Sh3Evaluator::streamElem(g.mPrevCommon);
            W[p].kn = g.mNextCommon.getSeed();
            W[p].kp = g.mPrevCommon.getSeed();
            W[p].evV1 = c->newEvent(); W[p].evV2 = c->newEvent(); W[p].evDone = c->newEvent();
            W[p].evT1 = c->newEvent(); W[p].evT2 = c->newEvent(); W[p].evG = c->newEvent();
            if (p == 0) W[p].side.reset(new gpu::Context(c->device()));
        }
        void* evFork = c0->newEvent();
        const u64* it = (const u64*)dIter.ptr();
        u64 kernels = 0;

        // one iteration of all three parties; dependencies between parties are events, so the same code runs
        // eagerly (first iteration) and under stream capture (the graph)
        // One iteration of all three parties.  Every protocol step is ONE launch over the three parties' pointers
        // (aby3cu_*_batch): the same kernel would otherwise be launched once per party (and per share plane).
        // Main stream: the critical chain  gather -> XX*w -> open/truncate -> XX^T*error -> open/truncate -> w -= update.
        // Side stream: what does not sit on it -- both truncation pairs of all parties (they depend only on the iteration
        // counter), error -= YY (applied to RTrunc BEFORE the opened value is added: addition commutes mod 2^64) and the
        // transposes of the batch.  The same code runs eagerly (first iteration) and under stream capture (the graph).
        aby3cu_ctx* hm = c0->h();
        aby3cu_ctx* hs = W[0].side->h();
        auto issue = [&] {
            kernels = 0;
            gpu::check(aby3cu_event_record(hm, evFork));
            gpu::check(aby3cu_event_wait(hs, evFork));
            {   // extractBatch for every party: rows of X and Y, both planes
                const int64_t* in[12]; int64_t* out[12]; uint64_t cols[12];
                for (int p = 0; p < 3; ++p) {
                    auto& X = *P[p].X; auto& Y = *P[p].Y;
                    in[4 * p] = X[0].dev(); in[4 * p + 1] = X[1].dev(); in[4 * p + 2] = Y[0].dev(); in[4 * p + 3] = Y[1].dev();
                    out[4 * p] = (i64*)W[p].XX[0].ptr(); out[4 * p + 1] = (i64*)W[p].XX[1].ptr();
                    out[4 * p + 2] = (i64*)W[p].YY[0].ptr(); out[4 * p + 3] = (i64*)W[p].YY[1].ptr();
                    cols[4 * p] = cols[4 * p + 1] = F; cols[4 * p + 2] = cols[4 * p + 3] = 1;
                }
                gpu::check(aby3cu_gather_rows_multi_at(hm, 12, in, cols, out, (const u64*)dIdx.ptr(), B, it));
                gpu::check(aby3cu_event_record(hm, W[0].evG));
                ++kernels;
            }
            {   // side: V1 = -r, E = RTrunc (shift D, n = B) and V2 = -r', U = RTrunc' (shift D + aB, n = F) for every party
                const uint8_t* kn[6]; const uint8_t* kp[6]; uint64_t en[6], ep[6], sh[6], cnt[6]; int64_t* negr[6]; int64_t* r0[6]; int64_t* r1[6];
                for (int p = 0; p < 3; ++p) {
                    kn[p] = kn[3 + p] = W[p].kn.data(); kp[p] = kp[3 + p] = W[p].kp.data();
                    en[p] = W[p].en; ep[p] = W[p].ep; sh[p] = D; cnt[p] = B;
                    negr[p] = (i64*)W[p].V1.ptr(); r0[p] = (i64*)W[p].E[0].ptr(); r1[p] = (i64*)W[p].E[1].ptr();
                    en[3 + p] = W[p].en + B; ep[3 + p] = W[p].ep + B; sh[3 + p] = D + aB; cnt[3 + p] = F;
                    negr[3 + p] = (i64*)W[p].V2.ptr(); r0[3 + p] = (i64*)W[p].U[0].ptr(); r1[3 + p] = (i64*)W[p].U[1].ptr();
                }
                gpu::check(aby3cu_trunc_tuple_batch_at(hs, 6, kn, en, kp, ep, sh, cnt, negr, r0, r1, it, S));
                gpu::check(aby3cu_event_record(hs, W[0].evT1));
                gpu::check(aby3cu_event_wait(hs, W[0].evG));
                const int64_t* x[6]; const int64_t* y[6]; int64_t* o[6]; const int64_t* tin[6]; int64_t* tout[6];
                for (int p = 0; p < 3; ++p)
                    for (int s2 = 0; s2 < 2; ++s2) {
                        x[2 * p + s2] = (const i64*)W[p].E[s2].ptr(); y[2 * p + s2] = (const i64*)W[p].YY[s2].ptr(); o[2 * p + s2] = (i64*)W[p].E[s2].ptr();
                        tin[2 * p + s2] = (const i64*)W[p].XX[s2].ptr(); tout[2 * p + s2] = (i64*)W[p].XT[s2].ptr();
                    }
                gpu::check(aby3cu_share_op_batch(hs, ABY3CU_OP_SUB, 6, x, y, o, B));                       // error -= YY
                gpu::check(aby3cu_transpose_i64_batch(hs, 6, tin, B, F, tout));                            // XX^T
                gpu::check(aby3cu_event_record(hs, W[0].evT2));
                kernels += 3;
            }
            const int64_t* a0[3]; const int64_t* a1[3]; const int64_t* b0[3]; const int64_t* b1[3]; int64_t* cc[3];
            const int64_t* f0[2]; const int64_t* f1[2]; const int64_t* f2[2]; int64_t* fc[2];
            {   // error = mul(XX, w): V1 += XX*w   (Sh3Evaluator.cpp:651-700), then parties 0 and 1 open and truncate (:703-724)
                for (int p = 0; p < 3; ++p) {
                    auto& w = *P[p].w;
                    a0[p] = (const i64*)W[p].XX[0].ptr(); a1[p] = (const i64*)W[p].XX[1].ptr(); b0[p] = w[0].dev(); b1[p] = w[1].dev();
                    cc[p] = (i64*)W[p].V1.ptr();
                }
                gpu::check(aby3cu_event_wait(hm, W[0].evT1));
                gpu::check(aby3cu_gemv_cross_batch(hm, 3, a0, a1, b0, b1, B, F, cc, 1));
                for (int p = 0; p < 2; ++p) {
                    f0[p] = (const i64*)W[(p + 1) % 3].V1.ptr(); f1[p] = (const i64*)W[(p + 2) % 3].V1.ptr(); f2[p] = (const i64*)W[p].V1.ptr();
                    fc[p] = (i64*)W[p].E[p].ptr();
                }
                gpu::check(aby3cu_event_wait(hm, W[0].evT2));          // side stream done: E -= YY, XX^T, V2, U
                gpu::check(aby3cu_trunc_finish_batch(hm, 2, f0, f1, f2, fc, B, D));
                kernels += 2;
            }
            {   // update = mulTruncate(XX^T, error, aB)
                for (int p = 0; p < 3; ++p) {
                    a0[p] = (const i64*)W[p].XT[0].ptr(); a1[p] = (const i64*)W[p].XT[1].ptr();
                    b0[p] = (const i64*)W[p].E[0].ptr(); b1[p] = (const i64*)W[p].E[1].ptr(); cc[p] = (i64*)W[p].V2.ptr();
                }
                gpu::check(aby3cu_gemv_cross_batch(hm, 3, a0, a1, b0, b1, F, B, cc, 1));
                for (int p = 0; p < 2; ++p) {
                    f0[p] = (const i64*)W[(p + 1) % 3].V2.ptr(); f1[p] = (const i64*)W[(p + 2) % 3].V2.ptr(); f2[p] = (const i64*)W[p].V2.ptr();
                    fc[p] = (i64*)W[p].U[p].ptr();
                }
                gpu::check(aby3cu_trunc_finish_batch(hm, 2, f0, f1, f2, fc, F, D + aB));
                kernels += 2;
            }
            {   // w -= update, in place: the graph must find w where it left it
                const int64_t* x[6]; const int64_t* y[6]; int64_t* o[6];
                for (int p = 0; p < 3; ++p)
                    for (int s2 = 0; s2 < 2; ++s2) {
                        i64* wp = (*P[p].w)[s2].devMut();
                        x[2 * p + s2] = wp; y[2 * p + s2] = (const i64*)W[p].U[s2].ptr(); o[2 * p + s2] = wp;
                    }
                gpu::check(aby3cu_share_op_batch(hm, ABY3CU_OP_SUB, 6, x, y, o, F));
                gpu::check(aby3cu_counter_add(hm, (u64*)dIter.ptr(), 1));
                kernels += 2;
            }
        };

        // every party stream must be idle-ordered behind what it did before: the fork event covers party 0's stream,
        // the other two are drained once
        for (int p = 1; p < 3; ++p) P[p].ctx->sync();
        issue();                                             // iteration 0, eagerly (also sets kernel attributes)
        void* exec = nullptr;
        if (iters > 1) {
            gpu::check(aby3cu_capture_begin(c0->h()));
            for (int p = 0; p < 3; ++p) P[p].ctx->setCapturing(true);        // pool releases record no events meanwhile
            try { issue(); } catch (...) {
                for (int p = 0; p < 3; ++p) P[p].ctx->setCapturing(false);
                void* dead = nullptr; aby3cu_capture_end(c0->h(), &dead); aby3cu_graph_destroy(dead); throw;
            }
            for (int p = 0; p < 3; ++p) P[p].ctx->setCapturing(false);
            gpu::check(aby3cu_capture_end(c0->h(), &exec));
            for (u64 i = 1; i < iters; ++i) gpu::check(aby3cu_graph_launch(c0->h(), exec, kernels));
        }
        c0->sync();
        for (int p = 1; p < 3; ++p) P[p].ctx->sync();
        if (exec) aby3cu_graph_destroy(exec);
        for (int p = 0; p < 3; ++p) {
            P[p].ctx->recycleEvent(W[p].evV1); P[p].ctx->recycleEvent(W[p].evV2); P[p].ctx->recycleEvent(W[p].evDone);
            P[p].ctx->recycleEvent(W[p].evT1); P[p].ctx->recycleEvent(W[p].evT2); P[p].ctx->recycleEvent(W[p].evG);
            if (W[p].side) W[p].side->sync();
            auto& g = P[p].eval->mShareGen;
            g.mNextCommon.skip(8 * S * iters);               // what the kernels consumed
            g.mPrevCommon.skip(8 * S * iters);
        }
        c0->recycleEvent(evFork);
        return kernels;
    }

    // whether runFused() takes this problem (else run() -- the graph replay -- does)
    static bool fusedSupports(u64 B, u64 F) { return F >= 2 && F % 2 == 0 && B >= 1 && B <= 2048; }

    // The same training run as ONE persistent kernel (csrc/sgd_fused.cu): the grid walks the iterations itself, two grid
    // barriers per iteration, batch rows read straight from X.  Same arithmetic and keystream offsets as run().
    static void runFused(std::array<PartyRef, 3> P, const RegressionParam& params, const std::vector<u64>& batchIndices) {
        const u64 B = params.mBatchSize, iters = params.mIterations;
        const u64 F = P[0].X->cols(), rows = P[0].X->rows();
        if (!iters) return;
        if (batchIndices.size() != iters * B) throw std::runtime_error(LOCATION);
        if (!fusedSupports(B, F)) throw std::runtime_error("ColocatedSgdLinear::runFused: unsupported shape " LOCATION);
        for (auto& p : P) {
            if (p.ctx->device() != P[0].ctx->device()) throw std::runtime_error("ColocatedSgdLinear: the parties must share one GPU " LOCATION);
            if (p.X->rows() != rows || p.X->cols() != F || p.Y->rows() != rows || p.Y->cols() != 1 || p.w->rows() != F || p.w->cols() != 1)
                throw std::runtime_error(LOCATION);
            if (p.eval->DEBUG_disable_randomization) throw std::runtime_error("ColocatedSgdLinear: randomisation must be on " LOCATION);
        }
        for (u64 i : batchIndices) if (i >= rows) throw std::runtime_error("ColocatedSgdLinear: batch index out of range " LOCATION);
        const u64 aB = (u64)std::log2(1 / (params.mLearningRate / B));      // Regression.h:139
        const u64 S = B + F;

        gpu::Context* c0 = P[0].ctx;
        gpu::Buffer dIdx(c0, std::max<size_t>(batchIndices.size() * 8, 16)), work(c0, aby3cu_sgd_linear_colocated_work_bytes(B));
        gpu::check(aby3cu_h2d(c0->h(), dIdx.ptr(), batchIndices.data(), batchIndices.size() * 8));
        const int64_t* X[6]; const int64_t* Y[6]; int64_t* w[6];
        const uint8_t* kn[3]; const uint8_t* kp[3]; uint64_t en[3], ep[3];
        block seedN[3], seedP[3];
        for (int p = 0; p < 3; ++p) {
            for (int s = 0; s < 2; ++s) {
                X[2 * p + s] = (*P[p].X)[s].dev(); Y[2 * p + s] = (*P[p].Y)[s].dev(); w[2 * p + s] = (*P[p].w)[s].devMut();
            }
            auto& g = P[p].eval->mShareGen;
            seedN[p] = g.mNextCommon.getSeed(); seedP[p] = g.mPrevCommon.getSeed();
            kn[p] = seedN[p].data(); kp[p] = seedP[p].data();
            en[p] = Sh3Evaluator::streamElem(g.mNextCommon); ep[p] = Sh3Evaluator::streamElem(g.mPrevCommon);
        }
        // the other parties' streams may still be producing X, Y or w: party 0's stream runs the kernel behind them
        for (int p = 1; p < 3; ++p) {
            void* e = P[p].ctx->recordEvent();
            gpu::check(aby3cu_event_wait(c0->h(), e));
            P[p].ctx->recycleEvent(e);
        }
        gpu::check(aby3cu_sgd_linear_colocated(c0->h(), X, Y, w, (const u64*)dIdx.ptr(), F, B, iters, D, D + aB, kn, en, kp, ep, work.ptr()));
        c0->sync();                                                         // batchIndices may be pageable; w is final
        for (int p = 0; p < 3; ++p) {
            auto& g = P[p].eval->mShareGen;
            g.mNextCommon.skip(8 * S * iters);               // what the kernel consumed
            g.mPrevCommon.skip(8 * S * iters);
        }
    }
};

}  // namespace aby3
