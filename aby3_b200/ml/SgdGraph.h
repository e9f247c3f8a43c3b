// SgdGraph.h -- SGD_Linear (aby3-ML/Regression.h:112-184) for three parties that share ONE GPU, replayed as a
// CUDA graph.  An iteration of the reference loop is ~28 small kernels spread over three party threads; issued
// one by one it is bound by driver calls, not by the device.  Here ONE thread issues the three parties' kernels of
// an iteration on their three streams (the reshare between co-located parties is a pointer: the opened xy - r is
// read in place, ordered by events), captures that once, and replays it with a single launch per iteration.
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
            W[p].side.reset(new gpu::Context(c->device()));
        }
        void* evFork = c0->newEvent();
        const u64* it = (const u64*)dIter.ptr();
        u64 kernels = 0;

        // one iteration of all three parties; dependencies between parties are events, so the same code runs
        // eagerly (first iteration) and under stream capture (the graph)
        auto issue = [&] {
            kernels = 0;
            gpu::check(aby3cu_event_record(c0->h(), evFork));
            for (int p = 1; p < 3; ++p) gpu::check(aby3cu_event_wait(P[p].ctx->h(), evFork));
            // Main stream of a party: the critical chain  gather -> XX*w -> open/truncate -> XX^T*error -> open/truncate -> w -= update.
            // Side stream: both truncation pairs (they depend only on the iteration counter) and error -= YY (applied to
            // RTrunc BEFORE the opened value is added: addition commutes mod 2^64).  The transpose of the batch fills the
            // main stream's wait for the first truncation pair.
            for (int p = 0; p < 3; ++p) {
                aby3cu_ctx* h = P[p].ctx->h();
                auto& X = *P[p].X; auto& Y = *P[p].Y;
                const int64_t* in[4] = {X[0].dev(), X[1].dev(), Y[0].dev(), Y[1].dev()};
                int64_t* out[4] = {(i64*)W[p].XX[0].ptr(), (i64*)W[p].XX[1].ptr(), (i64*)W[p].YY[0].ptr(), (i64*)W[p].YY[1].ptr()};
                const uint64_t cols[4] = {F, F, 1, 1};
                gpu::check(aby3cu_gather_rows_multi_at(h, 4, in, cols, out, (const u64*)dIdx.ptr(), B, it));            // extractBatch
                gpu::check(aby3cu_event_record(h, W[p].evG));
                gpu::check(aby3cu_transpose_i64_2(h, (const i64*)W[p].XX[0].ptr(), (const i64*)W[p].XX[1].ptr(), B, F,
                                                  (i64*)W[p].XT[0].ptr(), (i64*)W[p].XT[1].ptr()));                                             // XX^T
                kernels += 2;
            }
            for (int p = 0; p < 3; ++p) {
                aby3cu_ctx* hs = W[p].side->h();
                i64* e0 = (i64*)W[p].E[0].ptr();
                i64* e1 = (i64*)W[p].E[1].ptr();
                gpu::check(aby3cu_event_wait(hs, evFork));
                gpu::check(aby3cu_trunc_tuple_at(hs, W[p].kn.data(), W[p].en, W[p].kp.data(), W[p].ep, it, S, D, nullptr,
                                                 (i64*)W[p].V1.ptr(), e0, e1, B));                                       // V1 = -r, E = RTrunc
                gpu::check(aby3cu_event_record(hs, W[p].evT1));
                gpu::check(aby3cu_trunc_tuple_at(hs, W[p].kn.data(), W[p].en + B, W[p].kp.data(), W[p].ep + B, it, S, D + aB, nullptr,
                                                 (i64*)W[p].V2.ptr(), (i64*)W[p].U[0].ptr(), (i64*)W[p].U[1].ptr(), F));
                gpu::check(aby3cu_event_wait(hs, W[p].evG));
                gpu::check(aby3cu_share_op2(hs, ABY3CU_OP_SUB, e0, (const i64*)W[p].YY[0].ptr(), e0, e1, (const i64*)W[p].YY[1].ptr(), e1, B));  // error -= YY
                gpu::check(aby3cu_event_record(hs, W[p].evT2));
                kernels += 3;
            }
            for (int p = 0; p < 3; ++p) {
                // error = mul(XX, w): V1 += XX*w   (Sh3Evaluator.cpp:651-700)
                aby3cu_ctx* h = P[p].ctx->h();
                auto& w = *P[p].w;
                gpu::check(aby3cu_event_wait(h, W[p].evT1));
                gpu::check(aby3cu_gemm_cross(h, ABY3CU_GEMM_AUTO, (const i64*)W[p].XX[0].ptr(), (const i64*)W[p].XX[1].ptr(),
                                             w[0].dev(), w[1].dev(), B, F, 1, (i64*)W[p].V1.ptr(), 1));
                gpu::check(aby3cu_event_record(h, W[p].evV1));
                ++kernels;
            }
            for (int p = 0; p < 3; ++p) {
                aby3cu_ctx* h = P[p].ctx->h();
                gpu::check(aby3cu_event_wait(h, W[p].evT2));       // side stream done: E -= YY, XX^T, V2 = -r, U = RTrunc
                if (p < 2) {                                       // parties 0 and 1 open xy - r and truncate (:703-724)
                    const int nx = (p + 1) % 3, pv = (p + 2) % 3;
                    gpu::check(aby3cu_event_wait(h, W[nx].evV1));
                    gpu::check(aby3cu_event_wait(h, W[pv].evV1));
                    gpu::check(aby3cu_trunc_finish(h, (const i64*)W[nx].V1.ptr(), (const i64*)W[pv].V1.ptr(), (const i64*)W[p].V1.ptr(),
                                                   (i64*)W[p].E[p].ptr(), B, D));
                    ++kernels;
                }
                // update = mulTruncate(XX^T, error, aB)
                gpu::check(aby3cu_gemm_cross(h, ABY3CU_GEMM_AUTO, (const i64*)W[p].XT[0].ptr(), (const i64*)W[p].XT[1].ptr(),
                                             (const i64*)W[p].E[0].ptr(), (const i64*)W[p].E[1].ptr(), F, B, 1, (i64*)W[p].V2.ptr(), 1));
                gpu::check(aby3cu_event_record(h, W[p].evV2));
                ++kernels;
            }
            for (int p = 0; p < 3; ++p) {
                aby3cu_ctx* h = P[p].ctx->h();
                if (p < 2) {
                    const int nx = (p + 1) % 3, pv = (p + 2) % 3;
                    gpu::check(aby3cu_event_wait(h, W[nx].evV2));
                    gpu::check(aby3cu_event_wait(h, W[pv].evV2));
                    gpu::check(aby3cu_trunc_finish(h, (const i64*)W[nx].V2.ptr(), (const i64*)W[pv].V2.ptr(), (const i64*)W[p].V2.ptr(),
                                                   (i64*)W[p].U[p].ptr(), F, D + aB));
                    ++kernels;
                }
                // w -= update, in place: the graph must find w where it left it
                auto& w = *P[p].w;
                i64* w0 = w[0].devMut();
                i64* w1 = w[1].devMut();
                gpu::check(aby3cu_share_op2(h, ABY3CU_OP_SUB, w0, (const i64*)W[p].U[0].ptr(), w0, w1, (const i64*)W[p].U[1].ptr(), w1, F));
                gpu::check(aby3cu_event_record(h, W[p].evDone));
                ++kernels;
            }
            for (int p = 1; p < 3; ++p) gpu::check(aby3cu_event_wait(c0->h(), W[p].evDone));
            gpu::check(aby3cu_counter_add(c0->h(), (u64*)dIter.ptr(), 1));
            ++kernels;
        };

        // every party stream must be idle-ordered behind what it did before: the fork event covers party 0's stream,
        // the other two are drained once
        for (int p = 1; p < 3; ++p) P[p].ctx->sync();
        issue();                                             // iteration 0, eagerly (also sets kernel attributes)
        void* exec = nullptr;
        if (iters > 1) {
            gpu::check(aby3cu_capture_begin(c0->h()));
            try { issue(); } catch (...) { void* dead = nullptr; aby3cu_capture_end(c0->h(), &dead); aby3cu_graph_destroy(dead); throw; }
            gpu::check(aby3cu_capture_end(c0->h(), &exec));
            for (u64 i = 1; i < iters; ++i) gpu::check(aby3cu_graph_launch(c0->h(), exec, kernels));
        }
        c0->sync();
        for (int p = 1; p < 3; ++p) P[p].ctx->sync();
        if (exec) aby3cu_graph_destroy(exec);
        for (int p = 0; p < 3; ++p) {
            P[p].ctx->recycleEvent(W[p].evV1); P[p].ctx->recycleEvent(W[p].evV2); P[p].ctx->recycleEvent(W[p].evDone);
            P[p].ctx->recycleEvent(W[p].evT1); P[p].ctx->recycleEvent(W[p].evT2); P[p].ctx->recycleEvent(W[p].evG);
            W[p].side->sync();
            auto& g = P[p].eval->mShareGen;
            g.mNextCommon.skip(8 * S * iters);               // what the kernels consumed
            g.mPrevCommon.skip(8 * S * iters);
        }
        c0->recycleEvent(evFork);
        return kernels;
    }
};

}  // namespace aby3
