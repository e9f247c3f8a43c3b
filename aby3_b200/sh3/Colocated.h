// Colocated.h -- what three parties that share ONE GPU can do together.
//
// In a replicated sharing the second plane a party holds is the previous party's first plane, so when the three parties of a
// GEMV-shaped product (N = 1: logistic inference, aby3-ML/aby3ML.h:102-139) sit on one GPU, each of them streaming both of its
// planes of A reads every plane twice.  The group lets the three party threads meet: each posts the pointers of its cross term
// (Sh3Evaluator.cpp:662-665), the last one to arrive launches ONE kernel for all three (aby3cu_gemv_ring: every plane read
// once) on its own stream behind the others' "inputs ready" events, and the others' streams are ordered behind that launch.
// Same share words as three separate launches as long as the sharings are consistent -- which every sharing the protocols
// produce is.  Host-side rendezvous only; the parties still draw their own keystreams and exchange their own messages.
#pragma once
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>

#include "Gpu.h"

namespace aby3 {
namespace gpu {

class ColocatedGroup {
public:
    struct GemvJob {
        Context* ctx = nullptr;
        const i64 *a0 = nullptr, *b0 = nullptr, *b1 = nullptr;
        i64* c = nullptr;                 // accumulated onto (it holds -r of the truncation pair)
        void* ready = nullptr;            // event on the party's stream: a0, b0, b1 are valid
        void* ready2 = nullptr;           // optional second event (the truncation pair written by the party's second stream)
        u64 M = 0, K = 0;
    };

    // Called by each of the three parties with the same M and K.  Returns once the common launch has been enqueued and the
    // caller's stream waits for it.  Throws on every party if the shapes disagree, the launch fails or a party never arrives.
    void ringGemv(int party, const GemvJob& job) {
        if (party < 0 || party > 2 || !job.ctx) throw std::runtime_error("ColocatedGroup: bad party " LOCATION);
        std::unique_lock<std::mutex> lk(mMtx);
        const u64 gen = mGen;
        mJobs[party] = job;
        if (++mArrived == 3) {
            mArrived = 0;
            mErr.clear();
            try { launch(party); } catch (const std::exception& e) { mErr = e.what(); }
            ++mGen;
            mCv.notify_all();
        } else if (!mCv.wait_for(lk, std::chrono::seconds(120), [&] { return mGen != gen; })) {
            --mArrived;
            throw std::runtime_error("ColocatedGroup: the other parties did not reach the common product " LOCATION);
        }
        if (!mErr.empty()) throw std::runtime_error(mErr);
        if (party != mLeader) {
            void* e = mDone[party];
            mDone[party] = nullptr;
            lk.unlock();
            check(aby3cu_event_wait(job.ctx->h(), e));
            EventPool::put(mLeaderDevice, e);
        }
    }

private:
    void launch(int leader) {
        Context* ctx = mJobs[leader].ctx;
        const u64 M = mJobs[leader].M, K = mJobs[leader].K;
        const i64 *a0[3], *b0[3], *b1[3];
        i64* c[3];
        for (int p = 0; p < 3; ++p) {
            const GemvJob& j = mJobs[p];
            if (j.M != M || j.K != K) throw std::runtime_error("ColocatedGroup: the parties' shapes differ " LOCATION);
            if (j.ctx->device() != ctx->device()) throw std::runtime_error("ColocatedGroup: the parties are not on one GPU " LOCATION);
            if (p != leader && j.ready) check(aby3cu_event_wait(ctx->h(), j.ready));
            if (j.ready2) check(aby3cu_event_wait(ctx->h(), j.ready2));
            a0[p] = j.a0; b0[p] = j.b0; b1[p] = j.b1; c[p] = j.c;
        }
        check(aby3cu_gemv_ring(ctx->h(), a0, b0, b1, M, K, c, 1));
        mLeader = leader;
        mLeaderDevice = ctx->device();
        for (int p = 0; p < 3; ++p)
            if (p != leader) mDone[p] = ctx->recordEvent();
    }

    std::mutex mMtx;
    std::condition_variable mCv;
    GemvJob mJobs[3];
    void* mDone[3] = {nullptr, nullptr, nullptr};
    std::string mErr;
    u64 mGen = 0;
    int mArrived = 0, mLeader = 0, mLeaderDevice = 0;
};

}  // namespace gpu
}  // namespace aby3
