// harness.cpp -- a small C API over the sh3 facade so that tests/ and bench.py
// (Python, ctypes) can drive three in-process parties exactly the way the
// reference's unit tests do (three threads, one Sh3Runtime / Sh3Encryptor /
// Sh3Evaluator each: aby3_tests/Sh3EvaluatorTests.cpp:20-135).  Every party is a
// persistent worker thread with its own device context and stream; the parties
// may sit on one GPU or on three.  Not part of the reference API.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <thread>

#include "Sh3BinaryEvaluator.h"
#include "Sh3Encryptor.h"
#include "Sh3Evaluator.h"
#include "Sh3Converter.h"
#include "Sh3Piecewise.h"
#include "../basic/Basics.h"
#include "../ml/Regression.h"
#include "../ml/SgdGraph.h"

using namespace aby3;

namespace {

thread_local std::string g_err;

struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has = false, stop = false, done = false;
    std::string error;
    void loop() {
        for (;;) {
            std::function<void()> j;
            {
                std::unique_lock<std::mutex> l(m);
                cv.wait(l, [&] { return has || stop; });
                if (stop) return;
                j = std::move(job);
                has = false;
            }
            std::string err;
            try { j(); } catch (const std::exception& e) { err = e.what(); } catch (...) { err = "unknown exception"; }
            {
                std::lock_guard<std::mutex> l(m);
                error = err;
                done = true;
            }
            cv.notify_all();
        }
    }
    void submit(std::function<void()> j) {
        {
            std::lock_guard<std::mutex> l(m);
            job = std::move(j); has = true; done = false;
        }
        cv.notify_all();
    }
    std::string wait() {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&] { return done; });
        return error;
    }
};

struct Party {
    std::unique_ptr<gpu::Context> ctx;
    std::unique_ptr<gpu::Context> copy;      // upload stream of the party: overlapped h2d of plaintext matrices
    std::unique_ptr<gpu::Context> copyOut;   // download stream: PCIe is full duplex, results must not queue behind the next inputs
    CommPkg comm;
    Sh3Runtime rt;
    Sh3Encryptor enc;
    Sh3Evaluator eval;
    Sh3Converter conv;
    bool convInit = false;
    std::map<int, std::unique_ptr<si64Matrix>> ints;
    std::map<int, std::unique_ptr<sbMatrix>> bins;
    std::map<int, std::unique_ptr<i64Matrix>> plains;
    std::map<int, std::unique_ptr<sPackedBin>> packs;
    void* ev_start = nullptr;
    void* ev_stop = nullptr;
};

}  // namespace

struct sh3h {
    Party p[3];
    Worker w[3];
    int next_handle = 1;
    void* ev_end = nullptr;
    std::vector<oc::detail::NcclApi::comm_t> nccl_comms;

    // run f(party) on the three party threads, wait for all; returns 0 / sets g_err
    int run(const std::function<void(int)>& f) {
        for (int i = 0; i < 3; ++i)
            w[i].submit([=] {
                f(i);
                // NCCL transport: nothing may stay queued when the party goes idle
                if (p[i].comm.mNext.isConnected()) { p[i].comm.mNext.flush(); p[i].comm.mPrev.flush(); }
            });
        std::string err;
        for (int i = 0; i < 3; ++i) {
            std::string e = w[i].wait();
            if (!e.empty() && err.empty()) err = "party " + std::to_string(i) + ": " + e;
        }
        if (!err.empty()) { g_err = err; return 1; }
        return 0;
    }
};

extern "C" {

const char* sh3h_last_error(void) { return g_err.c_str(); }

// seeds: [party][0 = prev, 1 = next][16] for the encryptor and the evaluator
static sh3h* create_impl(int dev0, int dev1, int dev2, const uint8_t* enc_seeds, const uint8_t* eval_seeds, int use_nccl);

sh3h* sh3h_create(int dev0, int dev1, int dev2, const uint8_t* enc_seeds, const uint8_t* eval_seeds) {
    return create_impl(dev0, dev1, dev2, enc_seeds, eval_seeds, 0);
}
// three parties on ONE GPU sharing ONE stream: enqueue order alone orders the parties' kernels, so a
// message hand-over needs no CUDA event (latency-bound protocols such as SGD count driver calls)
sh3h* sh3h_create_shared_stream(int dev, const uint8_t* enc_seeds, const uint8_t* eval_seeds) {
    return create_impl(dev, dev, dev, enc_seeds, eval_seeds, 2);
}
// parties on three DIFFERENT GPUs, reshare by ncclSend/ncclRecv over NVLink
sh3h* sh3h_create_nccl(int dev0, int dev1, int dev2, const uint8_t* enc_seeds, const uint8_t* eval_seeds) {
    return create_impl(dev0, dev1, dev2, enc_seeds, eval_seeds, 1);
}

static sh3h* create_impl(int dev0, int dev1, int dev2, const uint8_t* enc_seeds, const uint8_t* eval_seeds, int use_nccl) {
    try {
        std::unique_ptr<sh3h> h(new sh3h);
        const int dev[3] = {dev0, dev1, dev2};
        for (int i = 0; i < 3; ++i) {
            if (use_nccl == 2 && i > 0) h->p[i].ctx.reset(new gpu::Context(dev[i], h->p[0].ctx->stream()));
            else h->p[i].ctx.reset(new gpu::Context(dev[i]));
        }
        if (use_nccl == 1) {
            if (dev0 == dev1 || dev1 == dev2 || dev0 == dev2)
                throw std::runtime_error("NCCL transport needs three different GPUs (kernels that wait on each other must not share a device)");
            auto& api = oc::detail::NcclApi::get();
            h->nccl_comms.resize(3);
            api.check(api.CommInitAll(h->nccl_comms.data(), 3, dev), "CommInitAll");
            for (int i = 0; i < 3; ++i) {
                auto ep = std::make_shared<oc::detail::NcclEndpoint>();
                ep->comm = h->nccl_comms[i];
                ep->ctx = h->p[i].ctx.get();
                h->p[i].comm = CommPkg{oc::Channel::makeNccl(ep, (i + 2) % 3), oc::Channel::makeNccl(ep, (i + 1) % 3)};
            }
        } else {
        // chl01 / chl02 / chl12 exactly as Sh3EvaluatorTests.cpp:23-36; comm = {prev, next}
        auto c01 = oc::Channel::makePair(h->p[0].ctx.get(), h->p[1].ctx.get());
        auto c02 = oc::Channel::makePair(h->p[0].ctx.get(), h->p[2].ctx.get());
        auto c12 = oc::Channel::makePair(h->p[1].ctx.get(), h->p[2].ctx.get());
        h->p[0].comm = CommPkg{c02.first, c01.first};
        h->p[1].comm = CommPkg{c01.second, c12.first};
        h->p[2].comm = CommPkg{c12.second, c02.second};
        }
        for (int i = 0; i < 3; ++i) h->w[i].th = std::thread([hp = h.get(), i] { hp->w[i].loop(); });
        // three parties on one GPU: GEMV-shaped products meet in one launch (sh3/Colocated.h)
        if (use_nccl != 1 && dev0 == dev1 && dev1 == dev2) {
            auto group = std::make_shared<gpu::ColocatedGroup>();
            for (int i = 0; i < 3; ++i) h->p[i].eval.mColocated = group;
        }
        auto blk = [](const uint8_t* p) { block b; memcpy(b.data(), p, 16); return b; };
        int rc = h->run([&](int i) {
            Party& P = h->p[i];
            gpu::setCurrent(P.ctx.get());
            P.rt.init(i, P.comm);
            P.enc.init(i, blk(enc_seeds + (2 * i) * 16), blk(enc_seeds + (2 * i + 1) * 16));
            P.eval.init(i, blk(eval_seeds + (2 * i) * 16), blk(eval_seeds + (2 * i + 1) * 16));
            // parties on different GPUs: the opened xy - r leaves in row blocks while the product is still running
            if (dev[0] != dev[1] || dev[1] != dev[2]) {
                const char* e = std::getenv("ABY3_OPEN_BLOCKS");
                // (measured on 3 x B200: the opens are 0.2 ms messages next to a 2.7 ms contraction -- one message is as fast
                // as 2 blocks and faster than 4; profiles/r2_distributed_open_blocks.jsonl)
                P.eval.mOpenBlocks = e ? std::max(1, atoi(e)) : 1;
            }
            gpu::check(aby3cu_event_create(P.ctx->h(), &P.ev_start));
            gpu::check(aby3cu_event_create(P.ctx->h(), &P.ev_stop));
        });
        if (rc) return nullptr;
        gpu::check(aby3cu_event_create(h->p[0].ctx->h(), &h->ev_end));
        return h.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void sh3h_destroy(sh3h* h) {
    if (!h) return;
    h->run([&](int i) {
        Party& P = h->p[i];
        P.ints.clear(); P.bins.clear(); P.plains.clear(); P.packs.clear();
        if (P.copy) P.copy->sync();
        if (P.copyOut) P.copyOut->sync();
        P.ctx->sync();
    });
    for (int i = 0; i < 3; ++i) {
        { std::lock_guard<std::mutex> l(h->w[i].m); h->w[i].stop = true; }
        h->w[i].cv.notify_all();
        h->w[i].th.join();
    }
    if (!h->nccl_comms.empty()) {
        auto& api = oc::detail::NcclApi::get();
        for (auto c : h->nccl_comms) api.CommDestroy(c);
    }
    delete h;
}

int sh3h_set_disable_randomization(sh3h* h, int on) {
    for (int i = 0; i < 3; ++i) h->p[i].eval.DEBUG_disable_randomization = on != 0;
    return 0;
}
int sh3h_set_open_blocks(sh3h* h, uint64_t blocks) {
    for (int i = 0; i < 3; ++i) h->p[i].eval.mOpenBlocks = blocks ? blocks : 1;
    return 0;
}
int sh3h_set_gemm_algo(sh3h* h, int algo) {
    for (int i = 0; i < 3; ++i) h->p[i].eval.mGemmAlgo = algo;
    return 0;
}

// cursors: same six numbers as orc_session_cursors
int sh3h_cursors(sh3h* h, int party, uint64_t c[6]) {
    Party& P = h->p[party];
    c[0] = P.enc.mShareGen.mShareElemIdx; c[1] = P.eval.mShareGen.mShareElemIdx;
    c[2] = P.eval.mShareGen.mPrevCommon.byteCursor(); c[3] = P.eval.mShareGen.mNextCommon.byteCursor();
    c[4] = P.enc.mShareGen.mPrevCommon.byteCursor(); c[5] = P.enc.mShareGen.mNextCommon.byteCursor();
    return 0;
}

// A plaintext matrix living at `owner` (page-locked host storage).  Returns a handle;
// *host_ptr is where the caller writes the values (rows*cols int64, row-major).
int sh3h_plain_create(sh3h* h, int owner, uint64_t rows, uint64_t cols, int64_t** host_ptr) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        if (i != owner) return;
        auto m = std::make_unique<i64Matrix>(rows, cols);
        *host_ptr = m->data();
        h->p[i].plains[id] = std::move(m);
    });
    return rc ? -1 : id;
}
// after the caller has (re)written the host values
int sh3h_plain_touch(sh3h* h, int owner, int id) {
    return h->run([&](int i) { if (i == owner) (void)h->p[i].plains.at(id)->data(); });
}

// ---- overlapped transfers (eMatrix::prefetchDevice / fetchHostAsync) -------------------------------------------
static gpu::Context* copyCtx(Party& P) {
    if (!P.copy) P.copy.reset(new gpu::Context(P.ctx->device()));
    return P.copy.get();
}
static gpu::Context* copyOutCtx(Party& P) {
    if (!P.copyOut) P.copyOut.reset(new gpu::Context(P.ctx->device()));
    return P.copyOut.get();
}
// start the upload of a plaintext matrix on the owner's copy stream; a later sh3h_share waits for it on the device
int sh3h_plain_prefetch(sh3h* h, int owner, int id) {
    return h->run([&](int i) { if (i == owner) h->p[i].plains.at(id)->prefetchDevice(copyCtx(h->p[i])); });
}
// enc.revealAll on every party; party `who` reveals into its plaintext matrix `plain_id` and starts the download on its
// copy stream WITHOUT waiting for it (sh3h_plain_wait completes it)
int sh3h_reveal_plain_async(sh3h* h, int id, int who, int plain_id) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        if (i == who) {
            i64Matrix& dest = *P.plains.at(plain_id);
            P.enc.revealAll(P.comm, *P.ints.at(id), dest);
            dest.fetchHostAsync(copyOutCtx(P));
        } else {
            i64Matrix dest;
            P.enc.revealAll(P.comm, *P.ints.at(id), dest);
        }
    });
}
int sh3h_plain_wait(sh3h* h, int who, int plain_id) {
    return h->run([&](int i) { if (i == who) h->p[i].plains.at(plain_id)->waitHost(); });
}

// Sh3Encryptor::localIntMatrix at `owner`, remoteIntMatrix elsewhere (binary: *BinMatrix).
int sh3h_share(sh3h* h, int owner, int plain_id, uint64_t rows, uint64_t cols, int binary, uint64_t bit_count) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        if (!binary) {
            auto m = std::make_unique<si64Matrix>(rows, cols);
            if (i == owner) P.enc.localIntMatrix(P.comm, *P.plains.at(plain_id), *m);
            else P.enc.remoteIntMatrix(P.comm, *m);
            P.ints[id] = std::move(m);
        } else {
            auto m = std::make_unique<sbMatrix>(rows, bit_count);
            if (i == owner) P.enc.localBinMatrix(P.comm, *P.plains.at(plain_id), *m);
            else P.enc.remoteBinMatrix(P.comm, *m);
            P.bins[id] = std::move(m);
        }
    });
    return rc ? -1 : id;
}

// install raw share planes (tests: start from the oracle's shares). shares = [3][2][rows*cols]
int sh3h_set_shares(sh3h* h, const int64_t* shares, uint64_t rows, uint64_t cols, int binary, uint64_t bit_count) {
    const int id = h->next_handle++;
    const uint64_t wcols = binary ? (bit_count + 63) / 64 : cols;
    const uint64_t n = rows * wcols;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        eMatrix<i64>* pl[2];
        if (!binary) {
            auto m = std::make_unique<si64Matrix>(rows, cols);
            pl[0] = &m->mShares[0]; pl[1] = &m->mShares[1];
            P.ints[id] = std::move(m);
        } else {
            auto m = std::make_unique<sbMatrix>(rows, bit_count);
            pl[0] = &m->mShares[0]; pl[1] = &m->mShares[1];
            P.bins[id] = std::move(m);
        }
        for (int s = 0; s < 2; ++s) {
            memcpy(pl[s]->data(), shares + ((uint64_t)i * 2 + s) * n, n * 8);
            (void)pl[s]->dev();
        }
        P.ctx->sync();
    });
    return rc ? -1 : id;
}

int sh3h_get_shares(sh3h* h, int id, int binary, int64_t* out) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        const eMatrix<i64>* pl[2];
        if (!binary) { auto& m = *P.ints.at(id); pl[0] = &m.mShares[0]; pl[1] = &m.mShares[1]; }
        else { auto& m = *P.bins.at(id); pl[0] = &m.mShares[0]; pl[1] = &m.mShares[1]; }
        const uint64_t n = pl[0]->size();
        for (int s = 0; s < 2; ++s) memcpy(out + ((uint64_t)i * 2 + s) * n, pl[s]->hostData(), n * 8);
    });
}

int sh3h_shape(sh3h* h, int id, int binary, uint64_t* rows, uint64_t* cols) {
    try {
        Party& P = h->p[0];
        if (!binary) { *rows = P.ints.at(id)->rows(); *cols = P.ints.at(id)->cols(); }
        else { *rows = P.bins.at(id)->rows(); *cols = P.bins.at(id)->i64Cols(); }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int sh3h_free(sh3h* h, int id) {
    return h->run([&](int i) { h->p[i].ints.erase(id); h->p[i].bins.erase(id); h->p[i].plains.erase(id); h->p[i].packs.erase(id); });
}

// eval.asyncMul(rt, A, B, C [, shift]).get() on every party.  shift < 0: no truncation.
// out_id == 0: allocate a new result handle; otherwise reuse that matrix.
int sh3h_mul(sh3h* h, int a, int b, int64_t shift, int out_id) {
    const int id = out_id ? out_id : h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto& A = *P.ints.at(a);
        auto& B = *P.ints.at(b);
        auto& slot = P.ints[id];
        if (!slot) slot = std::make_unique<si64Matrix>();
        if (shift < 0) P.eval.asyncMul(P.rt, A, B, *slot).get();
        else P.eval.asyncMul(P.rt, A, B, *slot, (u64)shift).get();
    });
    return rc ? -1 : id;
}

// c = b * a: eval.asyncMul(rt, si64Matrix, sbMatrix, c) (pub != 0: the public constant `a_pub`)
int sh3h_mul_bit(sh3h* h, int a_id, int b_id, int pub, int64_t a_pub) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<si64Matrix>();
        if (pub) P.eval.asyncMul(P.rt, (const i64&)a_pub, *P.bins.at(b_id), *m).get();
        else P.eval.asyncMul(P.rt, *P.ints.at(a_id), *P.bins.at(b_id), *m).get();
        P.ints[id] = std::move(m);
    });
    return rc ? -1 : id;
}

// Sh3Piecewise::eval on an n x 1 fixed-point sharing.  Region r has coef_counts[r] coefficients
// (constant first); coefficient k is the integer coef_int[k] when coef_is_int[k], else coef_dbl[k].
int sh3h_piecewise(sh3h* h, int in_id, const double* thresholds, int n_thresholds, const int* coef_counts,
                   const int* coef_is_int, const int64_t* coef_int, const double* coef_dbl, uint64_t D) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        Sh3Piecewise pw;
        for (int t = 0; t < n_thresholds; ++t) pw.mThresholds.emplace_back(thresholds[t]);
        int k = 0;
        pw.mCoefficients.resize(n_thresholds + 1);
        for (int r = 0; r <= n_thresholds; ++r)
            for (int c = 0; c < coef_counts[r]; ++c, ++k) {
                if (coef_is_int[k]) pw.mCoefficients[r].emplace_back((i64)coef_int[k]);
                else pw.mCoefficients[r].emplace_back(coef_dbl[k]);
            }
        auto& in = *P.ints.at(in_id);
        auto out = std::make_unique<si64Matrix>(in.rows(), 1);
        pw.eval(P.rt.noDependencies(), in, *out, D, P.eval).get();
        P.ints[id] = std::move(out);
        P.ctx->sync();
    });
    return rc ? -1 : id;
}

// ---- bit-sliced sharings (Sh3Encryptor::localPackedBinary / remotePackedBinary / revealAll) ----------
// rows secrets of 64 * cols bits from `owner`'s plaintext matrix; shares_out (may be null): [3][2][bits * simd] words
int sh3h_share_packed(sh3h* h, int owner, int plain_id, uint64_t rows, uint64_t cols, int use_task, int64_t* shares_out) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<sPackedBin>(rows, 64 * cols);
        if (use_task) {
            if (i == owner) P.enc.localPackedBinary(P.rt.noDependencies(), *P.plains.at(plain_id), *m).get();
            else P.enc.remotePackedBinary(P.rt.noDependencies(), *m).get();
        } else {
            if (i == owner) P.enc.localPackedBinary(P.comm, *P.plains.at(plain_id), *m);
            else P.enc.remotePackedBinary(P.comm, *m);
        }
        if (shares_out)
            for (int s = 0; s < 2; ++s)
                memcpy(shares_out + ((uint64_t)i * 2 + s) * m->size(), m->mShares[s].hostData(), m->size() * 8);
        P.packs[id] = std::move(m);
    });
    return rc ? -1 : id;
}
// enc.revealAll(comm, sPackedBin, dest) on every party; party `who`'s rows x words result is copied out
int sh3h_reveal_packed(sh3h* h, int id, int who, int64_t* out) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        i64Matrix dest;
        P.enc.revealAll(P.comm, *P.packs.at(id), dest);
        if (i == who) memcpy(out, dest.hostData(), dest.size() * 8);
        else P.ctx->sync();
    });
}

// ---- Sh3Converter --------------------------------------------------------------------
static Sh3Converter& converter(Party& P) {
    if (!P.convInit) { P.conv.init(P.rt, P.eval.mShareGen); P.convInit = true; }
    return P.conv;
}
// conv.init(rt, eval.mShareGen) on every party (idempotent; the conversions below call it on first use)
int sh3h_conv_init(sh3h* h) {
    return h->run([&](int i) { (void)converter(h->p[i]); });
}
// conv.toBinaryMatrix(rt, si64Matrix in, sbMatrix dest).get(): arithmetic -> binary, 64 bits per word
int sh3h_conv_a2b(sh3h* h, int in_id, uint64_t bits) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<sbMatrix>();
        if (bits) m->resize(P.ints.at(in_id)->rows(), bits);        // a pre-sized destination fixes the bit count (ragged last word)
        converter(P).toBinaryMatrix(P.rt.noDependencies(), *P.ints.at(in_id), *m).get();
        // the closure completes with the last circuit round; getOutput and the state-keeping continuation hang
        // off the inner closure (Sh3BinaryEvaluator.cpp:467-473, Sh3Converter.cpp:112) and run with the queue,
        // as in the reference's test (run(t0, t1, t2), aby3_tests/Sh3ConverterTests.cpp:336)
        P.rt.runAll();
        P.bins[id] = std::move(m);
        P.ctx->sync();
    });
    return rc ? -1 : id;
}
// conv.bitInjection(rt, sbMatrix in, si64Matrix dest, twoRounds).get(): one arithmetic element per input bit
int sh3h_conv_bit_injection(sh3h* h, int in_id, int two_rounds) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<si64Matrix>();
        converter(P).bitInjection(P.rt.noDependencies(), *P.bins.at(in_id), *m, two_rounds != 0).get();
        P.ints[id] = std::move(m);
        P.ctx->sync();
    });
    return rc ? -1 : id;
}
// toPackedBin followed by toBinaryMatrix(sPackedBin); the packed planes ([3][2][bitCount * simd] words) are copied
// to packed_out (may be null) for comparison with the oracle's bit transpose
int sh3h_conv_packed_roundtrip(sh3h* h, int in_id, int64_t* packed_out, uint64_t* simd_width) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        sPackedBin pk;
        converter(P).toPackedBin(*P.bins.at(in_id), pk);
        if (i == 0 && simd_width) *simd_width = pk.simdWidth();
        if (packed_out)
            for (int s = 0; s < 2; ++s)
                memcpy(packed_out + ((uint64_t)i * 2 + s) * pk.size(), pk.mShares[s].hostData(), pk.size() * 8);
        auto m = std::make_unique<sbMatrix>();
        converter(P).toBinaryMatrix(pk, *m);
        P.bins[id] = std::move(m);
        P.ctx->sync();
    });
    return rc ? -1 : id;
}

// ---- aby3-Basic building blocks (basic/Basics.h) -----------------------------------
// res = (A > B) on arithmetic sharings: cipher_gt (BuildingBlocks.cpp:525-532)
int sh3h_cipher_gt(sh3h* h, int a_id, int b_id) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<sbMatrix>();
        basic::cipher_gt(i, *P.ints.at(a_id), *P.ints.at(b_id), *m, P.eval, P.rt);
        P.bins[id] = std::move(m);
    });
    return rc ? -1 : id;
}
// (max, min) of two binary sharings of 64-bit values: bool_cipher_max_min_split (BoolBasic.cpp:275-312)
int sh3h_max_min_split(sh3h* h, int a_id, int b_id, int* max_id, int* min_id) {
    *max_id = h->next_handle++;
    *min_id = h->next_handle++;
    return h->run([&](int i) {
        Party& P = h->p[i];
        auto mx = std::make_unique<sbMatrix>(), mn = std::make_unique<sbMatrix>();
        basic::bool_cipher_max_min_split(i, *P.bins.at(a_id), *P.bins.at(b_id), *mx, *mn, P.enc, P.eval, P.rt);
        P.bins[*max_id] = std::move(mx);
        P.bins[*min_id] = std::move(mn);
    });
}
// odd_even_merge of two sorted binary sharings (Sort.cpp:327-406)
int sh3h_odd_even_merge(sh3h* h, int a_id, int b_id) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<sbMatrix>();
        basic::odd_even_merge(*P.bins.at(a_id), *P.bins.at(b_id), *m, i, P.enc, P.eval, P.rt);
        P.bins[id] = std::move(m);
    });
    return rc ? -1 : id;
}

// C = A + B / A - B on shares (local)
int sh3h_addsub(sh3h* h, int a, int b, int sub) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        auto m = std::make_unique<si64Matrix>();
        *m = sub ? (*P.ints.at(a) - *P.ints.at(b)) : (*P.ints.at(a) + *P.ints.at(b));
        P.ints[id] = std::move(m);
    });
    return rc ? -1 : id;
}

// enc.revealAll on every party; party `who`'s result is copied to out
int sh3h_reveal(sh3h* h, int id, int binary, int who, int64_t* out) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        i64Matrix dest;
        if (!binary) P.enc.revealAll(P.comm, *P.ints.at(id), dest);
        else P.enc.revealAll(P.comm, *P.bins.at(id), dest);
        if (i == who) memcpy(out, dest.hostData(), dest.size() * 8);
        else P.ctx->sync();
    });
}

// enc.revealAll on every party; party `who` reveals straight into its page-locked plaintext
// matrix `plain_id` (the d2h copy lands in the buffer the caller reads: no host-side copy)
int sh3h_reveal_plain(sh3h* h, int id, int who, int plain_id) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        if (i == who) {
            i64Matrix& dest = *P.plains.at(plain_id);
            P.enc.revealAll(P.comm, *P.ints.at(id), dest);
            (void)dest.hostData();
        } else {
            i64Matrix dest;
            P.enc.revealAll(P.comm, *P.ints.at(id), dest);
            P.ctx->sync();
        }
    });
}

// getTruncationTuple on one party (advances its cursors): R, RT0, RT1 of n = rows*cols
int sh3h_trunc_tuple(sh3h* h, int party, uint64_t rows, uint64_t cols, uint64_t d, int64_t* R, int64_t* RT0, int64_t* RT1) {
    return h->run([&](int i) {
        if (i != party) return;
        Party& P = h->p[i];
        TruncationPair t = P.eval.getTruncationTuple(rows, cols, d);
        const uint64_t n = rows * cols;
        memcpy(R, t.mR.hostData(), n * 8);
        memcpy(RT0, t.mRTrunc.mShares[0].hostData(), n * 8);
        memcpy(RT1, t.mRTrunc.mShares[1].hostData(), n * 8);
    });
}

}  // extern "C"

// ---- aby3-ML linear regression (ml/Regression.h) on sf64<D16> ---------------------
namespace {
struct PartyEngine {
    Party& P;
    template <Decimal D>
    sf64Matrix<D> mul(const sf64Matrix<D>& l, const sf64Matrix<D>& r) {
        sf64Matrix<D> d;
        P.eval.asyncMul(P.rt, l, r, d).get();
        return d;
    }
    template <Decimal D>
    sf64Matrix<D> mulTruncate(const sf64Matrix<D>& l, const sf64Matrix<D>& r, u64 shift) {
        sf64Matrix<D> d;
        P.eval.asyncMul(P.rt, l, r, d, shift).get();
        return d;
    }
    Sh3Piecewise mLogistic;
    template <Decimal D>
    sf64Matrix<D> logisticFunc(const sf64Matrix<D>& Y) {
        aby3ML::setLogistic(mLogistic);
        sf64Matrix<D> out(Y.rows(), Y.cols());
        mLogistic.eval<D>(P.rt.noDependencies(), Y, out, P.eval).get();
        return out;
    }
};
}  // namespace

extern "C" {

int sh3h_linreg(sh3h* h, int x_id, int y_id, int w_id, const uint64_t* batch_idx, uint64_t iters,
                           uint64_t batch, double lr) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        // same layout cast the reference uses between si64Matrix and sf64Matrix<D> (Sh3Encryptor.h:214-219)
        auto& X = reinterpret_cast<sf64Matrix<D16>&>(*P.ints.at(x_id));
        auto& Y = reinterpret_cast<sf64Matrix<D16>&>(*P.ints.at(y_id));
        auto& W = reinterpret_cast<sf64Matrix<D16>&>(*P.ints.at(w_id));
        RegressionParam params{iters, batch, lr};
        std::vector<u64> idx(batch_idx, batch_idx + iters * batch);
        PartyEngine eng{P};
        SGD_Linear(params, eng, X, Y, W, idx);
    });
}

// The same training loop for parties that share one GPU, replayed as ONE CUDA graph per iteration (ml/SgdGraph.h).
// Issued from the calling thread; the party threads stay idle.  Result and PRNG cursors are those of sh3h_linreg.
int sh3h_linreg_graph(sh3h* h, int x_id, int y_id, int w_id, const uint64_t* batch_idx, uint64_t iters, uint64_t batch, double lr) {
    try {
        std::array<ColocatedSgdLinear<D16>::PartyRef, 3> P;
        for (int i = 0; i < 3; ++i) {
            Party& p = h->p[i];
            P[i] = {p.ctx.get(), &p.eval, &reinterpret_cast<sf64Matrix<D16>&>(*p.ints.at(x_id)),
                    &reinterpret_cast<sf64Matrix<D16>&>(*p.ints.at(y_id)), &reinterpret_cast<sf64Matrix<D16>&>(*p.ints.at(w_id))};
        }
        RegressionParam params{iters, batch, lr};
        std::vector<u64> idx(batch_idx, batch_idx + iters * batch);
        ColocatedSgdLinear<D16>::run(P, params, idx);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// The same as ONE persistent kernel for the whole run (csrc/sgd_fused.cu); shapes it does not take go to the graph.
int sh3h_linreg_fused(sh3h* h, int x_id, int y_id, int w_id, const uint64_t* batch_idx, uint64_t iters, uint64_t batch, double lr) {
    try {
        std::array<ColocatedSgdLinear<D16>::PartyRef, 3> P;
        for (int i = 0; i < 3; ++i) {
            Party& p = h->p[i];
            P[i] = {p.ctx.get(), &p.eval, &reinterpret_cast<sf64Matrix<D16>&>(*p.ints.at(x_id)),
                    &reinterpret_cast<sf64Matrix<D16>&>(*p.ints.at(y_id)), &reinterpret_cast<sf64Matrix<D16>&>(*p.ints.at(w_id))};
        }
        RegressionParam params{iters, batch, lr};
        std::vector<u64> idx(batch_idx, batch_idx + iters * batch);
        if (ColocatedSgdLinear<D16>::fusedSupports(batch, P[0].X->cols())) ColocatedSgdLinear<D16>::runFused(P, params, idx);
        else ColocatedSgdLinear<D16>::run(P, params, idx);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int sh3h_logreg(sh3h* h, int x_id, int y_id, int w_id, const uint64_t* batch_idx, uint64_t iters, uint64_t batch, double lr) {
    return h->run([&](int i) {
        Party& P = h->p[i];
        auto& X = reinterpret_cast<sf64Matrix<D16>&>(*P.ints.at(x_id));
        auto& Y = reinterpret_cast<sf64Matrix<D16>&>(*P.ints.at(y_id));
        auto& W = reinterpret_cast<sf64Matrix<D16>&>(*P.ints.at(w_id));
        RegressionParam params{iters, batch, lr};
        std::vector<u64> idx(batch_idx, batch_idx + iters * batch);
        PartyEngine eng{P};
        SGD_Logistic(params, eng, X, Y, W, idx);
    });
}

// ---- binary engine -------------------------------------------------------------
// Evaluate a circuit given as flat arrays (same layout as the oracle's orc_circuit) on
// sbMatrix inputs; returns output handles.  width = number of instances (rows).
int sh3h_bin_eval(sh3h* h, const uint32_t* gates, uint32_t gate_count, uint32_t wire_count,
                  const uint32_t* level_gates, uint32_t level_count,
                  const uint32_t* input_first, const uint32_t* input_bits, uint32_t num_inputs,
                  const uint32_t* output_off, const uint32_t* output_bits, const uint32_t* output_wires,
                  const uint8_t* output_invert, uint32_t num_outputs,
                  const int* input_ids, int* output_ids) {
    for (uint32_t k = 0; k < num_outputs; ++k) output_ids[k] = h->next_handle++;
    return h->run([&](int i) {
        Party& P = h->p[i];
        oc::BetaCircuit cir;
        cir.loadFlat(gates, gate_count, wire_count, level_gates, level_count, input_first, input_bits, num_inputs,
                     output_off, output_bits, output_wires, output_invert, num_outputs);
        Sh3BinaryEvaluator ev;
        const u64 width = P.bins.at(input_ids[0])->rows();
        ev.sharePlanes(P.rt.mComm);              // the inputs are sharings made by the encryptor
        ev.setCir(&cir, width, P.eval.mShareGen);
        for (uint32_t k = 0; k < num_inputs; ++k) ev.setInput(k, *P.bins.at(input_ids[k]));
        ev.asyncEvaluate(P.rt).get();
        for (uint32_t k = 0; k < num_outputs; ++k) {
            auto m = std::make_unique<sbMatrix>(width, output_bits[k]);
            ev.getOutput(k, *m);
            P.bins[output_ids[k]] = std::move(m);
        }
        P.ctx->sync();
    });
}

// Evaluate a circuit and then run the device-side shadow check (Sh3BinaryEvaluator::enableDebug / validateMemory) over
// separate debug channels.  tamper_wire >= 0: before the check, party `tamper_party` flips instance 0 of that wire in its
// own share plane (a fault the check must find).  *mismatches = instance-gates that disagree (summed over the parties'
// identical views / 3).  Returns 0 also when the check fails: the count is the result.
int sh3h_bin_eval_check(sh3h* h, const uint32_t* gates, uint32_t gate_count, uint32_t wire_count,
                        const uint32_t* level_gates, uint32_t level_count,
                        const uint32_t* input_first, const uint32_t* input_bits, uint32_t num_inputs,
                        const uint32_t* output_off, const uint32_t* output_bits, const uint32_t* output_wires,
                        const uint8_t* output_invert, uint32_t num_outputs,
                        const int* input_ids, int64_t tamper_wire, int tamper_party, uint64_t* mismatches) {
    // debug channels: a second, independent set of pipes between the same contexts
    auto d01 = oc::Channel::makePair(h->p[0].ctx.get(), h->p[1].ctx.get());
    auto d02 = oc::Channel::makePair(h->p[0].ctx.get(), h->p[2].ctx.get());
    auto d12 = oc::Channel::makePair(h->p[1].ctx.get(), h->p[2].ctx.get());
    oc::Channel dprev[3] = {d02.first, d01.second, d12.second}, dnext[3] = {d01.first, d12.first, d02.second};
    std::atomic<uint64_t> total{0};
    int rc = h->run([&](int i) {
        Party& P = h->p[i];
        oc::BetaCircuit cir;
        cir.loadFlat(gates, gate_count, wire_count, level_gates, level_count, input_first, input_bits, num_inputs,
                     output_off, output_bits, output_wires, output_invert, num_outputs);
        Sh3BinaryEvaluator ev;
        const u64 width = P.bins.at(input_ids[0])->rows();
        ev.setCir(&cir, width, P.eval.mShareGen);
        for (uint32_t k = 0; k < num_inputs; ++k) ev.setInput(k, *P.bins.at(input_ids[k]));
        ev.asyncEvaluate(P.rt).get();
        if (tamper_wire >= 0 && i == tamper_party) {
            u64 word = 0;
            u8* at = (u8*)ev.planeDevice(0) + (u64)tamper_wire * ev.rowBytes();
            gpu::check(aby3cu_d2h(P.ctx->h(), &word, at, 8));
            P.ctx->sync();
            word ^= 1;
            gpu::check(aby3cu_h2d(P.ctx->h(), at, &word, 8));
            P.ctx->sync();
        }
        ev.enableDebug(i, dprev[i], dnext[i]);
        try { ev.validateMemory(); } catch (const std::runtime_error&) {}
        total += ev.mDebugMismatches;
    });
    *mismatches = total.load();
    return rc;
}

// Same evaluation, but inputs and outputs travel through the engine as sPackedBin (bit-sliced):
// sbMatrix -> sPackedBin by a first engine pass, setInput(sPackedBin), getOutput(sPackedBin),
// and the packed output is transposed back to an sbMatrix for comparison.
int sh3h_bin_eval_packed(sh3h* h, const uint32_t* gates, uint32_t gate_count, uint32_t wire_count,
                         const uint32_t* level_gates, uint32_t level_count,
                         const uint32_t* input_first, const uint32_t* input_bits, uint32_t num_inputs,
                         const uint32_t* output_off, const uint32_t* output_bits, const uint32_t* output_wires,
                         const uint8_t* output_invert, uint32_t num_outputs,
                         const int* input_ids, int* output_ids) {
    for (uint32_t k = 0; k < num_outputs; ++k) output_ids[k] = h->next_handle++;
    return h->run([&](int i) {
        Party& P = h->p[i];
        oc::BetaCircuit cir;
        cir.loadFlat(gates, gate_count, wire_count, level_gates, level_count, input_first, input_bits, num_inputs,
                     output_off, output_bits, output_wires, output_invert, num_outputs);
        const u64 width = P.bins.at(input_ids[0])->rows();
        std::vector<sPackedBin> packed(num_inputs);
        {
            // conversion pass (no PRNG draw: explicit keys), wires read back bit-sliced
            Sh3BinaryEvaluator conv;
            conv.setCir(&cir, width, oc::ZeroBlock, oc::ZeroBlock);
            for (uint32_t k = 0; k < num_inputs; ++k) {
                conv.setInput(k, *P.bins.at(input_ids[k]));
                conv.getOutput(cir.mInputs[k].mWires, packed[k]);
            }
        }
        Sh3BinaryEvaluator ev;
        ev.sharePlanes(P.rt.mComm);
        ev.setCir(&cir, width, P.eval.mShareGen);
        for (uint32_t k = 0; k < num_inputs; ++k) ev.setInput(k, packed[k]);
        ev.asyncEvaluate(P.rt).get();
        for (uint32_t k = 0; k < num_outputs; ++k) {
            sPackedBin po;
            ev.getOutput(k, po);
            if (po.shareCount() != width || po.bitCount() != output_bits[k]) throw RTE_LOC;
            auto m = std::make_unique<sbMatrix>(width, output_bits[k]);
            for (int s = 0; s < 2; ++s) {
                i64* dst = m->mShares[s].devOut();
                gpu::check(aby3cu_memset(P.ctx->h(), dst, 0, m->i64Size() * 8));
                gpu::check(aby3cu_bit_transpose(P.ctx->h(), po.mShares[s].dev(), output_bits[k], width, po.simdWidth() * 8,
                                                dst, m->i64Cols() * 8, nullptr));
            }
            P.bins[output_ids[k]] = std::move(m);
        }
        P.ctx->sync();
    });
}

// ---- device timing over all three parties' streams -------------------------------
// begin: every party stream waits for one common start event; end: returns the time
// from that event until the last party stream has drained.  (Parties on one GPU.)
int sh3h_timer_begin(sh3h* h) {
    int rc = h->run([&](int i) { h->p[i].ctx->sync(); });
    if (rc) return rc;
    try {
        gpu::check(aby3cu_event_record(h->p[0].ctx->h(), h->p[0].ev_start));
        for (int i = 1; i < 3; ++i) gpu::check(aby3cu_event_wait(h->p[i].ctx->h(), h->p[0].ev_start));
        // work a party issues ahead on its second stream belongs to the timed region too
        for (int i = 0; i < 3; ++i) gpu::check(aby3cu_event_wait(h->p[i].ctx->aux()->h(), h->p[0].ev_start));
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
    return 0;
}
int sh3h_timer_end(sh3h* h, float* ms) {
    try {
        for (int i = 0; i < 3; ++i) gpu::check(aby3cu_event_record(h->p[i].ctx->h(), h->p[i].ev_stop));
        for (int i = 1; i < 3; ++i) gpu::check(aby3cu_event_wait(h->p[0].ctx->h(), h->p[i].ev_stop));
        gpu::check(aby3cu_event_record(h->p[0].ctx->h(), h->ev_end));
        gpu::check(aby3cu_event_sync(h->ev_end));
        gpu::check(aby3cu_event_elapsed_ms(h->p[0].ev_start, h->ev_end, ms));
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
    return 0;
}
int sh3h_sync(sh3h* h) { return h->run([&](int i) { h->p[i].ctx->sync(); }); }

// ---- interop with another CUDA user of the same process (bench.py: torch.distributed / NCCL collectives) -------------
// the six device pointers [party][plane] of an arithmetic sharing (valid in stream order on the owning party's stream)
int sh3h_device_ptrs(sh3h* h, int id, void* out[6]) {
    return h->run([&](int i) {
        auto& m = *h->p[i].ints.at(id);
        out[2 * i + 0] = (void*)m.mShares[0].devMut();
        out[2 * i + 1] = (void*)m.mShares[1].devMut();
    });
}
// an arithmetic sharing whose planes are allocated on the device but not written (the caller fills them through
// sh3h_device_ptrs, e.g. as the destination of a broadcast)
int sh3h_alloc_shares(sh3h* h, uint64_t rows, uint64_t cols) {
    const int id = h->next_handle++;
    int rc = h->run([&](int i) {
        auto m = std::make_unique<si64Matrix>(rows, cols);
        (void)m->mShares[0].devOut();
        (void)m->mShares[1].devOut();
        h->p[i].ints[id] = std::move(m);
    });
    return rc ? -1 : id;
}
// cudaStream_t of a party (wrap it, do not destroy it)
void* sh3h_party_stream(sh3h* h, int party) { return h->p[party].ctx->stream(); }
uint64_t sh3h_launch_count(sh3h* h) {
    uint64_t n = 0;
    for (int i = 0; i < 3; ++i) n += h->p[i].ctx->launchCount();
    return n;
}
// return every cached device block of the parties' pools to the driver
int sh3h_trim(sh3h* h) { return h->run([&](int i) { h->p[i].ctx->trim(); }); }

// ABY3_POOL_GUARD=1 self-test: party 0 takes a block from its pool, writes `overrun` bytes past its end (0 = a clean
// block) and gives it back to the driver.  Returns 1 when the pool's guard check threw, 0 when it passed, -1 when the
// guard mode is off.
int sh3h_guard_selftest(sh3h* h, uint64_t bytes, uint64_t overrun) {
    if (!gpu::Context::guardEnabled()) return -1;
    int caught = 0;
    h->run([&](int i) {
        if (i != 0) return;
        gpu::Context* c = h->p[0].ctx.get();
        c->trim();
        void* p = c->alloc(bytes);
        const size_t cls = gpu::Context::roundSize(bytes);
        gpu::check(aby3cu_memset(c->h(), p, 0x11, cls + overrun));
        c->release(p, bytes);
        try { c->trim(); } catch (const std::exception&) { caught = 1; }
    });
    return caught;
}
// driver allocations / frees behind the parties' buffer pools: [0] mallocs, [1] bytes, [2] frees
void sh3h_pool_stats(sh3h* h, uint64_t out[3]) {
    out[0] = out[1] = out[2] = 0;
    for (int i = 0; i < 3; ++i) { out[0] += h->p[i].ctx->mallocCount(); out[1] += h->p[i].ctx->mallocBytes(); out[2] += h->p[i].ctx->freeCount(); }
}
uint64_t sh3h_bytes_sent(sh3h* h) {
    uint64_t n = 0;
    for (int i = 0; i < 3; ++i) n += h->p[i].comm.mNext.getTotalDataSent() + h->p[i].comm.mPrev.getTotalDataSent();
    return n;
}

}  // extern "C"

// ---- circuit library export (host only; no device needed) -------------------------
// name: "and" | "or" | "xor" | "add" (ripple) | "add_depth" (prefix) | "add_msb" | "lt" | "eq"
extern "C" {

struct sh3h_circuit {
    oc::BetaLibrary lib;
    oc::BetaCircuit own;
    oc::BetaCircuit* cir = nullptr;
};

sh3h_circuit* sh3h_circuit_build(const char* name, uint32_t bits) {
    try {
        std::unique_ptr<sh3h_circuit> c(new sh3h_circuit);
        const std::string n(name);
        if (n == "and") c->cir = c->lib.int_int_bitwiseAnd(bits, bits, bits);
        else if (n == "or") c->cir = c->lib.int_int_bitwiseOr(bits, bits, bits);
        else if (n == "xor") c->cir = c->lib.int_int_bitwiseXor(bits, bits, bits);
        else if (n == "nor") c->cir = c->lib.bits_nor_helper(bits);
        else if (n == "add") c->cir = c->lib.int_int_add(bits, bits, bits, oc::BetaLibrary::Optimized::Size);
        else if (n == "add_depth") c->cir = c->lib.int_int_add(bits, bits, bits, oc::BetaLibrary::Optimized::Depth);
        else if (n == "add_msb") c->cir = c->lib.int_int_add_msb(bits);
        else if (n == "lt") c->cir = c->lib.int_int_lt_ab(bits, bits);
        else if (n == "lt_swapped") c->cir = c->lib.int_int_lt(bits, bits);
        else if (n == "sub") c->cir = c->lib.int_int_subtract(bits, bits, bits);
        else if (n == "eq") c->cir = c->lib.int_eq(bits);
        else if (n.rfind("piecewise", 0) == 0) c->cir = c->lib.int_Sh3Piecewise_helper(bits, std::stoul(n.substr(9)));
        else if (n == "a2b") {            // Sh3Converter::getArithToBinCircuit(64, bits), levelised as setCir does
            Sh3Converter conv;
            c->own = conv.getArithToBinCircuit(64, bits);
            c->own.levelByAndDepth();
            c->cir = &c->own;
        }
        else { g_err = "unknown circuit " + n; return nullptr; }
        return c.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void sh3h_circuit_free(sh3h_circuit* c) { delete c; }
// sizes: gates, wires, levels, inputs, outputs, total output wires, nonlinear gates
void sh3h_circuit_sizes(const sh3h_circuit* c, uint32_t s[7]) {
    const auto& cir = *c->cir;
    s[0] = (uint32_t)cir.mGates.size(); s[1] = cir.mWireCount; s[2] = (uint32_t)cir.mLevelCounts.size();
    s[3] = (uint32_t)cir.mInputs.size(); s[4] = (uint32_t)cir.mOutputs.size();
    uint32_t ow = 0;
    for (auto& o : cir.mOutputs) ow += (uint32_t)o.size();
    s[5] = ow; s[6] = (uint32_t)cir.mNonlinearGateCount;
}
void sh3h_circuit_copy(const sh3h_circuit* c, uint32_t* gates, uint32_t* level_gates, uint32_t* input_first,
                       uint32_t* input_bits, uint32_t* output_off, uint32_t* output_bits, uint32_t* output_wires,
                       uint8_t* output_invert) {
    const auto& cir = *c->cir;
    for (size_t g = 0; g < cir.mGates.size(); ++g) {
        gates[4 * g] = cir.mGates[g].mInput[0]; gates[4 * g + 1] = cir.mGates[g].mInput[1];
        gates[4 * g + 2] = cir.mGates[g].mOutput; gates[4 * g + 3] = (uint32_t)cir.mGates[g].mType;
    }
    for (size_t l = 0; l < cir.mLevelCounts.size(); ++l) level_gates[l] = (uint32_t)cir.mLevelCounts[l];
    for (size_t k = 0; k < cir.mInputs.size(); ++k) { input_first[k] = cir.mInputs[k].front(); input_bits[k] = (uint32_t)cir.mInputs[k].size(); }
    uint32_t off = 0;
    for (size_t k = 0; k < cir.mOutputs.size(); ++k) {
        output_off[k] = off; output_bits[k] = (uint32_t)cir.mOutputs[k].size();
        for (size_t b = 0; b < cir.mOutputs[k].size(); ++b) {
            output_wires[off + b] = cir.mOutputs[k][b];
            output_invert[off + b] = cir.isInvert(cir.mOutputs[k][b]) ? 1 : 0;
        }
        off += (uint32_t)cir.mOutputs[k].size();
    }
}

}  // extern "C"
