// compat/apps_driver.cpp -- C entry points over the REFERENCE's own application sources (aby3-ML: aby3ML.cpp,
// Regression.h, main-linear.cpp, LinearModelGen.cpp; aby3-Basic: BoolBasic.cpp, ArithBasic.cpp, BuildingBlocks.cpp,
// Sort.cpp, Basic.cpp, debug.cpp), compiled UNMODIFIED where they lie under /root/reference against compat/include --
// forwarding headers that resolve <aby3/sh3/*.h>, <cryptoTools/...> and <Eigen/Dense> to the B200 facade
// (aby3_b200/sh3).  This file is ours; it only sets up three party threads (as aby3-ML/main-linear.cpp:181-250 does)
// and moves share planes in and out.  tests/test_compat.py compares what comes out with the same sources running on
// the CPU (oracle/_ref) and with the facade's own paths.  Same entry-point shapes as oracle/ref_driver.cpp.
#include <aby3/sh3/Sh3BinaryEvaluator.h>
#include <aby3/sh3/Sh3Encryptor.h>
#include <aby3/sh3/Sh3Evaluator.h>
#include <aby3/sh3/Sh3Runtime.h>
#include <aby3-ML/aby3ML.h>
#include <aby3-ML/Regression.h>
#include <aby3-Basic/Basics.h>
#include <aby3-Basic/BuildingBlocks.h>
#include <aby3-Basic/Sort.h>
#include <cryptoTools/Common/CLP.h>
#include <aby3_tests/Test.h>
#include <fstream>
#include <map>
#include <cryptoTools/Network/IOService.h>

#include <atomic>
#include <chrono>
#include <thread>

using namespace aby3;

int linear_main_3pc_sh(oc::CLP& cmd);        // aby3-ML/main-linear.cpp:173

namespace {
thread_local std::string g_err;

struct CmpParty {
    gpu::Context* ctx = nullptr;
    CommPkg comm;
    Sh3Runtime rt;
    Sh3Encryptor enc;
    Sh3Evaluator eval;
};
block blk(const uint8_t* p) { block b; memcpy(b.data(), p, 16); return b; }

// the session set-up of aby3-ML/main-linear.cpp:187-204 for party i
void openSessions(oc::IOService& ios, const std::string& tag, u64 i, oc::Session& epPrev, oc::Session& epNext) {
    const u64 next = (i + 1) % 3, prev = (i + 2) % 3;
    auto name = [&](u64 a, u64 b) { return tag + std::to_string(std::min(a, b)) + std::to_string(std::max(a, b)); };
    epNext.start(ios, "127.0.0.1", 1212 + (u32)std::min(i, next), i < next ? oc::SessionMode::Server : oc::SessionMode::Client, name(i, next));
    epPrev.start(ios, "127.0.0.1", 1212 + (u32)std::min(i, prev), i < prev ? oc::SessionMode::Server : oc::SessionMode::Client, name(i, prev));
}
std::atomic<int> g_instance{0};
}  // namespace

struct cmp_session {
    oc::IOService ios;
    CmpParty p[3];
    int run(const std::function<void(int)>& f) {
        std::string errs[3];
        std::thread th[3];
        for (int i = 0; i < 3; ++i)
            th[i] = std::thread([&, i] {
                try {
                    if (p[i].ctx) gpu::setCurrent(p[i].ctx);
                    f(i);
                    if (p[i].ctx) p[i].ctx->sync();
                } catch (const std::exception& e) { errs[i] = e.what(); } catch (...) { errs[i] = "unknown exception"; }
            });
        for (auto& t : th) t.join();
        for (int i = 0; i < 3; ++i)
            if (!errs[i].empty()) { g_err = "party " + std::to_string(i) + ": " + errs[i]; return 1; }
        return 0;
    }
};

namespace {
void loadInt(si64Matrix& m, const int64_t* shares, int party, u64 rows, u64 cols) {
    m.resize(rows, cols);
    const u64 n = rows * cols;
    for (int s = 0; s < 2; ++s) memcpy(m.mShares[s].data(), shares + (u64(party) * 2 + s) * n, n * 8);
}
void storeInt(const si64Matrix& m, int64_t* shares, int party) {
    const u64 n = m.size();
    for (int s = 0; s < 2; ++s) memcpy(shares + (u64(party) * 2 + s) * n, m.mShares[s].data(), n * 8);
}
void loadBin(sbMatrix& m, const int64_t* shares, int party, u64 rows, u64 bits) {
    m.resize(rows, bits);
    const u64 n = m.i64Size();
    for (int s = 0; s < 2; ++s) memcpy(m.mShares[s].data(), shares + (u64(party) * 2 + s) * n, n * 8);
}
void storeBin(const sbMatrix& m, int64_t* shares, int party) {
    const u64 n = m.i64Size();
    for (int s = 0; s < 2; ++s) memcpy(shares + (u64(party) * 2 + s) * n, m.mShares[s].data(), n * 8);
}
double since(std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
}  // namespace

extern "C" {

const char* cmp_last_error(void) { return g_err.c_str(); }

// seeds: [party][0 = prev, 1 = next][16]; enc.init / eval.init as aby3_tests/Sh3EvaluatorTests.cpp:41-47
cmp_session* cmp_session_new(const uint8_t* enc_seeds, const uint8_t* eval_seeds) {
    std::unique_ptr<cmp_session> s(new cmp_session);
    const std::string tag = "cmp" + std::to_string(g_instance++) + "_";
    int rc = s->run([&](int i) {
        CmpParty& P = s->p[i];
        oc::Session epPrev, epNext;
        openSessions(s->ios, tag, i, epPrev, epNext);
        P.comm = CommPkg{epPrev.addChannel(), epNext.addChannel()};
        P.ctx = gpu::current();                       // created by the IOService for this thread
        P.rt.init(i, P.comm);
        P.enc.init(i, blk(enc_seeds + (2 * i) * 16), blk(enc_seeds + (2 * i + 1) * 16));
        P.eval.init(i, blk(eval_seeds + (2 * i) * 16), blk(eval_seeds + (2 * i + 1) * 16));
    });
    return rc ? nullptr : s.release();
}
void cmp_session_free(cmp_session* s) {
    if (!s) return;
    s->run([&](int) {});          // drains the three party streams
    delete s;
}

int cmp_share_bin(cmp_session* s, int owner, const int64_t* plain, int64_t* shares, uint64_t rows, uint64_t words) {
    return s->run([&](int i) {
        CmpParty& P = s->p[i];
        sbMatrix m(rows, words * 64);
        if (i == owner) {
            i64Matrix pl(rows, words);
            memcpy(pl.data(), plain, rows * words * 8);
            P.enc.localBinMatrix(P.comm, pl, m);
        } else P.enc.remoteBinMatrix(P.comm, m);
        storeBin(m, shares, i);
    });
}
int cmp_share_int(cmp_session* s, int owner, const int64_t* plain, int64_t* shares, uint64_t rows, uint64_t cols) {
    return s->run([&](int i) {
        CmpParty& P = s->p[i];
        si64Matrix m(rows, cols);
        if (i == owner) {
            i64Matrix pl(rows, cols);
            memcpy(pl.data(), plain, rows * cols * 8);
            P.enc.localIntMatrix(P.comm, pl, m);
        } else P.enc.remoteIntMatrix(P.comm, m);
        storeInt(m, shares, i);
    });
}
int cmp_reveal_all(cmp_session* s, const int64_t* shares, uint64_t rows, uint64_t cols, int binary, int64_t* out) {
    return s->run([&](int i) {
        CmpParty& P = s->p[i];
        i64Matrix dest(rows, cols);
        if (!binary) { si64Matrix m; loadInt(m, shares, i, rows, cols); P.enc.revealAll(P.comm, m, dest); }
        else { sbMatrix m; loadBin(m, shares, i, rows, cols * 64); P.enc.revealAll(P.comm, m, dest); }
        memcpy(out + u64(i) * rows * cols, dest.data(), rows * cols * 8);
    });
}

// aby3-Basic on binary sharings of 64-bit values: same op codes as ref_basic_bool (oracle/ref_driver.cpp)
int cmp_basic_bool(cmp_session* s, int op, const int64_t* A, const int64_t* B, uint64_t n, int64_t* out, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        CmpParty& P = s->p[i];
        sbMatrix a, b, r;
        loadBin(a, A, i, n, 64);
        loadBin(b, B, i, n, 64);
        auto t0 = std::chrono::steady_clock::now();
        switch (op) {
        case 0: bool_cipher_lt(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 1: bool_cipher_eq(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 2: bool_cipher_and(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 3: bool_cipher_or(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 4: bool_cipher_add(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 5: bool_cipher_max(i, a, b, r, P.enc, P.eval, P.rt); break;
        case 6: bool_cipher_min(i, a, b, r, P.enc, P.eval, P.rt); break;
        default: throw std::runtime_error("cmp_basic_bool: unknown op");
        }
        P.ctx->sync();
        t[i] = since(t0);
        if (r.rows() != n) throw std::runtime_error("cmp_basic_bool: unexpected result shape");
        storeBin(r, out, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
int cmp_basic_cipher_gt(cmp_session* s, const int64_t* A, const int64_t* B, uint64_t n, int64_t* out, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        CmpParty& P = s->p[i];
        si64Matrix a, b;
        sbMatrix r;
        loadInt(a, A, i, n, 1);
        loadInt(b, B, i, n, 1);
        auto t0 = std::chrono::steady_clock::now();
        cipher_gt(i, a, b, r, P.eval, P.rt);
        P.ctx->sync();
        t[i] = since(t0);
        if (r.rows() != n) throw std::runtime_error("cmp_basic_cipher_gt: unexpected result shape");
        storeBin(r, out, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
int cmp_basic_max_min_split(cmp_session* s, const int64_t* A, const int64_t* B, uint64_t n, int64_t* mx, int64_t* mn, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        CmpParty& P = s->p[i];
        sbMatrix a, b, hi, lo;
        loadBin(a, A, i, n, 64);
        loadBin(b, B, i, n, 64);
        auto t0 = std::chrono::steady_clock::now();
        bool_cipher_max_min_split(i, a, b, hi, lo, P.enc, P.eval, P.rt);
        P.ctx->sync();
        t[i] = since(t0);
        storeBin(hi, mx, i);
        storeBin(lo, mn, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
int cmp_basic_odd_even_merge(cmp_session* s, const int64_t* A, uint64_t n1, const int64_t* B, uint64_t n2, int64_t* out, double* secs) {
    double t[3] = {0, 0, 0};
    int rc = s->run([&](int i) {
        CmpParty& P = s->p[i];
        sbMatrix a, b, r;
        loadBin(a, A, i, n1, 64);
        loadBin(b, B, i, n2, 64);
        auto t0 = std::chrono::steady_clock::now();
        odd_even_merge(a, b, r, i, P.enc, P.eval, P.rt);
        P.ctx->sync();
        t[i] = since(t0);
        if (r.rows() != n1 + n2) throw std::runtime_error("cmp_basic_odd_even_merge: unexpected result shape");
        storeBin(r, out, i);
    });
    if (secs) *secs = std::max(t[0], std::max(t[1], t[2]));
    return rc;
}
// cipher_mul on arithmetic sharings (BuildingBlocks.cpp:215-228: this fork's element-wise product of n x 1 vectors)
int cmp_basic_cipher_mul(cmp_session* s, const int64_t* A, const int64_t* B, uint64_t n, int64_t* out) {
    return s->run([&](int i) {
        CmpParty& P = s->p[i];
        si64Matrix a, b, r(n, 1);
        loadInt(a, A, i, n, 1);
        loadInt(b, B, i, n, 1);
        cipher_mul(i, a, b, r, P.eval, P.enc, P.rt);
        storeInt(r, out, i);
    });
}

// aby3-ML/main-linear.cpp's own entry point on the facade, argv-style
int cmp_main_linear(int argc, const char* const* argv) {
    try {
        oc::CLP cmd;
        cmd.parse(argc, argv);
        return linear_main_3pc_sh(cmd);
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// aby3-ML linear regression as main-linear.cpp:31-133 sets it up (aby3ML engine seeded toBlock(pIdx), D16, party 0 inputs
// everything), data passed in; SGD_Linear = the reference's Regression.h:112-184, every product a facade asyncMul.
double cmp_sgd_linear(const double* x, const double* y, uint64_t N, uint64_t F, uint64_t B, uint64_t iters, double lr, int64_t* w_shares) {
    const std::string tag = "sgd" + std::to_string(g_instance++) + "_";
    oc::IOService ios;
    std::string errs[3];
    double secs[3] = {0, 0, 0};
    std::thread th[3];
    for (int i = 0; i < 3; ++i)
        th[i] = std::thread([&, i] {
            try {
                oc::Session epPrev, epNext;
                openSessions(ios, tag, i, epPrev, epNext);
                const Decimal D = D16;
                {
                aby3ML p;
                p.mPrint = false;
                p.init(i, epPrev, epNext, oc::toBlock(i));
                sf64Matrix<D> X, Y, W;
                if (i == 0) {
                    eMatrix<double> vx(N, F), vy(N, 1), vw(F, 1);
                    memcpy(vx.data(), x, N * F * sizeof(double));
                    memcpy(vy.data(), y, N * sizeof(double));
                    vw.setZero();
                    X = p.localInput<D>(vx); Y = p.localInput<D>(vy); W = p.localInput<D>(vw);
                } else {
                    X = p.remoteInput<D>(0); Y = p.remoteInput<D>(0); W = p.remoteInput<D>(0);
                }
                RegressionParam params;
                params.mBatchSize = B; params.mIterations = iters; params.mLearningRate = lr;
                gpu::current()->sync();
                auto t0 = std::chrono::steady_clock::now();
                SGD_Linear(params, p, X, Y, W);
                gpu::current()->sync();
                secs[i] = since(t0);
                if (w_shares) storeInt(W.i64Cast(), w_shares, i);
                }
                gpu::current()->sync();
                gpu::setCurrent(nullptr);
            } catch (const std::exception& e) { errs[i] = e.what(); } catch (...) { errs[i] = "unknown exception"; }
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < 3; ++i)
        if (!errs[i].empty()) { g_err = "party " + std::to_string(i) + ": " + errs[i]; return -1.0; }
    return std::max(secs[0], std::max(secs[1], secs[2]));
}

}  // extern "C"

// ---- the fork's own role tests (aby3_tests/Test.cpp, BoolTest.cpp, SortTest.cpp; frontend/main.cpp:16-54 dispatches them
// ---- by flag, Eval/dis_exec.sh starts one process per role).  Here: the named test, compiled unmodified, on three
// ---- threads of this process with "-role i"; its check_result() lines are counted from DEBUG_FILE.
namespace {
typedef int (*RoleTest)(oc::CLP&);
RoleTest find_role_test(const std::string& n) {
    static const std::map<std::string, RoleTest> t = {
        {"arith_basic_test", arith_basic_test}, {"bool_basic_test", bool_basic_test}, {"bool_basic_test2", bool_basic_test2},
        {"bool_aggregation_test", bool_aggregation_test}, {"get_first_zero_test", get_first_zero_test},
        {"share_conversion_test", share_conversion_test}, {"initialization_test", initialization_test},
        {"bc_sort_test", bc_sort_test}, {"bc_sort_corner_test", bc_sort_corner_test}, {"bc_sort_multiple_times", bc_sort_multiple_times},
        {"quick_sort_test", quick_sort_test}, {"quick_sort_with_duplicate_elements_test", quick_sort_with_duplicate_elements_test},
        {"odd_even_merge_test", odd_even_merge_test}, {"shuffle_test", shuffle_test}, {"correlation_test", correlation_test}};
    auto it = t.find(n);
    return it == t.end() ? nullptr : it->second;
}
}  // namespace

extern "C" int cmp_role_test(const char* name, int* n_success, int* n_error) {
    RoleTest fn = find_role_test(name);
    if (!fn) { g_err = std::string("unknown role test ") + name; return 1; }
    { std::ofstream trunc(DEBUG_FILE, std::ios_base::trunc); }
    std::string errs[3];
    std::thread th[3];
    for (int i = 0; i < 3; ++i)
        th[i] = std::thread([&, i] {
            try {
                const std::string role = std::to_string(i);
                const char* argv[] = {"frontend", "-role", role.c_str()};
                oc::CLP cmd;
                cmd.parse(3, argv);
                fn(cmd);
                if (gpu::currentSlot()) gpu::setCurrent(nullptr);
            } catch (const std::exception& e) { errs[i] = e.what(); } catch (...) { errs[i] = "unknown exception"; }
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < 3; ++i)
        if (!errs[i].empty()) { g_err = "party " + std::to_string(i) + ": " + errs[i]; return 1; }
    int ok = 0, bad = 0;
    std::ifstream in(DEBUG_FILE);
    for (std::string line; std::getline(in, line);) {
        if (line.find("SUCCESS") != std::string::npos) ++ok;
        if (line.find("ERROR") != std::string::npos) ++bad;
    }
    if (n_success) *n_success = ok;
    if (n_error) *n_error = bad;
    return 0;
}
