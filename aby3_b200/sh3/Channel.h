// Channel.h -- the point-to-point channel the sh3 protocols talk through.
// The reference uses cryptoTools' oc::Channel over Boost.Asio TCP even when all
// three parties live in one process (aby3_tests/Sh3EvaluatorTests.cpp:23-36).
// Here a channel has one of two transports:
//
//  * local  -- the parties are host threads of one process.  A message is either
//    a small host byte string or a device buffer handed over with the CUDA event
//    that marks it complete; the receiver's stream waits on the event and copies
//    (same GPU: D2D; different GPUs: cudaMemcpyPeerAsync over NVLink).  No host
//    synchronisation on the device path.
//  * NCCL   -- parties on different GPUs: ncclSend / ncclRecv over NVLink on the
//    party's own stream (NcclEndpoint below).  Sends and receives are collected
//    and issued as ONE ncclGroup per flush, because a ring step (send to next,
//    receive from prev) deadlocks if the two are separate kernels on one stream.
//    Flush points: waiting on a receive, and the end of every runtime task.
//
// Method subset = what aby3/sh3, aby3-ML and aby3-Basic call:
// asyncSendCopy / asyncSend / send / recv / asyncRecv (future) / getTotalDataSent
// / resetStats, plus the device variants used by the facade itself.
#pragma once
#include <dlfcn.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>

#include "Gpu.h"

namespace oc {

namespace detail {

struct Message {
    std::vector<u8> host;                 // host payload (when dev is empty)
    aby3::gpu::Buffer dev;                // device payload (staging buffer owned by the sender's pool)
    void* ready = nullptr;                // event: payload complete on the sender's stream
    size_t bytes = 0;
};

// one direction of a local channel
struct Pipe {
    std::mutex mtx;
    std::condition_variable cv;
    std::deque<Message> q;
    std::atomic<u64> count{0};
    void push(Message&& m) {
        {
            std::lock_guard<std::mutex> g(mtx);
            q.push_back(std::move(m));
            count.fetch_add(1, std::memory_order_release);
        }
        cv.notify_all();
    }
    Message pop() {
        // latency-bound protocols (SGD: two rounds per iteration) wait for a peer thread that is
        // microseconds away: poll briefly before paying for a futex sleep / wake-up
        for (int spin = 0; spin < 20000 && count.load(std::memory_order_acquire) == 0; ++spin) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        std::unique_lock<std::mutex> l(mtx);
        cv.wait(l, [&] { return !q.empty(); });
        count.fetch_sub(1, std::memory_order_relaxed);
        Message m = std::move(q.front());
        q.pop_front();
        return m;
    }
};

// receive operations posted on one endpoint complete strictly in posting order,
// whichever future is waited on first
struct RecvQueue {
    std::mutex mtx;
    u64 posted = 0, done = 0;
    std::deque<std::function<void()>> ops;
    u64 post(std::function<void()> f) {
        std::lock_guard<std::mutex> g(mtx);
        ops.push_back(std::move(f));
        return ++posted;
    }
    void completeUpTo(u64 ticket) {
        for (;;) {
            std::function<void()> f;
            {
                std::lock_guard<std::mutex> g(mtx);
                if (done >= ticket) return;
                f = std::move(ops.front());
                ops.pop_front();
                ++done;
            }
            f();
        }
    }
    void drain() {
        u64 t;
        { std::lock_guard<std::mutex> g(mtx); t = posted; }
        completeUpTo(t);
    }
};

// transport behind a Channel endpoint
struct Transport {
    virtual ~Transport() = default;
    virtual void sendHost(const u8* p, size_t n) = 0;
    virtual void sendDevice(const void* d, size_t n) = 0;       // copy semantics
    // Receives are two-phase: post() registers the operation (NCCL: queues it so that it
    // joins the same group as the sends of this protocol step) and returns the completion,
    // which blocks the host until the bytes are in p / ordered on the stream for d.
    virtual std::function<void()> postRecvHost(u8* p, size_t n) = 0;
    virtual std::function<void()> postRecvDevice(void* d, size_t n) = 0;
    virtual void flush() {}
    virtual aby3::gpu::Context* context() const = 0;
};

struct LocalTransport : Transport {
    std::shared_ptr<Pipe> out, in;
    aby3::gpu::Context* ctx = nullptr;
    aby3::gpu::Context* context() const override { return ctx; }
    void requireCtx() const {
        if (!ctx) throw std::runtime_error("Channel: device transfer on a channel without a device context " LOCATION);
    }
    void sendHost(const u8* p, size_t n) override {
        Message m;
        m.bytes = n;
        m.host.assign(p, p + n);
        out->push(std::move(m));
    }
    std::function<void()> postRecvHost(u8* p, size_t n) override { return [this, p, n] { recvHost(p, n); }; }
    std::function<void()> postRecvDevice(void* d, size_t n) override { return [this, d, n] { recvDevice(d, n); }; }
    void recvHost(u8* p, size_t n) {
        Message m = in->pop();
        if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
        if (m.dev) {
            // device message consumed on the host (e.g. reveal into an i64Matrix)
            requireCtx();
            aby3::gpu::check(aby3cu_event_wait(ctx->h(), m.ready));
            if (n) aby3::gpu::check(aby3cu_d2h(ctx->h(), p, m.dev.ptr(), n));
            ctx->sync();
            aby3cu_event_destroy(m.ready);
            m.dev.free();
        } else if (n) {
            memcpy(p, m.host.data(), n);
        }
    }
    void sendDevice(const void* d, size_t n) override {
        requireCtx();
        Message m;
        m.bytes = n;
        m.dev.reset(ctx, std::max<size_t>(n, 16));
        if (n) aby3::gpu::check(aby3cu_d2d(ctx->h(), m.dev.ptr(), ctx->device(), d, ctx->device(), n));
        aby3::gpu::check(aby3cu_event_create(ctx->h(), &m.ready));
        aby3::gpu::check(aby3cu_event_record(ctx->h(), m.ready));
        out->push(std::move(m));
    }
    void recvDevice(void* d, size_t n) {
        requireCtx();
        Message m = in->pop();
        if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
        if (m.dev) {
            aby3::gpu::check(aby3cu_event_wait(ctx->h(), m.ready));
            if (n) aby3::gpu::check(aby3cu_d2d(ctx->h(), d, ctx->device(), m.dev.ptr(), m.dev.ctx()->device(), n));
            aby3cu_event_destroy(m.ready);
            // the staging buffer returns to the sender's pool once OUR copy has run
            void* done = nullptr;
            aby3::gpu::check(aby3cu_event_create(ctx->h(), &done));
            aby3::gpu::check(aby3cu_event_record(ctx->h(), done));
            m.dev.free(done);
        } else if (n) {
            aby3::gpu::check(aby3cu_h2d(ctx->h(), d, m.host.data(), n));
            ctx->sync();   // m.host dies with this scope
        }
    }
};

// ---- NCCL, loaded lazily so that nothing depends on libnccl unless it is used -----
struct NcclApi {
    typedef void* comm_t;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, void*) = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    static NcclApi& get() {
        static NcclApi api = load();
        return api;
    }
    void check(int rc, const char* what) const {
        if (rc != 0) throw std::runtime_error(std::string("NCCL ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "error"));
    }
private:
    static NcclApi load() {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw std::runtime_error(std::string("cannot load libnccl: ") + dlerror());
        auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) throw std::runtime_error(std::string("libnccl lacks ") + n); return p; };
        a.CommInitAll = (int (*)(comm_t*, int, const int*))sym("ncclCommInitAll");
        a.CommDestroy = (int (*)(comm_t))sym("ncclCommDestroy");
        a.Send = (int (*)(const void*, size_t, int, int, comm_t, void*))sym("ncclSend");
        a.Recv = (int (*)(void*, size_t, int, int, comm_t, void*))sym("ncclRecv");
        a.GroupStart = (int (*)())sym("ncclGroupStart");
        a.GroupEnd = (int (*)())sym("ncclGroupEnd");
        a.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        return a;
    }
};

// One party's NCCL state, shared by its two channels: pending operations are issued
// together as one group (ring steps need send and receive in the same group).
struct NcclEndpoint {
    NcclApi::comm_t comm = nullptr;
    aby3::gpu::Context* ctx = nullptr;
    struct Op { bool send; int peer; void* ptr; size_t bytes; };
    std::vector<Op> pending;
    std::vector<aby3::gpu::Buffer> staging;       // send copies, released after the flush
    void flush() {
        if (pending.empty()) return;
        auto& api = NcclApi::get();
        void* stream = aby3cu_ctx_stream(ctx->h());
        api.check(api.GroupStart(), "GroupStart");
        for (auto& op : pending) {
            if (op.send) api.check(api.Send(op.ptr, op.bytes, /*ncclUint8*/ 1, op.peer, comm, stream), "Send");
            else api.check(api.Recv(op.ptr, op.bytes, /*ncclUint8*/ 1, op.peer, comm, stream), "Recv");
        }
        api.check(api.GroupEnd(), "GroupEnd");
        pending.clear();
        staging.clear();                           // stream-ordered reuse: the sends above run first
    }
};

struct NcclTransport : Transport {
    std::shared_ptr<NcclEndpoint> ep;
    int peer = -1;
    aby3::gpu::Context* context() const override { return ep->ctx; }
    void sendDevice(const void* d, size_t n) override {
        if (!n) return;
        aby3::gpu::Buffer b(ep->ctx, n);
        aby3::gpu::check(aby3cu_d2d(ep->ctx->h(), b.ptr(), ep->ctx->device(), d, ep->ctx->device(), n));
        ep->pending.push_back({true, peer, b.ptr(), n});
        ep->staging.push_back(std::move(b));
    }
    std::function<void()> postRecvDevice(void* d, size_t n) override {
        if (n) ep->pending.push_back({false, peer, d, n});
        auto e = ep;
        return [e] { e->flush(); };
    }
    void sendHost(const u8* p, size_t n) override {
        if (!n) return;
        aby3::gpu::Buffer b(ep->ctx, n);
        aby3::gpu::check(aby3cu_h2d(ep->ctx->h(), b.ptr(), p, n));
        ep->ctx->sync();                           // p may be a temporary
        ep->pending.push_back({true, peer, b.ptr(), n});
        ep->staging.push_back(std::move(b));
    }
    std::function<void()> postRecvHost(u8* p, size_t n) override {
        auto b = std::make_shared<aby3::gpu::Buffer>(ep->ctx, std::max<size_t>(n, 16));
        if (n) ep->pending.push_back({false, peer, b->ptr(), n});
        auto e = ep;
        return [e, b, p, n] {
            e->flush();
            if (n) aby3::gpu::check(aby3cu_d2h(e->ctx->h(), p, b->ptr(), n));
            e->ctx->sync();
        };
    }
    void flush() override { ep->flush(); }
};

}  // namespace detail

class Channel {
public:
    Channel() = default;

    // two connected local endpoints; ctxA / ctxB are the device contexts of the parties
    // holding each end (may be null for host-only use, e.g. scheduler tests)
    static std::pair<Channel, Channel> makePair(aby3::gpu::Context* ctxA = nullptr, aby3::gpu::Context* ctxB = nullptr) {
        auto ab = std::make_shared<detail::Pipe>();
        auto ba = std::make_shared<detail::Pipe>();
        auto ta = std::make_shared<detail::LocalTransport>();
        auto tb = std::make_shared<detail::LocalTransport>();
        ta->out = ab; ta->in = ba; ta->ctx = ctxA;
        tb->out = ba; tb->in = ab; tb->ctx = ctxB;
        return {Channel(ta), Channel(tb)};
    }
    // NCCL endpoint towards rank `peer` of the party's communicator
    static Channel makeNccl(std::shared_ptr<detail::NcclEndpoint> ep, int peer) {
        auto t = std::make_shared<detail::NcclTransport>();
        t->ep = std::move(ep);
        t->peer = peer;
        return Channel(t);
    }

    bool isConnected() const { return (bool)mT; }
    aby3::gpu::Context* context() const { return mT ? mT->context() : nullptr; }
    // issue everything this party has queued on an NCCL transport (no-op for local channels)
    void flush() { if (mT) mT->flush(); }

    // ------------------------------------------------------------- host send --
    template <typename T>
    void asyncSendCopy(const T* p, u64 n) { sendBytes(reinterpret_cast<const u8*>(p), n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value>::type
    asyncSendCopy(const T& v) { sendBytes(reinterpret_cast<const u8*>(&v), sizeof(T)); }
    template <typename T>
    void asyncSendCopy(const span<T>& s) { sendBytes(reinterpret_cast<const u8*>(s.data()), s.size() * sizeof(T)); }
    template <typename T>
    void asyncSendCopy(const std::vector<T>& v) { sendBytes(reinterpret_cast<const u8*>(v.data()), v.size() * sizeof(T)); }
    // both transports copy eagerly, so the no-copy forms are aliases
    template <typename T>
    void asyncSend(const T* p, u64 n) { asyncSendCopy(p, n); }
    template <typename T>
    void asyncSend(std::vector<T>&& v) { asyncSendCopy(v); v.clear(); }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value>::type
    asyncSend(const T& v) { asyncSendCopy(v); }
    template <typename T>
    void send(const T* p, u64 n) { asyncSendCopy(p, n); }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value>::type
    send(const T& v) { asyncSendCopy(v); }

    // ------------------------------------------------------------- host recv --
    template <typename T>
    void recv(T* p, u64 n) { recvBytes(reinterpret_cast<u8*>(p), n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value>::type
    recv(T& v) { recvBytes(reinterpret_cast<u8*>(&v), sizeof(T)); }
    template <typename T>
    void recv(std::vector<T>& v) { recvBytes(reinterpret_cast<u8*>(v.data()), v.size() * sizeof(T)); }

    template <typename T>
    std::future<void> asyncRecv(T* p, u64 n) {
        require();
        return post(mT->postRecvHost(reinterpret_cast<u8*>(p), n * sizeof(T)));
    }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value, std::future<void>>::type
    asyncRecv(T& v) {
        require();
        return post(mT->postRecvHost(reinterpret_cast<u8*>(&v), sizeof(T)));
    }
    template <typename T>
    std::future<void> asyncRecv(std::vector<T>& v) {
        require();
        return post(mT->postRecvHost(reinterpret_cast<u8*>(v.data()), v.size() * sizeof(T)));
    }

    // ---------------------------------------------------------- device path ---
    // Send `bytes` of device memory that is (or will be, in stream order) valid on
    // this party's stream.  Copy semantics: the caller may overwrite d_src right
    // after the call (asyncSendCopy at Sh3Evaluator.cpp:109, :681-684).
    void asyncSendDevice(const void* d_src, size_t bytes) {
        require();
        mBytesSent += bytes;
        mT->sendDevice(d_src, bytes);
    }
    // Receive into device memory: blocks the HOST at most until the peer has posted
    // the message; the transfer itself is ordered on this party's stream.
    void recvDevice(void* d_dst, size_t bytes) {
        require();
        auto done = mT->postRecvDevice(d_dst, bytes);
        mRecvQ->drain();
        done();
    }
    std::future<void> asyncRecvDevice(void* d_dst, size_t bytes) {
        require();
        return post(mT->postRecvDevice(d_dst, bytes));
    }

    // ------------------------------------------------------------------ stats -
    u64 getTotalDataSent() const { return mBytesSent; }
    void resetStats() { mBytesSent = 0; }

private:
    explicit Channel(std::shared_ptr<detail::Transport> t) : mT(std::move(t)), mRecvQ(std::make_shared<detail::RecvQueue>()) {}
    void require() const {
        if (!mT) throw std::runtime_error("Channel: not connected " LOCATION);
    }
    std::future<void> post(std::function<void()> f) {
        if (!mRecvQ) throw std::runtime_error("Channel: not connected " LOCATION);
        auto q = mRecvQ;
        const u64 ticket = q->post(std::move(f));
        return std::async(std::launch::deferred, [q, ticket] { q->completeUpTo(ticket); });
    }
    void sendBytes(const u8* p, size_t n) {
        require();
        mBytesSent += n;
        mT->sendHost(p, n);
    }
    void recvBytes(u8* p, size_t n) {
        require();
        auto done = mT->postRecvHost(p, n);
        mRecvQ->drain();
        done();
    }

    std::shared_ptr<detail::Transport> mT;
    std::shared_ptr<detail::RecvQueue> mRecvQ;
    u64 mBytesSent = 0;
};

}  // namespace oc
