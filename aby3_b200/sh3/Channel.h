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
    std::vector<u8> host;                 // host payload (when dev / shared are empty)
    aby3::gpu::Buffer dev;                // device payload (staging buffer owned by the sender's pool)
    std::shared_ptr<aby3::gpu::SharedBuffer> shared;   // device payload sent without a copy
    size_t offset = 0;                    // payload starts this many bytes into `shared`
    void* ready = nullptr;                // event: payload complete on the sender's stream (null: same stream)
    int readyDevice = 0;
    size_t bytes = 0;
};

// What a zero-copy receive hands to the protocol: a pointer that is valid (in stream order) on the
// receiver's stream, and the object keeping it alive.  release() after the last kernel reading it
// has been enqueued.
struct Borrowed {
    const void* ptr = nullptr;
    size_t bytes = 0;
    aby3::gpu::Buffer own;                                  // staging buffer taken over from the message, or an own copy
    std::shared_ptr<aby3::gpu::SharedBuffer> shared;        // sender-shared buffer
    bool ownedByPeer = false;                               // `own` belongs to the sender's pool
    void release(aby3::gpu::Context* reader) {
        if (shared) {
            if (shared->ctx()->stream() != reader->stream()) shared->addReader(reader->device(), reader->recordEvent());
            shared.reset();
        }
        if (own) {
            if (ownedByPeer && own.ctx()->stream() != reader->stream()) own.free(reader->recordEvent(), reader->device());
            else own.free();
        }
        ptr = nullptr;
    }
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
    // zero-copy forms; the defaults fall back to the copying ones
    virtual void sendDeviceShared(const std::shared_ptr<aby3::gpu::SharedBuffer>& b, size_t n) { sendDevice(b->ptr(), n); }
    // `on` (may be null = the party's own stream): the stream that performs a cross-device copy, so that two incoming
    // messages from two different GPUs travel side by side instead of one after the other
    virtual std::function<void()> postRecvDeviceBorrow(size_t n, Borrowed* out, aby3::gpu::Context* on = nullptr) {
        (void)on;
        out->own.reset(context(), std::max<size_t>(n, 16));
        out->ptr = out->own.ptr();
        out->bytes = n;
        return postRecvDevice(out->own.ptr(), n);
    }
    // The same on ANOTHER stream of the party (`on`, e.g. gpu::Context::comm()): a slice of a buffer the protocol never
    // writes again leaves while the party's own stream keeps computing (the opened xy - r, block by block).
    virtual void sendDeviceSharedOn(aby3::gpu::Context* on, const std::shared_ptr<aby3::gpu::SharedBuffer>& b, size_t off, size_t n) = 0;
    virtual std::function<void()> postRecvDeviceOn(aby3::gpu::Context* on, void* d, size_t n) = 0;
    virtual void flushOn(aby3::gpu::Context*) {}
    virtual void flush() {}
    // both ends are device contexts on the SAME GPU (known when the pair is made, so both ends always agree)
    virtual bool colocated() const { return false; }
    virtual aby3::gpu::Context* context() const = 0;
    // late binding of the endpoint's device context (a Session hands out channels before it knows which party
    // thread -- and so which device -- will hold each end)
    virtual void bind(aby3::gpu::Context*) { throw std::runtime_error("Channel: this transport has a fixed context " LOCATION); }
};

struct LocalTransport : Transport {
    std::shared_ptr<Pipe> out, in;
    aby3::gpu::Context* ctx = nullptr;
    void* peerStream = nullptr;           // the stream of the context at the other end
    aby3::gpu::Context* context() const override { return ctx; }
    void bind(aby3::gpu::Context* c) override { ctx = c; }
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
        if (m.dev || m.shared) {
            // device message consumed on the host (e.g. reveal into an i64Matrix)
            requireCtx();
            waitReady(m);
            if (n) aby3::gpu::check(aby3cu_d2h(ctx->h(), p, m.dev ? m.dev.ptr() : (const u8*)m.shared->ptr() + m.offset, n));
            ctx->sync();
            m.dev.free();
            m.shared.reset();
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
        markReady(m);
        out->push(std::move(m));
    }
    void sendDeviceShared(const std::shared_ptr<aby3::gpu::SharedBuffer>& b, size_t n) override {
        requireCtx();
        if (b->ctx() != ctx) throw std::runtime_error("Channel: a shared buffer must come from the sender's own context " LOCATION);
        Message m;
        m.bytes = n;
        m.shared = b;
        markReady(m);
        out->push(std::move(m));
    }
    void sendDeviceSharedOn(aby3::gpu::Context* on, const std::shared_ptr<aby3::gpu::SharedBuffer>& b, size_t off, size_t n) override {
        requireCtx();
        if (b->ctx() != ctx) throw std::runtime_error("Channel: a shared buffer must come from the sender's own context " LOCATION);
        Message m;
        m.bytes = n;
        m.shared = b;
        m.offset = off;
        m.ready = on->recordEvent();              // everything `on` was ordered behind (the block's GEMM) precedes it
        m.readyDevice = on->device();
        out->push(std::move(m));
    }
    std::function<void()> postRecvDeviceOn(aby3::gpu::Context* on, void* d, size_t n) override {
        return [this, on, d, n] {
            requireCtx();
            Message m = in->pop();
            if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
            deliver(m, d, n, on);
        };
    }
    // Parties whose contexts share ONE stream are ordered by enqueue order alone: the receiver's host thread
    // only learns of a message after the sender has enqueued its producer, so no event is needed.
    bool sameStream() const { return peerStream && peerStream == ctx->stream(); }
    bool colocated() const override { return ctx && peerStream != nullptr; }
    void markReady(Message& m) {
        if (sameStream()) return;
        m.ready = ctx->recordEvent();
        m.readyDevice = ctx->device();
    }
    void waitReady(Message& m) {
        if (!m.ready) return;
        aby3::gpu::check(aby3cu_event_wait(ctx->h(), m.ready));
        aby3::gpu::EventPool::put(m.readyDevice, m.ready);
        m.ready = nullptr;
    }
    std::function<void()> postRecvDeviceBorrow(size_t n, Borrowed* b, aby3::gpu::Context* on = nullptr) override {
        return [this, n, b, on] {
            requireCtx();
            Message m = in->pop();
            if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
            b->bytes = n;
            aby3::gpu::Context* owner = m.shared ? m.shared->ctx() : (m.dev ? m.dev.ctx() : nullptr);
            if (owner && owner->device() == ctx->device()) {
                waitReady(m);
                if (m.shared) { b->shared = std::move(m.shared); b->ptr = (const u8*)b->shared->ptr() + m.offset; }
                else { b->own = std::move(m.dev); b->ownedByPeer = true; b->ptr = b->own.ptr(); }
                return;
            }
            // another GPU (NVLink peer copy) or a host payload: land it in a buffer of our own
            b->own.reset(ctx, std::max<size_t>(n, 16));
            b->ptr = b->own.ptr();
            if (on) {       // the buffer comes from the party's pool (own-stream order): `on` starts behind its last use
                void* e = ctx->recordEvent();
                aby3::gpu::check(aby3cu_event_wait(on->h(), e));
                ctx->recycleEvent(e);
            }
            deliver(m, b->own.ptr(), n, on);
        };
    }
    void recvDevice(void* d, size_t n) {
        requireCtx();
        Message m = in->pop();
        if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
        deliver(m, d, n);
    }
    // `on`: the stream (context) of this party that performs the copy; defaults to the party's own
    void deliver(Message& m, void* d, size_t n, aby3::gpu::Context* on = nullptr) {
        aby3::gpu::Context* c = on ? on : ctx;
        if (m.dev || m.shared) {
            aby3::gpu::Context* owner = m.dev ? m.dev.ctx() : m.shared->ctx();
            const void* src = m.dev ? m.dev.ptr() : (const void*)((const u8*)m.shared->ptr() + m.offset);
            const bool same = !m.ready;
            if (m.ready) {
                aby3::gpu::check(aby3cu_event_wait(c->h(), m.ready));
                aby3::gpu::EventPool::put(m.readyDevice, m.ready);
                m.ready = nullptr;
            }
            if (n) aby3::gpu::check(aby3cu_d2d(c->h(), d, c->device(), src, owner->device(), n));
            // the payload returns to the sender's pool once OUR copy has run
            if (m.dev) m.dev.free(same ? nullptr : c->recordEvent(), c->device());
            else {
                if (!same) m.shared->addReader(c->device(), c->recordEvent());
                m.shared.reset();
            }
        } else if (n) {
            aby3::gpu::check(aby3cu_h2d(c->h(), d, m.host.data(), n));
            c->sync();   // m.host dies with this scope
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
    struct Op { bool send; int peer; void* ptr; size_t bytes; std::shared_ptr<aby3::gpu::SharedBuffer> keep; };
    std::vector<Op> pending;
    std::vector<aby3::gpu::Buffer> staging;       // send copies, released after the flush
    void flush() { flushOn(ctx); }
    // issue everything queued as ONE group on `on`'s stream (the party's own, or its communication stream)
    void flushOn(aby3::gpu::Context* on) {
        if (pending.empty()) return;
        auto& api = NcclApi::get();
        void* stream = aby3cu_ctx_stream(on->h());
        api.check(api.GroupStart(), "GroupStart");
        for (auto& op : pending) {
            if (op.send) api.check(api.Send(op.ptr, op.bytes, /*ncclUint8*/ 1, op.peer, comm, stream), "Send");
            else api.check(api.Recv(op.ptr, op.bytes, /*ncclUint8*/ 1, op.peer, comm, stream), "Recv");
        }
        api.check(api.GroupEnd(), "GroupEnd");
        // buffers sent without a staging copy stay alive until the send has run on `on`
        for (auto& op : pending)
            if (op.keep) op.keep->addReader(on->device(), on->recordEvent());
        pending.clear();
        if (on == ctx) staging.clear();            // stream-ordered reuse: the sends above run first (staging is drawn on ctx's stream)
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
        ep->pending.push_back({true, peer, b.ptr(), n, nullptr});
        ep->staging.push_back(std::move(b));
    }
    void sendDeviceSharedOn(aby3::gpu::Context*, const std::shared_ptr<aby3::gpu::SharedBuffer>& b, size_t off, size_t n) override {
        if (!n) return;
        ep->pending.push_back({true, peer, (u8*)b->ptr() + off, n, b});
    }
    std::function<void()> postRecvDeviceOn(aby3::gpu::Context* on, void* d, size_t n) override {
        if (n) ep->pending.push_back({false, peer, d, n, nullptr});
        auto e = ep;
        return [e, on] { e->flushOn(on); };
    }
    void flushOn(aby3::gpu::Context* on) override { ep->flushOn(on); }
    std::function<void()> postRecvDevice(void* d, size_t n) override {
        if (n) ep->pending.push_back({false, peer, d, n, nullptr});
        auto e = ep;
        return [e] { e->flush(); };
    }
    void sendHost(const u8* p, size_t n) override {
        if (!n) return;
        aby3::gpu::Buffer b(ep->ctx, n);
        aby3::gpu::check(aby3cu_h2d(ep->ctx->h(), b.ptr(), p, n));
        ep->ctx->sync();                           // p may be a temporary
        ep->pending.push_back({true, peer, b.ptr(), n, nullptr});
        ep->staging.push_back(std::move(b));
    }
    std::function<void()> postRecvHost(u8* p, size_t n) override {
        auto b = std::make_shared<aby3::gpu::Buffer>(ep->ctx, std::max<size_t>(n, 16));
        if (n) ep->pending.push_back({false, peer, b->ptr(), n, nullptr});
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

using Borrowed = detail::Borrowed;

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
        if (ctxA && ctxB && ctxA->device() == ctxB->device()) { ta->peerStream = ctxB->stream(); tb->peerStream = ctxA->stream(); }
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
    // the party at the other end runs on the same GPU: device pointers of one are valid for the other
    bool colocated() const { return mT && mT->colocated(); }
    void waitForConnection() {}
    void close() {}
    void cancel() {}
    void bindContext(aby3::gpu::Context* c) { require(); mT->bind(c); }
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
    // any other contiguous container with data() / size() -- a share plane, for one (aby3-Basic/BoolBasic.cpp:872)
    template <typename C>
    auto asyncSendCopy(const C& c) -> typename std::enable_if<!std::is_trivially_copyable<C>::value,
                                                              decltype((void)c.data(), (void)c.size(), void())>::type {
        sendBytes(reinterpret_cast<const u8*>(c.data()), c.size() * sizeof(*c.data()));
    }
    // both transports copy eagerly, so the no-copy forms are aliases
    template <typename T>
    void asyncSend(const T* p, u64 n) { asyncSendCopy(p, n); }
    template <typename T>
    void asyncSend(std::vector<T>&& v) { asyncSendCopy(v); v.clear(); }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value>::type
    asyncSend(const T& v) { asyncSendCopy(v); }
    // sends copy eagerly, so the completion future is ready on return (aby3-Basic/Basic.cpp:17)
    template <typename T>
    std::future<void> asyncSendFuture(const T* p, u64 n) {
        asyncSendCopy(p, n);
        std::promise<void> pr;
        pr.set_value();
        return pr.get_future();
    }
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
        *mBytesSent += bytes;
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
    // Zero-copy forms for buffers the protocol never writes again (the opened xy - r of the truncating
    // product, Sh3Evaluator.cpp:676-700).  Local transport on one GPU: no staging copy on either side;
    // other transports copy as above.
    void asyncSendDeviceShared(const std::shared_ptr<aby3::gpu::SharedBuffer>& buf, size_t bytes) {
        require();
        *mBytesSent += bytes;
        mT->sendDeviceShared(buf, bytes);
    }
    std::future<void> asyncRecvDeviceBorrow(size_t bytes, Borrowed* out, aby3::gpu::Context* on = nullptr) {
        require();
        return post(mT->postRecvDeviceBorrow(bytes, out, on));
    }
    // Slices of such a buffer, moved by ANOTHER stream of the party (`on`): see Transport::sendDeviceSharedOn
    void asyncSendDeviceSharedOn(aby3::gpu::Context* on, const std::shared_ptr<aby3::gpu::SharedBuffer>& buf, size_t off, size_t bytes) {
        require();
        *mBytesSent += bytes;
        mT->sendDeviceSharedOn(on, buf, off, bytes);
    }
    std::future<void> asyncRecvDeviceOn(aby3::gpu::Context* on, void* d_dst, size_t bytes) {
        require();
        return post(mT->postRecvDeviceOn(on, d_dst, bytes));
    }
    void flushOn(aby3::gpu::Context* on) { if (mT) mT->flushOn(on); }

    // ------------------------------------------------------------------ stats -
    // copies of a Channel are handles to the same endpoint: they share the counter
    u64 getTotalDataSent() const { return mBytesSent ? mBytesSent->load() : 0; }
    void resetStats() { if (mBytesSent) mBytesSent->store(0); }

private:
    explicit Channel(std::shared_ptr<detail::Transport> t)
        : mT(std::move(t)), mRecvQ(std::make_shared<detail::RecvQueue>()), mBytesSent(std::make_shared<std::atomic<u64>>(0)) {}
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
        *mBytesSent += n;
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
    std::shared_ptr<std::atomic<u64>> mBytesSent;
};

}  // namespace oc
