// Channel.h -- the point-to-point channel the sh3 protocols talk through.
// The reference uses cryptoTools' oc::Channel over Boost.Asio TCP even when all
// three parties live in one process (aby3_tests/Sh3EvaluatorTests.cpp:23-36).
// Here the parties are host threads that share an NVSwitch domain, so a message
// is either a small host byte string or a device buffer handed over together
// with the CUDA event that marks it complete: the receiver's stream waits on the
// event and copies (same GPU: D2D; different GPUs: NVLink peer copy).  No host
// synchronisation happens on the device path.
//
// Method subset = what aby3/sh3, aby3-ML and aby3-Basic call:
// asyncSendCopy / asyncSend / send / recv / asyncRecv (future) / getTotalDataSent
// / resetStats, plus the device variants used by the facade itself.
#pragma once
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

// one direction of a channel
struct Pipe {
    std::mutex mtx;
    std::condition_variable cv;
    std::deque<Message> q;
    u64 bytesSent = 0;
    void push(Message&& m) {
        {
            std::lock_guard<std::mutex> g(mtx);
            bytesSent += m.bytes;
            q.push_back(std::move(m));
        }
        cv.notify_all();
    }
    Message pop() {
        std::unique_lock<std::mutex> l(mtx);
        cv.wait(l, [&] { return !q.empty(); });
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

}  // namespace detail

class Channel {
public:
    Channel() = default;

    // two connected endpoints; ctxA / ctxB are the device contexts of the parties
    // holding each end (may be null for host-only use, e.g. scheduler tests)
    static std::pair<Channel, Channel> makePair(aby3::gpu::Context* ctxA = nullptr, aby3::gpu::Context* ctxB = nullptr) {
        auto ab = std::make_shared<detail::Pipe>();
        auto ba = std::make_shared<detail::Pipe>();
        Channel a, b;
        a.mOut = ab; a.mIn = ba; a.mCtx = ctxA; a.mRecvQ = std::make_shared<detail::RecvQueue>();
        b.mOut = ba; b.mIn = ab; b.mCtx = ctxB; b.mRecvQ = std::make_shared<detail::RecvQueue>();
        return {a, b};
    }

    bool isConnected() const { return (bool)mOut; }
    aby3::gpu::Context* context() const { return mCtx; }

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
    // the in-process channel copies eagerly, so the no-copy forms are aliases
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
    void recv(T* p, u64 n) { mRecvQ->drain(); recvBytes(reinterpret_cast<u8*>(p), n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value>::type
    recv(T& v) { mRecvQ->drain(); recvBytes(reinterpret_cast<u8*>(&v), sizeof(T)); }
    template <typename T>
    void recv(std::vector<T>& v) { mRecvQ->drain(); recvBytes(reinterpret_cast<u8*>(v.data()), v.size() * sizeof(T)); }

    template <typename T>
    std::future<void> asyncRecv(T* p, u64 n) {
        return post([this, p, n] { recvBytes(reinterpret_cast<u8*>(p), n * sizeof(T)); });
    }
    template <typename T>
    typename std::enable_if<std::is_trivially_copyable<T>::value && !std::is_pointer<T>::value, std::future<void>>::type
    asyncRecv(T& v) {
        T* p = &v;
        return post([this, p] { recvBytes(reinterpret_cast<u8*>(p), sizeof(T)); });
    }
    template <typename T>
    std::future<void> asyncRecv(std::vector<T>& v) {
        std::vector<T>* p = &v;
        return post([this, p] { recvBytes(reinterpret_cast<u8*>(p->data()), p->size() * sizeof(T)); });
    }

    // ---------------------------------------------------------- device path ---
    // Send `bytes` of device memory that is (or will be, in stream order) valid on
    // this party's stream.  Copy semantics: the caller may overwrite d_src right
    // after the call (asyncSendCopy at Sh3Evaluator.cpp:109, :681-684).
    void asyncSendDevice(const void* d_src, size_t bytes) {
        requireCtx();
        detail::Message m;
        m.bytes = bytes;
        m.dev.reset(mCtx, std::max<size_t>(bytes, 16));
        if (bytes)
            aby3::gpu::check(aby3cu_d2d(mCtx->h(), m.dev.ptr(), mCtx->device(), d_src, mCtx->device(), bytes));
        aby3::gpu::check(aby3cu_event_create(mCtx->h(), &m.ready));
        aby3::gpu::check(aby3cu_event_record(mCtx->h(), m.ready));
        mOut->push(std::move(m));
    }
    // Receive into device memory: blocks the HOST only until the peer has posted
    // the message; the copy itself is enqueued on this party's stream behind the
    // sender's event.
    void recvDevice(void* d_dst, size_t bytes) { mRecvQ->drain(); recvDeviceNow(d_dst, bytes); }
    std::future<void> asyncRecvDevice(void* d_dst, size_t bytes) {
        return post([this, d_dst, bytes] { recvDeviceNow(d_dst, bytes); });
    }

    // ------------------------------------------------------------------ stats -
    u64 getTotalDataSent() const { return mOut ? mOut->bytesSent : 0; }
    void resetStats() { if (mOut) mOut->bytesSent = 0; }

private:
    void requireCtx() const {
        if (!mCtx) throw std::runtime_error("Channel: device transfer on a channel without a device context " LOCATION);
    }
    std::future<void> post(std::function<void()> f) {
        auto q = mRecvQ;
        const u64 ticket = q->post(std::move(f));
        return std::async(std::launch::deferred, [q, ticket] { q->completeUpTo(ticket); });
    }
    void sendBytes(const u8* p, size_t n) {
        if (!mOut) throw std::runtime_error("Channel: not connected " LOCATION);
        detail::Message m;
        m.bytes = n;
        m.host.assign(p, p + n);
        mOut->push(std::move(m));
    }
    void recvBytes(u8* p, size_t n) {
        if (!mIn) throw std::runtime_error("Channel: not connected " LOCATION);
        detail::Message m = mIn->pop();
        if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
        if (m.dev) {
            // device message consumed on the host (e.g. reveal into an i64Matrix)
            requireCtx();
            aby3::gpu::check(aby3cu_event_wait(mCtx->h(), m.ready));
            if (n) aby3::gpu::check(aby3cu_d2h(mCtx->h(), p, m.dev.ptr(), n));
            mCtx->sync();
            aby3cu_event_destroy(m.ready);
            m.dev.free();
        } else if (n) {
            memcpy(p, m.host.data(), n);
        }
    }
    void recvDeviceNow(void* d_dst, size_t n) {
        requireCtx();
        detail::Message m = mIn->pop();
        if (m.bytes != n) throw std::runtime_error("Channel: message size mismatch " LOCATION);
        if (m.dev) {
            aby3::gpu::check(aby3cu_event_wait(mCtx->h(), m.ready));
            if (n)
                aby3::gpu::check(aby3cu_d2d(mCtx->h(), d_dst, mCtx->device(), m.dev.ptr(), m.dev.ctx()->device(), n));
            aby3cu_event_destroy(m.ready);
            // the staging buffer returns to the sender's pool once OUR copy has run
            void* done = nullptr;
            aby3::gpu::check(aby3cu_event_create(mCtx->h(), &done));
            aby3::gpu::check(aby3cu_event_record(mCtx->h(), done));
            m.dev.free(done);
        } else if (n) {
            aby3::gpu::check(aby3cu_h2d(mCtx->h(), d_dst, m.host.data(), n));
            mCtx->sync();   // m.host dies with this scope
        }
    }

    std::shared_ptr<detail::Pipe> mOut, mIn;
    std::shared_ptr<detail::RecvQueue> mRecvQ;
    aby3::gpu::Context* mCtx = nullptr;
};

}  // namespace oc
