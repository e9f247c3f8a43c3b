#pragma once
// shim of cryptoTools/Network/Channel.h: an in-process, ordered, reliable byte-message pipe with the
// subset of the oc::Channel surface the reference calls (asyncSend/asyncSendCopy/send, recv,
// asyncRecv -> std::future<void>, asyncRecv with a completion callback).  Message k of a direction is
// delivered to the k-th receive posted on the other end, whichever side arrives first.
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <mutex>
#include "cryptoTools/Common/Defines.h"
namespace osuCrypto {
namespace shim {
struct RecvReq {
    std::function<u8*(u64 bytes)> place;          // where the payload goes (may resize a container)
    std::promise<void> done;
    std::function<void()> callback;
};
struct Pipe {
    std::mutex mtx;
    std::deque<std::vector<u8>> msgs;
    std::deque<std::unique_ptr<RecvReq>> reqs;
    u64 bytes = 0;
    static void complete(RecvReq& r, std::vector<u8>& m) {
        try {
            u8* d = r.place(m.size());
            if (m.size()) std::memcpy(d, m.data(), m.size());
            r.done.set_value();
        } catch (...) { r.done.set_exception(std::current_exception()); }
        if (r.callback) r.callback();
    }
    void push(std::vector<u8>&& m) {
        std::unique_ptr<RecvReq> r;
        {
            std::lock_guard<std::mutex> g(mtx);
            bytes += m.size();
            if (reqs.empty()) { msgs.push_back(std::move(m)); return; }
            r = std::move(reqs.front());
            reqs.pop_front();
        }
        complete(*r, m);
    }
    std::future<void> post(std::unique_ptr<RecvReq> r) {
        std::future<void> f = r->done.get_future();
        std::vector<u8> m;
        {
            std::lock_guard<std::mutex> g(mtx);
            if (msgs.empty() || !reqs.empty()) { reqs.push_back(std::move(r)); return f; }
            m = std::move(msgs.front());
            msgs.pop_front();
        }
        complete(*r, m);
        return f;
    }
};
}  // namespace shim

class Channel {
public:
    Channel() = default;
    // two connected endpoints
    static std::pair<Channel, Channel> makePair() {
        auto ab = std::make_shared<shim::Pipe>(), ba = std::make_shared<shim::Pipe>();
        Channel a, b;
        a.mOut = ab; a.mIn = ba; b.mOut = ba; b.mIn = ab;
        return {a, b};
    }
    bool isConnected() const { return (bool)mOut; }
    void waitForConnection() {}
    void close() {}
    void cancel() {}
    std::string getName() const { return "shim"; }
    u64 getTotalDataSent() const { return mOut ? mOut->bytes : 0; }
    u64 getTotalDataRecv() const { return mIn ? mIn->bytes : 0; }
    void resetStats() {}

    // ---- sends (all copy: the reference's lifetime rules are a superset) --------------------
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value>::type asyncSend(const T* p, u64 n) { sendBytes(p, n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value>::type asyncSendCopy(const T* p, u64 n) { sendBytes(p, n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value>::type send(const T* p, u64 n) { sendBytes(p, n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value && !is_container<T>::value>::type asyncSendCopy(const T& v) { sendBytes(&v, sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value && !is_container<T>::value>::type asyncSend(const T& v) { sendBytes(&v, sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value && !is_container<T>::value>::type send(const T& v) { sendBytes(&v, sizeof(T)); }
    template <typename C>
    typename std::enable_if<is_container<C>::value>::type asyncSendCopy(const C& c) { sendBytes(c.data(), c.size() * sizeof(*c.data())); }
    template <typename C>
    typename std::enable_if<is_container<C>::value>::type send(const C& c) { sendBytes(c.data(), c.size() * sizeof(*c.data())); }
    template <typename C>
    typename std::enable_if<is_container<typename std::remove_reference<C>::type>::value && !std::is_lvalue_reference<C>::value>::type
    asyncSend(C&& c) { sendBytes(c.data(), c.size() * sizeof(*c.data())); C drop(std::move(c)); (void)drop; }
    template <typename C>
    typename std::enable_if<is_container<C>::value>::type asyncSend(std::unique_ptr<C> c) { sendBytes(c->data(), c->size() * sizeof(*c->data())); }

    // sends copy, so the future is ready on return (aby3-Basic/Basic.cpp:17)
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value, std::future<void>>::type asyncSendFuture(const T* p, u64 n) {
        sendBytes(p, n * sizeof(T));
        std::promise<void> pr; pr.set_value();
        return pr.get_future();
    }

    // ---- receives ---------------------------------------------------------------------------
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value, std::future<void>>::type asyncRecv(T* p, u64 n) {
        return post(fixed((u8*)p, n * sizeof(T)), nullptr);
    }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value, std::future<void>>::type asyncRecv(T* p, u64 n, std::function<void()> cb) {
        return post(fixed((u8*)p, n * sizeof(T)), std::move(cb));
    }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value && !is_container<T>::value, std::future<void>>::type asyncRecv(T& v) {
        return post(fixed((u8*)&v, sizeof(T)), nullptr);
    }
    template <typename C>
    typename std::enable_if<is_resizable_container<C>::value, std::future<void>>::type asyncRecv(C& c) {
        C* cp = &c;
        return post([cp](u64 bytes) -> u8* {
            typedef typename std::remove_reference<decltype(*cp->data())>::type E;
            if (bytes % sizeof(E)) throw std::runtime_error("Channel shim: message size is not a multiple of the element size " LOCATION);
            cp->resize(bytes / sizeof(E));
            return (u8*)cp->data();
        }, nullptr);
    }
    template <typename C>
    typename std::enable_if<is_container<C>::value && !is_resizable_container<C>::value, std::future<void>>::type asyncRecv(C& c) {
        return post(fixed((u8*)c.data(), c.size() * sizeof(*c.data())), nullptr);
    }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value>::type recv(T* p, u64 n) { asyncRecv(p, n).get(); }
    template <typename T>
    void recv(T& v) { asyncRecv(v).get(); }

private:
    static std::function<u8*(u64)> fixed(u8* p, u64 want) {
        return [p, want](u64 bytes) -> u8* {
            if (bytes != want) throw std::runtime_error("Channel shim: received " + std::to_string(bytes) + " bytes, expected " + std::to_string(want) + " " LOCATION);
            return p;
        };
    }
    void sendBytes(const void* p, u64 n) {
        if (!mOut) throw std::runtime_error("Channel shim: not connected " LOCATION);
        std::vector<u8> m((const u8*)p, (const u8*)p + n);
        mOut->push(std::move(m));
    }
    std::future<void> post(std::function<u8*(u64)> place, std::function<void()> cb) {
        if (!mIn) throw std::runtime_error("Channel shim: not connected " LOCATION);
        std::unique_ptr<shim::RecvReq> r(new shim::RecvReq);
        r->place = std::move(place);
        r->callback = std::move(cb);
        return mIn->post(std::move(r));
    }
    std::shared_ptr<shim::Pipe> mOut, mIn;
};
}  // namespace osuCrypto
