#pragma once
// compat_net.h -- oc::IOService / oc::Session for UNCHANGED reference applications (aby3-ML/main-linear.cpp:181-210,
// aby3-Basic/BuildingBlocks.cpp:150-179) on top of the facade's oc::Channel.
//
// The reference API has no notion of a device, so the binding is implicit here: the first time a party thread touches
// the network (Session::addChannel) it is given a gpu::Context of its own -- device ABY3_PARTY_DEVICES[k] for the k-th
// thread to arrive (default: every party on device 0) -- which becomes the thread's current context; the IOService owns
// the contexts.  Two Sessions started with the same address and name (one Server, one Client) are the two ends of an
// in-process link; the k-th addChannel() on one side is connected to the k-th on the other, as over TCP.
#include <map>

#include "aby3_b200/sh3/Channel.h"

namespace oc {

enum class SessionMode { Client, Server };

class IOService {
public:
    explicit IOService(u64 = 0) {}
    IOService(const IOService&) = delete;
    ~IOService() { stop(); }
    void stop() {}
    void showErrorMessages(bool) {}
    bool mPrint = false;

    // the calling thread's device context (created on first use, owned by this IOService)
    aby3::gpu::Context* threadContext() {
        if (aby3::gpu::Context* c = aby3::gpu::currentSlot()) return c;
        std::lock_guard<std::mutex> g(mMtx);
        const int k = (int)mContexts.size();
        mContexts.emplace_back(new aby3::gpu::Context(deviceFor(k)));
        aby3::gpu::setCurrent(mContexts.back().get());
        return mContexts.back().get();
    }

private:
    static int deviceFor(int k) {
        const char* e = std::getenv("ABY3_PARTY_DEVICES");          // e.g. "0,1,2": party threads in arrival order
        if (!e || !*e) return 0;
        std::vector<int> d;
        std::stringstream ss(e);
        std::string t;
        while (std::getline(ss, t, ',')) if (!t.empty()) d.push_back(std::atoi(t.c_str()));
        return d.empty() ? 0 : d[k % d.size()];
    }
    std::mutex mMtx;
    std::vector<std::unique_ptr<aby3::gpu::Context>> mContexts;
};

namespace detail {
struct SessionLink {
    std::mutex mtx;
    std::vector<std::pair<Channel, Channel>> pairs;       // first = server end, second = client end
    u64 used[2] = {0, 0};
    Channel take(int side, aby3::gpu::Context* ctx) {
        Channel c;
        {
            std::lock_guard<std::mutex> g(mtx);
            const u64 k = used[side]++;
            while (pairs.size() <= k) pairs.push_back(Channel::makePair(nullptr, nullptr));
            c = side == 0 ? pairs[k].first : pairs[k].second;
        }
        c.bindContext(ctx);
        return c;
    }
};
// both ends find each other by "address|name"; the entry is dropped once the second end has taken it
inline std::shared_ptr<SessionLink> sessionRendezvous(const std::string& key) {
    static std::mutex m;
    static std::map<std::string, std::shared_ptr<SessionLink>> waiting;
    std::lock_guard<std::mutex> g(m);
    auto it = waiting.find(key);
    if (it != waiting.end()) { auto l = it->second; waiting.erase(it); return l; }
    auto l = std::make_shared<SessionLink>();
    waiting[key] = l;
    return l;
}
}  // namespace detail

class Session {
public:
    Session() = default;
    Session(IOService& ios, const std::string& addr, SessionMode m, const std::string& name = "") { start(ios, addr, m, name); }
    Session(IOService& ios, const std::string& ip, u32 port, SessionMode m, const std::string& name = "") { start(ios, ip + ":" + std::to_string(port), m, name); }
    void start(IOService& ios, const std::string& addr, SessionMode m, const std::string& name = "") {
        mIos = &ios;
        mLink = detail::sessionRendezvous(addr + "|" + name);
        mSide = m == SessionMode::Server ? 0 : 1;
    }
    void start(IOService& ios, const std::string& ip, u32 port, SessionMode m, const std::string& name = "") { start(ios, ip + ":" + std::to_string(port), m, name); }
    Channel addChannel(const std::string& = "", const std::string& = "") {
        if (!mLink) throw std::runtime_error("Session: not started " LOCATION);
        return mLink->take(mSide, mIos->threadContext());
    }
    void stop() { mLink.reset(); }
private:
    IOService* mIos = nullptr;
    std::shared_ptr<detail::SessionLink> mLink;
    int mSide = 0;
};

}  // namespace oc
