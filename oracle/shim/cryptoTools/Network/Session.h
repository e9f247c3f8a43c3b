#pragma once
// shim of cryptoTools/Network/Session.h: two Sessions created with the same address and name (one Server, one Client)
// are the two ends of an in-process link; the k-th addChannel() on one side is connected to the k-th on the other
// (aby3-ML/aby3ML.cpp:6-9 opens four channels per neighbour that way; aby3-Basic/BuildingBlocks.cpp:150-179).
#include <map>
#include "cryptoTools/Network/Channel.h"
namespace osuCrypto {
enum class SessionMode { Client, Server };
class IOService { public: explicit IOService(u64 = 0) {} void stop() {} void showErrorMessages(bool) {} bool mPrint = false; };
namespace shim {
struct Link {
    std::mutex mtx;
    std::vector<std::pair<Channel, Channel>> pairs;       // first = server end, second = client end
    u64 used[2] = {0, 0};
    Channel take(int side) {
        std::lock_guard<std::mutex> g(mtx);
        const u64 k = used[side]++;
        while (pairs.size() <= k) pairs.push_back(Channel::makePair());
        return side == 0 ? pairs[k].first : pairs[k].second;
    }
};
std::shared_ptr<Link> rendezvous(const std::string& key);          // shim.cpp
}  // namespace shim
class Session {
public:
    Session() = default;
    Session(IOService& ios, const std::string& addr, SessionMode m, const std::string& name = "") { start(ios, addr, m, name); }
    Session(IOService& ios, const std::string& ip, u32 port, SessionMode m, const std::string& name = "") { start(ios, ip + ":" + std::to_string(port), m, name); }
    void start(IOService&, const std::string& addr, SessionMode m, const std::string& name = "") {
        mLink = shim::rendezvous(addr + "|" + name);
        mSide = m == SessionMode::Server ? 0 : 1;
    }
    Channel addChannel(const std::string& = "", const std::string& = "") {
        if (!mLink) throw std::runtime_error("Session shim: not started " LOCATION);
        return mLink->take(mSide);
    }
    void stop() { mLink.reset(); }
private:
    std::shared_ptr<shim::Link> mLink;
    int mSide = 0;
};
}  // namespace osuCrypto
