#pragma once
// shim of cryptoTools/Network/Session.h (Sh3Runtime.h:4 includes it; nothing on the path uses a Session)
#include "cryptoTools/Network/Channel.h"
namespace osuCrypto {
enum class SessionMode { Client, Server };
class IOService { public: explicit IOService(u64 = 0) {} void stop() {} bool mPrint = false; };
class Session {
public:
    Session() = default;
    Session(IOService&, const std::string&, SessionMode, const std::string& = "") {}
    void start(IOService&, const std::string&, SessionMode, const std::string& = "") {}
    Channel addChannel(const std::string& = "", const std::string& = "") { throw std::runtime_error("Session shim: sockets are not part of the oracle " LOCATION); }
    void stop() {}
};
}  // namespace osuCrypto
