// oracle/shim/shim.cpp -- the few out-of-line definitions of the stand-in headers (TEST INFRASTRUCTURE ONLY)
#include <random>
#include "cryptoTools/Common/BitVector.h"
#include "cryptoTools/Common/Log.h"
#include "cryptoTools/Crypto/PRNG.h"
#include "cryptoTools/Network/Session.h"
namespace osuCrypto {
std::mutex gIoStreamMtx;
ostreamLocker lout(std::cout);
LogAdapter gLog;
const AES mAesFixedKey(toBlock(45345336, 197134319));
void setThreadName(const std::string&) {}
block sysRandomSeed() {
    std::random_device rd;
    u64 a = (u64(rd()) << 32) | rd(), b = (u64(rd()) << 32) | rd();
    return toBlock(a, b);
}
namespace shim {
// both ends of a link find each other by "address|name"; the entry is dropped once both ends have taken it
std::shared_ptr<Link> rendezvous(const std::string& key) {
    static std::mutex m;
    static std::map<std::string, std::shared_ptr<Link>> waiting;
    std::lock_guard<std::mutex> g(m);
    auto it = waiting.find(key);
    if (it != waiting.end()) { auto l = it->second; waiting.erase(it); return l; }
    auto l = std::make_shared<Link>();
    waiting[key] = l;
    return l;
}
}  // namespace shim
void BitVector::randomize(PRNG& prng) { prng.get(mData.data(), mData.size()); }
}  // namespace osuCrypto
