// oracle/shim/shim.cpp -- the few out-of-line definitions of the stand-in headers (TEST INFRASTRUCTURE ONLY)
#include <random>
#include "cryptoTools/Common/BitVector.h"
#include "cryptoTools/Common/Log.h"
#include "cryptoTools/Crypto/PRNG.h"
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
void BitVector::randomize(PRNG& prng) { prng.get(mData.data(), mData.size()); }
}  // namespace osuCrypto
