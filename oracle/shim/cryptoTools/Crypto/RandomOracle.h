#pragma once
// shim of cryptoTools/Crypto/RandomOracle.h.  Only the reference's DEBUG paths hash state with it
// (Sh3BinaryEvaluator.h:128, .cpp:422); a keyed AES Davies-Meyer chain stands in (NOT a cryptographic
// restatement of Blake2 -- values are only ever compared with values from this same shim).
#include "cryptoTools/Common/Defines.h"
#include "cryptoTools/Crypto/AES.h"
namespace osuCrypto {
class RandomOracle {
public:
    static const u64 MaxHashSize = 20, HashSize = 20;
    explicit RandomOracle(u64 outputLength = 20) : mOut(outputLength) { Reset(); }
    void Reset() { mState = toBlock(0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull); mFill = 0; std::memset(mBuf, 0, 16); mLen = 0; }
    void Reset(u64 outputLength) { mOut = outputLength; Reset(); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value>::type Update(const T* data, u64 n) { absorb((const u8*)data, n * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value>::type Update(const T& v) { absorb((const u8*)&v, sizeof(T)); }
    void Final(u8* out) {
        u64 len = mLen;
        u8 pad[16] = {0x80};
        absorb(pad, 16 - mFill ? 16 - mFill : 16);
        absorb((const u8*)&len, 8);
        u8 pad2[8] = {0};
        absorb(pad2, 8);
        u8 o[32];
        block s2 = AES(mState).ecbEncBlock(toBlock(1));
        std::memcpy(o, &mState, 16); std::memcpy(o + 16, &s2, 16);
        std::memcpy(out, o, std::min<u64>(mOut, 32));
    }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value && sizeof(T) <= 32>::type Final(T& v) {
        u8 o[32]; const u64 keep = mOut; mOut = 32; Final(o); mOut = keep; std::memcpy(&v, o, sizeof(T));
    }
private:
    void absorb(const u8* p, u64 n) {
        mLen += n;
        while (n) {
            const u64 step = std::min<u64>(n, 16 - mFill);
            std::memcpy(mBuf + mFill, p, step);
            mFill += step; p += step; n -= step;
            if (mFill == 16) { block m; std::memcpy(&m, mBuf, 16); mState = AES(m).ecbEncBlock(mState) ^ mState; mFill = 0; }
        }
    }
    block mState; u8 mBuf[16]; u64 mFill = 0, mLen = 0, mOut;
};
}  // namespace osuCrypto
