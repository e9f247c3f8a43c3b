#pragma once
// shim of cryptoTools/Crypto/AES.h: AES-128 (AES-NI), ECB and counter-mode helpers.
// ecbEncCounterMode(base, n, out): out[i] = AES_k(toBlock(base + i))   [assumption, see ../../README.md]
#include <wmmintrin.h>
#include "cryptoTools/Common/Defines.h"
namespace osuCrypto {
class AES {
public:
    AES() = default;
    explicit AES(const block& key) { setKey(key); }
    void setKey(const block& key) {
        mRoundKey[0] = key;
        expand<0x01>(0); expand<0x02>(1); expand<0x04>(2); expand<0x08>(3); expand<0x10>(4);
        expand<0x20>(5); expand<0x40>(6); expand<0x80>(7); expand<0x1B>(8); expand<0x36>(9);
    }
    void ecbEncBlock(const block& in, block& out) const { out = enc(in); }
    block ecbEncBlock(const block& in) const { return enc(in); }
    void ecbEncBlocks(const block* in, u64 n, block* out) const { for (u64 i = 0; i < n; ++i) out[i] = enc(in[i]); }
    void ecbEncTwoBlocks(const block* in, block* out) const { out[0] = enc(in[0]); out[1] = enc(in[1]); }
    void ecbEncCounterMode(u64 base, u64 n, block* out) const { for (u64 i = 0; i < n; ++i) out[i] = enc(toBlock(base + i)); }
    void ecbEncCounterMode(u64 base, block& out) const { out = enc(toBlock(base)); }
    void ecbEncCounterMode(block base, u64 n, block* out) const {
        u64 w[2]; std::memcpy(w, &base, 16);
        for (u64 i = 0; i < n; ++i) out[i] = enc(toBlock(w[1], w[0] + i));
    }
    const block& getKey() const { return mRoundKey[0]; }
    std::array<block, 11> mRoundKey;
private:
    template <int RC>
    void expand(int i) {
        __m128i k = mRoundKey[i], t = _mm_aeskeygenassist_si128(k, RC);
        t = _mm_shuffle_epi32(t, 0xFF);
        k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
        k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
        k = _mm_xor_si128(k, _mm_slli_si128(k, 4));
        mRoundKey[i + 1] = _mm_xor_si128(k, t);
    }
    block enc(const block& in) const {
        __m128i s = _mm_xor_si128(in, mRoundKey[0]);
        for (int r = 1; r < 10; ++r) s = _mm_aesenc_si128(s, mRoundKey[r]);
        return _mm_aesenclast_si128(s, mRoundKey[10]);
    }
};
extern const AES mAesFixedKey;
}  // namespace osuCrypto
