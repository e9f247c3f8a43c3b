#pragma once
// shim of cryptoTools/Crypto/PRNG.h.  ASSUMPTION (see ../../README.md): the output is ONE contiguous
// AES-128-CTR keystream AES_seed(0) || AES_seed(1) || ..., consumed byte-sequentially by get(); the
// 256-block buffer only batches the generation.  get<bool>() consumes one byte (its low bit).
#include "cryptoTools/Common/Defines.h"
#include "cryptoTools/Crypto/AES.h"
namespace osuCrypto {
class PRNG {
public:
    PRNG() = default;
    explicit PRNG(const block& seed, u64 bufferSize = 256) { SetSeed(seed, bufferSize); }
    PRNG(const PRNG&) = delete;
    PRNG(PRNG&& o) = default;
    PRNG& operator=(PRNG&&) = default;
    void SetSeed(const block& seed, u64 bufferSize = 256) {
        mSeed = seed;
        mAes.setKey(seed);
        mBlockIdx = 0;
        mBuffer.resize(bufferSize);
        mBufferByteCapacity = bufferSize * sizeof(block);
        refillBuffer();
    }
    const block getSeed() const { if (mBuffer.empty()) throw std::runtime_error("PRNG has not been keyed " LOCATION); return mSeed; }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value && !std::is_same<T, bool>::value, T>::type get() {
        T ret;
        implGet((u8*)&ret, sizeof(T));
        return ret;
    }
    template <typename T>
    typename std::enable_if<std::is_same<T, bool>::value, T>::type get() { u8 b; implGet(&b, 1); return (b & 1) != 0; }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value, void>::type get(T* dest, u64 length) { implGet((u8*)dest, length * sizeof(T)); }
    template <typename T>
    typename std::enable_if<std::is_pod<T>::value, void>::type get(span<T> dest) { implGet((u8*)dest.data(), dest.size() * sizeof(T)); }
    u8 getBit() { return get<bool>(); }
    // `block b = prng.get();` -- the target type picks the draw (Sh3BinaryEvaluator.h:99-100, Sh3Converter.h:32-38)
    struct Any {
        PRNG& mPrng;
        template <typename T, typename = typename std::enable_if<std::is_pod<T>::value>::type>
        operator T() { return mPrng.get<T>(); }
    };
    Any get() { return Any{*this}; }
    // std UniformRandomBitGenerator surface (std::shuffle etc.)
    typedef u64 result_type;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return ~0ull; }
    result_type operator()() { return get<result_type>(); }
    result_type operator()(u64 mod) { return get<u64>() % mod; }

    block mSeed;
    std::vector<block> mBuffer;
    AES mAes;
    u64 mBytesIdx = 0, mBlockIdx = 0, mBufferByteCapacity = 0;
private:
    void refillBuffer() {
        if (mBuffer.empty()) throw std::runtime_error("PRNG has not been keyed " LOCATION);
        mAes.ecbEncCounterMode(mBlockIdx, mBuffer.size(), mBuffer.data());
        mBlockIdx += mBuffer.size();
        mBytesIdx = 0;
    }
    void implGet(u8* dest, u64 n) {
        while (n) {
            const u64 step = std::min(n, mBufferByteCapacity - mBytesIdx);
            std::memcpy(dest, (u8*)mBuffer.data() + mBytesIdx, step);
            dest += step; n -= step; mBytesIdx += step;
            if (mBytesIdx == mBufferByteCapacity) refillBuffer();
        }
    }
};
}  // namespace osuCrypto
