#pragma once
// shim of cryptoTools/Common/BitVector.h + BitIterator.h (LSB-first bit addressing inside bytes)
#include "cryptoTools/Common/Defines.h"
#include "cryptoTools/Crypto/PRNG.h"
namespace osuCrypto {
class BitReference {
public:
    BitReference(u8* byte, u8 shift) : mByte(byte), mMask(u8(1u << shift)), mShift(shift) {}
    BitReference(const BitReference&) = default;
    void operator=(const BitReference& o) { *this = (u8)o; }
    void operator=(u8 n) { if (n) *mByte |= mMask; else *mByte &= u8(~mMask); }
    operator u8() const { return u8((*mByte & mMask) >> mShift); }
private:
    u8* mByte; u8 mMask, mShift;
};
class BitIterator {
public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef u8 value_type; typedef i64 difference_type; typedef void pointer; typedef BitReference reference;
    BitIterator() = default;
    BitIterator(u8* byte, u64 shift = 0) : mByte(byte + shift / 8), mShift(u8(shift & 7)) {}
    BitReference operator*() { return BitReference(mByte, mShift); }
    BitIterator& operator++() { if (++mShift == 8) { mShift = 0; ++mByte; } return *this; }
    BitIterator operator++(int) { BitIterator r = *this; ++*this; return r; }
    BitIterator operator+(i64 v) const { i64 p = i64(mShift) + v; BitIterator r; r.mByte = mByte + (p >> 3); r.mShift = u8(p & 7); return r; }
    BitIterator& operator+=(i64 v) { *this = *this + v; return *this; }
    bool operator==(const BitIterator& o) const { return mByte == o.mByte && mShift == o.mShift; }
    bool operator!=(const BitIterator& o) const { return !(*this == o); }
    u8* mByte = nullptr; u8 mShift = 0;
};
class BitVector {
public:
    BitVector() = default;
    explicit BitVector(u64 n) { resize(n); }
    BitVector(u8* data, u64 nbits) { resize(nbits); std::memcpy(mData.data(), data, sizeBytes()); }
    void resize(u64 n, u8 val = 0) { mData.resize((n + 7) / 8, val ? 0xFF : 0); mNumBits = n; }
    void reset(u64 n = 0) { mData.assign((n + 7) / 8, 0); mNumBits = n; }
    u64 size() const { return mNumBits; }
    u64 sizeBytes() const { return (mNumBits + 7) / 8; }
    u8* data() { return mData.data(); }
    const u8* data() const { return mData.data(); }
    BitReference operator[](u64 i) { return BitReference(mData.data() + (i >> 3), u8(i & 7)); }
    u8 operator[](u64 i) const { return u8((mData[i >> 3] >> (i & 7)) & 1); }
    BitIterator begin() { return BitIterator(mData.data(), 0); }
    BitIterator end() { return BitIterator(mData.data(), mNumBits); }
    void pushBack(u8 bit) { resize(mNumBits + 1); (*this)[mNumBits - 1] = bit; }
    void append(u8* data, u64 length, u64 offset = 0) {
        for (u64 i = 0; i < length; ++i) pushBack(u8((data[(offset + i) >> 3] >> ((offset + i) & 7)) & 1));
    }
    void randomize(PRNG& prng);
    bool operator==(const BitVector& o) const {
        if (mNumBits != o.mNumBits) return false;
        for (u64 i = 0; i < mNumBits; ++i) if ((*this)[i] != o[i]) return false;
        return true;
    }
    bool operator!=(const BitVector& o) const { return !(*this == o); }
    template <typename T> span<T> getSpan() { return span<T>((T*)mData.data(), sizeBytes() / sizeof(T)); }
    template <typename T> span<T> getArrayView() { return getSpan<T>(); }
private:
    std::vector<u8> mData;
    u64 mNumBits = 0;
};
inline std::ostream& operator<<(std::ostream& o, const BitVector& v) { for (u64 i = 0; i < v.size(); ++i) o << int(v[i]); return o; }
}  // namespace osuCrypto
