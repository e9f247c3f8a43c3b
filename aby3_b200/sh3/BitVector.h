// BitVector.h -- oc::BitVector as the sh3 applications use it (cryptoTools/Common/BitVector.h, absent from the reference
// tree): a resizable vector of bits addressed LSB-first inside bytes, operator[] returning an assignable bit proxy.
// Call sites: the choice bits of SharedOT::help / asyncRecv (aby3/OT/SharedOT.h:19-46, aby3-Basic/BuildingBlocks.cpp:357-380).
#pragma once
#include "Defines.h"

namespace oc {

class BitVector {
public:
    class Ref {
    public:
        Ref(u8* byte, u8 shift) : mByte(byte), mShift(shift) {}
        Ref& operator=(u8 bit) { *mByte = (u8)((*mByte & ~(1u << mShift)) | ((bit & 1u) << mShift)); return *this; }
        Ref& operator=(const Ref& o) { return *this = (u8)o; }
        operator u8() const { return (u8)((*mByte >> mShift) & 1u); }
    private:
        u8* mByte;
        u8 mShift;
    };
    BitVector() = default;
    explicit BitVector(u64 nbits) { resize(nbits); }
    void resize(u64 nbits, u8 val = 0) { mData.resize((nbits + 7) / 8, val ? 0xFF : 0); mBits = nbits; }
    void reset(u64 nbits = 0) { mData.assign((nbits + 7) / 8, 0); mBits = nbits; }
    u64 size() const { return mBits; }
    u64 sizeBytes() const { return mData.size(); }
    u8* data() { return mData.data(); }
    const u8* data() const { return mData.data(); }
    Ref operator[](u64 i) { return Ref(mData.data() + (i >> 3), (u8)(i & 7)); }
    u8 operator[](u64 i) const { return (u8)((mData[i >> 3] >> (i & 7)) & 1u); }
    void pushBack(u8 bit) { resize(mBits + 1); (*this)[mBits - 1] = bit; }
    bool operator==(const BitVector& o) const {
        if (mBits != o.mBits) return false;
        for (u64 i = 0; i < mBits; ++i) if ((*this)[i] != o[i]) return false;
        return true;
    }
    bool operator!=(const BitVector& o) const { return !(*this == o); }
private:
    std::vector<u8> mData;
    u64 mBits = 0;
};

}  // namespace oc
