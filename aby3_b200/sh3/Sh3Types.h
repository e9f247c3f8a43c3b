// Sh3Types.h -- share types of the sh3 API (aby3/sh3/Sh3Types.h:32-34, 128-387).
// Same names, members and meaning as the reference so callers compile unchanged:
// party i holds (x_i, x_{i-1}); mShares[0] / mData[0] is the party's own share,
// [1] the previous party's.  Storage is eMatrix (HBM-backed, host view on demand).
#pragma once
#include "Channel.h"
#include "Crypto.h"
#include "eMatrix.h"

namespace aby3 {

struct CommPkg {
    oc::Channel mPrev, mNext;
};

template <typename ShareType>
struct Ref {
    using ref_value_type = typename ShareType::value_type;
    std::array<ref_value_type*, 2> mData;
    Ref(ref_value_type& a0, ref_value_type& a1) { mData[0] = &a0; mData[1] = &a1; }
    const ShareType& operator=(const ShareType& copy) {
        *mData[0] = copy[0];
        *mData[1] = copy[1];
        return copy;
    }
    ref_value_type& operator[](u64 i) { return *mData[i]; }
    const ref_value_type& operator[](u64 i) const { return *mData[i]; }
};

// a replicated arithmetic share of one 64-bit value
template <typename T>
struct Share {
    using value_type = i64;
    std::array<value_type, 2> mData{};
    Share() = default;
    Share(const std::array<value_type, 2>& d) : mData(d) {}
    Share(const Ref<Share>& s) { mData[0] = *s.mData[0]; mData[1] = *s.mData[1]; }
    Share operator+(const Share& r) const {
        return Share(std::array<value_type, 2>{{(i64)((u64)mData[0] + (u64)r.mData[0]), (i64)((u64)mData[1] + (u64)r.mData[1])}});
    }
    Share operator-(const Share& r) const {
        return Share(std::array<value_type, 2>{{(i64)((u64)mData[0] - (u64)r.mData[0]), (i64)((u64)mData[1] - (u64)r.mData[1])}});
    }
    value_type& operator[](u64 i) { return mData[i]; }
    const value_type& operator[](u64 i) const { return mData[i]; }
};
using si64 = Share<i64>;
using sia64 = i64;

// a replicated binary (xor) share of 64 bits
struct sb64 {
    using value_type = i64;
    std::array<i64, 2> mData{};
    sb64() = default;
    sb64(const std::array<value_type, 2>& d) : mData(d) {}
    i64& operator[](u64 i) { return mData[i]; }
    const i64& operator[](u64 i) const { return mData[i]; }
    sb64 operator^(const sb64& x) const {
        return sb64(std::array<value_type, 2>{{mData[0] ^ x.mData[0], mData[1] ^ x.mData[1]}});
    }
};

template <typename T>
struct sMatrix {
    std::array<eMatrix<T>, 2> mShares;

    struct ConstRow { const sMatrix& mMtx; const u64 mIdx; };
    struct Row {
        sMatrix& mMtx; const u64 mIdx;
        const Row& operator=(const Row& r) { return assign(r.mMtx, r.mIdx); }
        const Row& operator=(const ConstRow& r) { return assign(r.mMtx, r.mIdx); }
    private:
        const Row& assign(const sMatrix& src, u64 srcIdx) {
            for (int s = 0; s < 2; ++s)
                for (u64 j = 0; j < mMtx.cols(); ++j) mMtx.mShares[s](mIdx, j) = src.mShares[s](srcIdx, j);
            return *this;
        }
    };
    struct ConstCol { const sMatrix& mMtx; const u64 mIdx; };
    struct Col {
        sMatrix& mMtx; const u64 mIdx;
        const Col& operator=(const Col& c) { return assign(c.mMtx, c.mIdx); }
        const Col& operator=(const ConstCol& c) { return assign(c.mMtx, c.mIdx); }
    private:
        const Col& assign(const sMatrix& src, u64 srcIdx) {
            for (int s = 0; s < 2; ++s)
                for (u64 i = 0; i < mMtx.rows(); ++i) mMtx.mShares[s](i, mIdx) = src.mShares[s](i, srcIdx);
            return *this;
        }
    };

    sMatrix() = default;
    sMatrix(u64 xSize, u64 ySize) { resize(xSize, ySize); }
    void resize(u64 xSize, u64 ySize) {
        mShares[0].resize(xSize, ySize);
        mShares[1].resize(xSize, ySize);
    }
    u64 rows() const { return mShares[0].rows(); }
    u64 cols() const { return mShares[0].cols(); }
    u64 size() const { return mShares[0].size(); }

    Ref<Share<T>> operator()(u64 x, u64 y) const {
        auto& s = const_cast<sMatrix&>(*this);
        return Ref<Share<T>>(s.mShares[0](x, y), s.mShares[1](x, y));
    }
    Ref<Share<T>> operator()(u64 xy) const {
        auto& s = const_cast<sMatrix&>(*this);
        return Ref<Share<T>>(s.mShares[0](xy), s.mShares[1](xy));
    }
    void operator()(u64 x, u64 y, Share<T> v) { mShares[0](x, y) = v[0]; mShares[1](x, y) = v[1]; }

    // local (communication-free) share arithmetic -- Sh3Types.h:805-820; on device
    // (both planes go out as one kernel launch)
    sMatrix operator+(const sMatrix& B) const {
        sMatrix r;
        eMatrix<T>::binary2(mShares[0], B.mShares[0], r.mShares[0], mShares[1], B.mShares[1], r.mShares[1], ABY3CU_OP_ADD);
        return r;
    }
    sMatrix operator-(const sMatrix& B) const {
        sMatrix r;
        eMatrix<T>::binary2(mShares[0], B.mShares[0], r.mShares[0], mShares[1], B.mShares[1], r.mShares[1], ABY3CU_OP_SUB);
        return r;
    }
    sMatrix transpose() const {
        sMatrix r;
        eMatrix<T>::transpose2(mShares[0], mShares[1], r.mShares[0], r.mShares[1]);
        return r;
    }
    void transposeInPlace() { eMatrix<T>::transpose2(mShares[0], mShares[1], mShares[0], mShares[1]); }

    Row row(u64 i) { return Row{*this, i}; }
    Col col(u64 i) { return Col{*this, i}; }
    ConstRow row(u64 i) const { return ConstRow{*this, i}; }
    ConstCol col(u64 i) const { return ConstCol{*this, i}; }

    bool operator==(const sMatrix& b) const { return rows() == b.rows() && cols() == b.cols() && mShares == b.mShares; }
    bool operator!=(const sMatrix& b) const { return !(*this == b); }
    eMatrix<T>& operator[](u64 i) { return mShares[i]; }
    const eMatrix<T>& operator[](u64 i) const { return mShares[i]; }
};
using si64Matrix = sMatrix<i64>;

// rows x bitCount binary sharing, stored as rows x ceil(bitCount/64) words
// (Sh3Types.h:335-387).  NB the second constructor argument is a BIT count.
struct sbMatrix {
    std::array<eMatrix<i64>, 2> mShares;
    u64 mBitCount = 0;
    sbMatrix() = default;
    sbMatrix(u64 xSize, u64 bitCount) { resize(xSize, bitCount); }
    void resize(u64 xSize, u64 bitCount) {
        mBitCount = bitCount;
        const u64 ySize = (bitCount + 63) / 64;
        mShares[0].resize(xSize, ySize);
        mShares[1].resize(xSize, ySize);
    }
    u64 rows() const { return mShares[0].rows(); }
    u64 i64Size() const { return mShares[0].size(); }
    u64 i64Cols() const { return mShares[0].cols(); }
    u64 bitCount() const { return mBitCount; }
    eMatrix<i64>& operator[](u64 i) { return mShares[i]; }
    const eMatrix<i64>& operator[](u64 i) const { return mShares[i]; }
    void trim() {
        const u64 rem = mBitCount % 64;
        if (!rem || !i64Cols()) return;
        const u64 mask = (1ull << rem) - 1;
        for (auto& s : mShares) {
            if (s.onDevice()) {
                gpu::check(aby3cu_mask_last_word(s.ctx()->h(), s.devMut(), s.rows(), s.cols(), mask));
            } else {
                for (u64 r = 0; r < s.rows(); ++r) s(r, s.cols() - 1) &= (i64)mask;
            }
        }
    }
    bool operator==(const sbMatrix& b) const {
        if (rows() != b.rows() || bitCount() != b.bitCount()) return false;
        sbMatrix x = *this, y = b;
        x.trim(); y.trim();
        return x.mShares == y.mShares;
    }
    bool operator!=(const sbMatrix& b) const { return !(*this == b); }
};

// A packed (bit-sliced) set of binary secrets: row i holds bit i of every share, one bit per
// instance, LSB first (Sh3Types.h:445-652).  Rows are `simdWidth()` 64-bit words long.
template <typename T = i64>
struct sPackedBinBase {
    static_assert(sizeof(T) == 8, "the device layout uses 64-bit words");
    u64 mShareCount = 0;
    std::array<eMatrix<i64>, 2> mShares;
    sPackedBinBase() = default;
    sPackedBinBase(u64 shareCount, u64 bitCount, u64 rowAlignment = 1) { reset(shareCount, bitCount, rowAlignment); }
    // shareCount independent secrets of bitCount bits each; rows start on multiples of rowAlignment words
    void reset(u64 shareCount, u64 bitCount, u64 rowAlignment = 1) {
        const u64 rowSizeT = oc::divCeil(shareCount, rowAlignment * 64) * rowAlignment;     // :545
        mShareCount = shareCount;
        mShares[0].resize(bitCount, rowSizeT);
        mShares[1].resize(bitCount, rowSizeT);
    }
    void reshape(u64 shareCount) {
        if (shareCount > mShares[0].cols() * 64) throw RTE_LOC;
        mShareCount = shareCount;
    }
    u64 size() const { return mShares[0].size(); }
    u64 shareCount() const { return mShareCount; }
    u64 bitCount() const { return mShares[0].rows(); }
    u64 simdWidth() const { return mShares[0].cols(); }
};
using sPackedBin = sPackedBinBase<i64>;

}  // namespace aby3
