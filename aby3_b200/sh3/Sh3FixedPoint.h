// Sh3FixedPoint.h -- fixed-point value and share types
// (aby3/sh3/Sh3FixedPoint.h:21-112 fp/f64, :120-255 fpMatrix/f64Matrix,
//  :258-443 sf64/sf64Matrix).  A value is an i64 holding round-toward-zero(v * 2^D).
#pragma once
#include "Sh3Types.h"

namespace aby3 {

enum Decimal { D0 = 0, D8 = 8, D16 = 16, D32 = 32 };

struct monostate {};

template <typename T, Decimal D>
struct fp {
    static const Decimal mDecimal = D;
    T mValue = 0;
    fp() = default;
    fp(const double v) { *this = v; }
    fp(T v, monostate) : mValue(v) {}
    fp operator+(const fp& r) const { return {(T)((u64)mValue + (u64)r.mValue), monostate{}}; }
    fp operator-(const fp& r) const { return {(T)((u64)mValue - (u64)r.mValue), monostate{}}; }
    // 128-bit product, truncating division by 2^D (Sh3FixedPoint.cpp:9-20)
    fp operator*(const fp& r) const {
        __int128 v = (__int128)mValue * (__int128)r.mValue;
        v = v / (__int128)(1ull << D);
        return {(T)v, monostate{}};
    }
    fp operator>>(i64 s) const { return {mValue >> s, monostate{}}; }
    fp operator<<(i64 s) const { return {(T)((u64)mValue << s), monostate{}}; }
    fp& operator+=(const fp& r) { mValue = (T)((u64)mValue + (u64)r.mValue); return *this; }
    fp& operator-=(const fp& r) { mValue = (T)((u64)mValue - (u64)r.mValue); return *this; }
    fp& operator*=(const fp& r) { *this = *this * r; return *this; }
    explicit operator double() const { return mValue / double(T(1) << D); }
    void operator=(const double& v) { mValue = T(v * (T(1) << D)); }     // :94-97
    bool operator==(const fp& v) const { return mValue == v.mValue; }
    bool operator!=(const fp& v) const { return mValue != v.mValue; }
};
template <Decimal D>
using f64 = fp<i64, D>;

template <typename T, Decimal D>
std::ostream& operator<<(std::ostream& o, const fp<T, D>& f) { return o << (double)f; }

template <typename T, Decimal D>
struct fpMatrix {
    using value_type = fp<T, D>;
    static const Decimal mDecimal = D;
    eMatrix<i64> mData;          // raw fixed-point integers

    fpMatrix() = default;
    fpMatrix(u64 r, u64 c) : mData(r, c) {}
    void resize(u64 r, u64 c) { mData.resize(r, c); }
    u64 rows() const { return mData.rows(); }
    u64 cols() const { return mData.cols(); }
    u64 size() const { return mData.size(); }

    fpMatrix operator+(const fpMatrix& r) const { fpMatrix x; x.mData = mData + r.mData; return x; }
    fpMatrix operator-(const fpMatrix& r) const { fpMatrix x; x.mData = mData - r.mData; return x; }
    // matrix product of the integers followed by >> D on every entry (:200-210)
    fpMatrix operator*(const fpMatrix& r) const {
        fpMatrix x;
        x.mData = mData * r.mData;
        for (u64 i = 0; i < x.size(); ++i) x.mData(i) >>= D;
        return x;
    }
    fpMatrix& operator+=(const fpMatrix& r) { mData += r.mData; return *this; }
    fpMatrix& operator-=(const fpMatrix& r) { mData -= r.mData; return *this; }
    fpMatrix& operator*=(const fpMatrix& r) { *this = *this * r; return *this; }

    // element access through a reference wrapper that reads/writes the raw integer
    struct ref {
        i64& v;
        operator value_type() const { return value_type(v, monostate{}); }
        ref& operator=(const value_type& x) { v = x.mValue; return *this; }
        ref& operator=(double d) { value_type t(d); v = t.mValue; return *this; }
        explicit operator double() const { return (double)value_type(v, monostate{}); }
    };
    ref operator()(u64 x, u64 y) { return ref{mData(x, y)}; }
    ref operator()(u64 xy) { return ref{mData(xy)}; }
    value_type operator()(u64 x, u64 y) const { return value_type(mData(x, y), monostate{}); }
    value_type operator()(u64 xy) const { return value_type(mData(xy), monostate{}); }

    eMatrix<i64>& i64Cast() { return mData; }
    const eMatrix<i64>& i64Cast() const { return mData; }
    bool operator==(const fpMatrix& o) const { return mData == o.mData; }
    bool operator!=(const fpMatrix& o) const { return !(*this == o); }
};
template <Decimal D>
using f64Matrix = fpMatrix<i64, D>;

template <Decimal D>
struct sf64 {
    static const Decimal mDecimal = D;
    using value_type = si64::value_type;
    si64 mShare;
    sf64() = default;
    sf64(const std::array<value_type, 2>& d) : mShare(d) {}
    sf64(const Ref<sf64<D>>& s) { mShare.mData[0] = *s.mData[0]; mShare.mData[1] = *s.mData[1]; }
    sf64 operator+(const sf64& r) const { sf64 x; x.mShare = mShare + r.mShare; return x; }
    sf64 operator-(const sf64& r) const { sf64 x; x.mShare = mShare - r.mShare; return x; }
    value_type& operator[](u64 i) { return mShare[i]; }
    const value_type& operator[](u64 i) const { return mShare[i]; }
    si64& i64Cast() { return mShare; }
    const si64& i64Cast() const { return mShare; }
};

template <Decimal D>
struct sf64Matrix : private si64Matrix {
    static const Decimal mDecimal = D;
    struct ConstRow { const sf64Matrix& mMtx; const u64 mIdx; };
    struct Row {
        sf64Matrix& mMtx; const u64 mIdx;
        const Row& operator=(const Row& r) { mMtx.i64Cast().row(mIdx) = r.mMtx.i64Cast().row(r.mIdx); return r; }
        const ConstRow& operator=(const ConstRow& r) {
            mMtx.i64Cast().row(mIdx) = r.mMtx.i64Cast().row(r.mIdx);
            return r;
        }
    };
    struct ConstCol { const sf64Matrix& mMtx; const u64 mIdx; };
    struct Col {
        sf64Matrix& mMtx; const u64 mIdx;
        const Col& operator=(const Col& c) { mMtx.i64Cast().col(mIdx) = c.mMtx.i64Cast().col(c.mIdx); return c; }
        const ConstCol& operator=(const ConstCol& c) {
            mMtx.i64Cast().col(mIdx) = c.mMtx.i64Cast().col(c.mIdx);
            return c;
        }
    };

    sf64Matrix() = default;
    sf64Matrix(u64 x, u64 y) { resize(x, y); }
    void resize(u64 x, u64 y) { si64Matrix::resize(x, y); }
    u64 rows() const { return mShares[0].rows(); }
    u64 cols() const { return mShares[0].cols(); }
    u64 size() const { return mShares[0].size(); }
    Ref<sf64<D>> operator()(u64 x, u64 y) { return Ref<sf64<D>>(mShares[0](x, y), mShares[1](x, y)); }
    Ref<sf64<D>> operator()(u64 xy) { return Ref<sf64<D>>(mShares[0](xy), mShares[1](xy)); }
    // both share planes as one kernel launch (eMatrix::binary2)
    static sf64Matrix& op2(const sf64Matrix& A, const sf64Matrix& B, sf64Matrix& out, int op) {
        eMatrix<i64>::binary2(A.mShares[0], B.mShares[0], out.mShares[0], A.mShares[1], B.mShares[1], out.mShares[1], op);
        return out;
    }
    sf64Matrix& operator+=(const sf64Matrix& B) { return op2(*this, B, *this, ABY3CU_OP_ADD); }
    sf64Matrix& operator-=(const sf64Matrix& B) { return op2(*this, B, *this, ABY3CU_OP_SUB); }
    sf64Matrix operator+(const sf64Matrix& B) const { sf64Matrix r; op2(*this, B, r, ABY3CU_OP_ADD); return r; }
    sf64Matrix operator-(const sf64Matrix& B) const { sf64Matrix r; op2(*this, B, r, ABY3CU_OP_SUB); return r; }
    sf64Matrix transpose() const { sf64Matrix r; eMatrix<i64>::transpose2(mShares[0], mShares[1], r.mShares[0], r.mShares[1]); return r; }
    void transposeInPlace() { eMatrix<i64>::transpose2(mShares[0], mShares[1], mShares[0], mShares[1]); }
    Row row(u64 i) { return Row{*this, i}; }
    Col col(u64 i) { return Col{*this, i}; }
    ConstRow row(u64 i) const { return ConstRow{*this, i}; }
    ConstCol col(u64 i) const { return ConstCol{*this, i}; }
    bool operator==(const sf64Matrix& b) const { return rows() == b.rows() && cols() == b.cols() && mShares == b.mShares; }
    bool operator!=(const sf64Matrix& b) const { return !(*this == b); }
    si64Matrix& i64Cast() { return static_cast<si64Matrix&>(*this); }
    const si64Matrix& i64Cast() const { return static_cast<const si64Matrix&>(*this); }
    eMatrix<i64>& operator[](u64 i) { return mShares[i]; }
    const eMatrix<i64>& operator[](u64 i) const { return mShares[i]; }
};

}  // namespace aby3
