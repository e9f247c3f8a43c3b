#pragma once
// shim of cryptoTools/Common/MatrixView.h: a non-owning row-major 2D view
#include "cryptoTools/Common/Defines.h"
namespace osuCrypto {
template <typename T>
class MatrixView {
public:
    typedef T value_type;
    typedef T* iterator;
    MatrixView() = default;
    MatrixView(const MatrixView&) = default;
    MatrixView& operator=(const MatrixView&) = default;
    MatrixView(T* d, u64 rows, u64 cols) : mData(d), mRows(rows), mStride(cols) {}
    MatrixView(T* b, T* e, u64 cols) : mData(b), mRows(cols ? u64(e - b) / cols : 0), mStride(cols) {}
    template <typename It, typename = typename std::enable_if<!std::is_pointer<It>::value &&
              std::is_same<typename std::iterator_traits<It>::iterator_category, std::random_access_iterator_tag>::value>::type>
    MatrixView(It b, It e, u64 cols) : mData(b == e ? nullptr : &*b), mRows(cols ? u64(e - b) / cols : 0), mStride(cols) {}
    // from an owning Matrix / another view with compatible element type
    template <typename M, typename = typename std::enable_if<
        std::is_convertible<decltype(std::declval<M&>().data()), T*>::value &&
        std::is_convertible<decltype(std::declval<M&>().rows()), u64>::value>::type>
    MatrixView(M& m) : mData(m.data()), mRows(m.rows()), mStride(m.cols()) {}
    template <typename M, typename = typename std::enable_if<
        std::is_convertible<decltype(std::declval<const M&>().data()), T*>::value &&
        std::is_convertible<decltype(std::declval<const M&>().rows()), u64>::value>::type>
    MatrixView(const M& m) : mData(m.data()), mRows(m.rows()), mStride(m.cols()) {}

    T* data() const { return mData; }
    T* data(u64 row) const { return mData + row * mStride; }
    u64 rows() const { return mRows; }
    u64 cols() const { return mStride; }
    u64 stride() const { return mStride; }
    u64 size() const { return mRows * mStride; }
    std::array<u64, 2> bounds() const { return {{mRows, mStride}}; }
    T* begin() const { return mData; }
    T* end() const { return mData + size(); }
    T& operator()(u64 i) const { return mData[i]; }
    T& operator()(u64 r, u64 c) const { return mData[r * mStride + c]; }
    span<T> operator[](u64 r) const { return span<T>(mData + r * mStride, mStride); }
protected:
    T* mData = nullptr;
    u64 mRows = 0, mStride = 0;
};
}  // namespace osuCrypto
