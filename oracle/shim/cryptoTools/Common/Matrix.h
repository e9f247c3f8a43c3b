#pragma once
// shim of cryptoTools/Common/Matrix.h: an owning row-major matrix.  NEW storage is always zero-filled, also
// for AllocType::Uninitialized: the reference relies on fresh allocations reading as zero (e.g. the "x1 = 0"
// operands of Sh3Converter::toBinaryMatrix, Sh3Converter.cpp:82-85, are resized but never written), which holds
// for the untouched pages a real run gets from the OS and must hold here for recycled heap blocks too.
#include "cryptoTools/Common/MatrixView.h"
namespace osuCrypto {
template <typename T>
class Matrix : public MatrixView<T> {
public:
    Matrix() = default;
    Matrix(u64 rows, u64 cols, AllocType t = AllocType::Zeroed) { resize(rows, cols, t); }
    Matrix(const Matrix& o) : MatrixView<T>() { *this = o; }
    Matrix(Matrix&& o) noexcept : MatrixView<T>() { *this = std::move(o); }
    Matrix(const MatrixView<T>& o) { resize(o.rows(), o.cols(), AllocType::Uninitialized); if (o.size()) std::memcpy(this->mData, o.data(), o.size() * sizeof(T)); }
    ~Matrix() { release(); }
    Matrix& operator=(const Matrix& o) {
        if (this != &o) {
            resize(o.rows(), o.cols(), AllocType::Uninitialized);
            if (o.size()) std::memcpy(this->mData, o.mData, o.size() * sizeof(T));
        }
        return *this;
    }
    Matrix& operator=(Matrix&& o) noexcept {
        if (this != &o) {
            release();
            this->mData = o.mData; this->mRows = o.mRows; this->mStride = o.mStride; mCapacity = o.mCapacity;
            o.mData = nullptr; o.mRows = o.mStride = 0; o.mCapacity = 0;
        }
        return *this;
    }
    // keeps the leading min(old, new) ELEMENTS (linear order), like the original
    void resize(u64 rows, u64 cols, AllocType t = AllocType::Zeroed) {
        const u64 n = rows * cols, old = this->size();
        if (n > mCapacity) {
            T* p = n ? static_cast<T*>(::operator new(n * sizeof(T), std::align_val_t(64))) : nullptr;
            if (p) std::memset((void*)p, 0, n * sizeof(T));
            const u64 keep = std::min(old, n);
            if (keep) std::memcpy(p, this->mData, keep * sizeof(T));
            release();
            this->mData = p; mCapacity = n;
        }
        if (t == AllocType::Zeroed && n > old) std::memset((void*)(this->mData + old), 0, (n - old) * sizeof(T));
        this->mRows = rows; this->mStride = cols;
    }
    void setZero() { if (this->size()) std::memset((void*)this->mData, 0, this->size() * sizeof(T)); }
    bool operator==(const Matrix& o) const {
        return this->rows() == o.rows() && this->cols() == o.cols() &&
               (this->size() == 0 || std::memcmp(this->mData, o.mData, this->size() * sizeof(T)) == 0);
    }
    bool operator!=(const Matrix& o) const { return !(*this == o); }
private:
    void release() {
        if (this->mData) ::operator delete((void*)this->mData, std::align_val_t(64));
        this->mData = nullptr; mCapacity = 0;
    }
    u64 mCapacity = 0;
};
}  // namespace osuCrypto
