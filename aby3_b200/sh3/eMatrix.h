// eMatrix.h -- the dense row-major matrix the sh3 API is written against.
// The reference aliases Eigen (aby3/sh3/Sh3Types.h:38-39,
// `eMatrix<T> = Eigen::Matrix<T, Dynamic, Dynamic, RowMajor>`); Eigen is not in
// the reference tree, and here the storage of record is HBM.  This class keeps
// the subset of the Eigen surface that aby3's sh3 / ML / Basic code uses
//  and adds lazy host<->device coherence:
//   * protocol code reads/writes the device copy (dev(), devMut(), devOut());
//   * application code that indexes m(i) / m.data() on the host gets a host
//     copy that is downloaded on first touch and re-uploaded on next device use.
// Bulk arithmetic on device-resident matrices runs as aby3cu kernels on the
// calling party's stream; purely host-resident (plaintext) matrices use plain
// loops, exactly like the reference's plaintext side.
#pragma once
#include <type_traits>

#include "Gpu.h"

namespace aby3 {

template <typename T>
class eMatrix {
    static_assert(std::is_trivially_copyable<T>::value, "eMatrix needs POD elements");
    static constexpr bool kDeviceOps = sizeof(T) == 8 && std::is_integral<T>::value;

public:
    using Scalar = T;
    using value_type = T;

    eMatrix() = default;
    eMatrix(u64 r, u64 c) { resize(r, c); }
    eMatrix(const eMatrix& o) { copyFrom(o); }
    eMatrix(eMatrix&& o) noexcept { moveFrom(std::move(o)); }
    eMatrix& operator=(const eMatrix& o) { if (this != &o) copyFrom(o); return *this; }
    eMatrix& operator=(eMatrix&& o) noexcept { if (this != &o) moveFrom(std::move(o)); return *this; }
    ~eMatrix() { settleUpload(); }

    // -------------------------------------------------------------- shape ----
    u64 rows() const { return mRows; }
    u64 cols() const { return mCols; }
    u64 size() const { return mRows * mCols; }
    // Eigen semantics: contents are unspecified after a size change
    void resize(u64 r, u64 c) {
        if (r == mRows && c == mCols) return;
        settleUpload();
        mRows = r; mCols = c;
        // no storage is touched here: a matrix that only ever lives on the device
        // never allocates (or zero-fills) a host copy
        mHost.clear();
        mHostValid = false; mDevValid = false;
        mDev.free();
    }
    template <typename M>
    void resizeLike(const M& m) { resize(m.rows(), m.cols()); }
    void conservativeResize(u64 r, u64 c) {
        eMatrix n(r, c);
        const u64 rr = std::min(r, mRows), cc = std::min(c, mCols);
        const T* src = hostData();
        T* dst = n.data();
        for (u64 i = 0; i < rr; ++i)
            for (u64 j = 0; j < cc; ++j) dst[i * c + j] = src[i * mCols + j];
        *this = std::move(n);
    }

    // ------------------------------------------------------ host access ------
    T* data() { touchHost(true); return mHost.data(); }
    const T* data() const { touchHost(false); return mHost.data(); }
    T& operator()(u64 i) { touchHost(true); return mHost[i]; }
    const T& operator()(u64 i) const { touchHost(false); return mHost[i]; }
    T& operator()(u64 r, u64 c) { touchHost(true); return mHost[r * mCols + c]; }
    const T& operator()(u64 r, u64 c) const { touchHost(false); return mHost[r * mCols + c]; }
    const T* hostData() const { touchHost(false); return mHost.data(); }
    // element iterators over the host view (aby3-Basic/BoolBasic.cpp:804-809 std::copy's share planes)
    T* begin() { return data(); }
    T* end() { return data() + size(); }
    const T* begin() const { return hostData(); }
    const T* end() const { return hostData() + size(); }

    // ---------------------------------------------------- device access ------
    bool onDevice() const { return mDevValid; }
    // read-only device pointer (uploads the host copy if that is the fresh one)
    const T* dev() const { touchDev(); return static_cast<const T*>(mDev.ptr()); }
    // read-write device pointer; the host copy becomes stale
    T* devMut() { touchDev(); mHostValid = false; return static_cast<T*>(mDev.ptr()); }
    // write-only device pointer: no upload, previous contents are dropped
    T* devOut() {
        orderBehindPrefetch();
        ensureDevBuffer();
        mDevValid = true; mHostValid = false;
        return static_cast<T*>(mDev.ptr());
    }
    // adopt a device buffer produced elsewhere (e.g. a received message)
    void adoptDevice(gpu::Buffer&& b) {
        mDev = std::move(b);
        mDevValid = true; mHostValid = false;
    }
    gpu::Context* ctx() const { return mDev.ctx() ? mDev.ctx() : gpu::current(); }
    // an upload started by prefetchDevice() that no stream has been ordered behind yet
    bool prefetchPending() const { return mDevReady != nullptr; }

    // ---------------------------------------------- overlapped transfers -----
    // Upload the host copy on ANOTHER context's stream (a copy stream), so that it overlaps the kernels queued on the
    // party's own stream; the first device use waits for it.  The caller guarantees that no kernel still reads the
    // previous device contents and must not write the host copy until the upload has run.
    void prefetchDevice(gpu::Context* copy) {
        if (!mHostValid || !size()) return;
        ensureDevBuffer();
        gpu::check(aby3cu_h2d(copy->h(), mDev.ptr(), mHost.data(), size() * sizeof(T)));
        dropEvent(mDevReady, mDevReadyDevice);
        mDevReady = copy->recordEvent();
        mDevReadyDevice = copy->device();
        mDevValid = true;
        noteUpload(copy);
    }
    // Download the device copy on a copy stream once the calling party's stream has produced it; waitHost() (or any
    // host access) completes it.  The host vector keeps its address when the shape is unchanged.
    void fetchHostAsync(gpu::Context* copy) {
        if (!mDevValid || !size()) return;
        touchDev();
        mHost.resize(size());
        void* produced = gpu::current()->recordEvent();
        gpu::check(aby3cu_event_wait(copy->h(), produced));
        gpu::current()->recycleEvent(produced);
        gpu::check(aby3cu_d2h(copy->h(), mHost.data(), mDev.ptr(), size() * sizeof(T)));
        dropEvent(mHostReady, mHostReadyDevice);
        mHostReady = copy->recordEvent();
        mHostReadyDevice = copy->device();
        mHostValid = true;
    }
    void waitHost() const {
        if (!mHostReady) return;
        gpu::check(aby3cu_event_sync(mHostReady));
        gpu::EventPool::put(mHostReadyDevice, mHostReady);
        mHostReady = nullptr;
    }

    // -------------------------------------------------------- fills ----------
    void setZero() {
        if (mDevValid && !mHostValid) {
            gpu::check(aby3cu_memset(ctx()->h(), mDev.ptr(), 0, size() * sizeof(T)));
        } else {
            settleUpload();
            mHost.assign(size(), T{});
            mHostValid = true; mDevValid = false;
        }
    }
    void setConstant(T v) {
        touchHost(true);
        std::fill(mHost.begin(), mHost.end(), v);
    }

    // ------------------------------------------------------ arithmetic -------
    eMatrix operator+(const eMatrix& b) const { return binary(b, ABY3CU_OP_ADD); }
    eMatrix operator-(const eMatrix& b) const { return binary(b, ABY3CU_OP_SUB); }
    eMatrix& operator+=(const eMatrix& b) { return inplace(b, ABY3CU_OP_ADD); }
    eMatrix& operator-=(const eMatrix& b) { return inplace(b, ABY3CU_OP_SUB); }
    eMatrix operator-() const {
        eMatrix z(mRows, mCols);
        z.setZero();                       // explicit: a fresh matrix has no defined contents on either side
        return z - *this;
    }
    // plaintext matrix product (wrapping), used by callers for expected values
    eMatrix operator*(const eMatrix& b) const {
        if (mCols != b.mRows) throw std::runtime_error("eMatrix product: shape mismatch " LOCATION);
        eMatrix c(mRows, b.mCols);
        const T* A = hostData();
        const T* B = b.hostData();
        T* C = c.data();
        for (u64 i = 0; i < mRows; ++i)
            for (u64 k = 0; k < mCols; ++k) {
                const u64 a = (u64)A[i * mCols + k];
                for (u64 j = 0; j < b.mCols; ++j)
                    C[i * b.mCols + j] = (T)((u64)C[i * b.mCols + j] + a * (u64)B[k * b.mCols + j]);
            }
        return c;
    }
    eMatrix& operator*=(const eMatrix& b) { *this = *this * b; return *this; }
    eMatrix operator*(T s) const {
        eMatrix c(mRows, mCols);
        const T* A = hostData();
        T* C = c.data();
        for (u64 i = 0; i < size(); ++i) C[i] = (T)((u64)A[i] * (u64)s);
        return c;
    }

    eMatrix transpose() const {
        eMatrix t;
        t.mRows = mCols; t.mCols = mRows;
        if constexpr (kDeviceOps) {
            if (mDevValid) {
                gpu::check(aby3cu_transpose_i64(ctx()->h(), (const int64_t*)dev(), mRows, mCols, (int64_t*)t.devOut()));
                return t;
            }
        }
        t.mHost.assign(size(), T{});
        t.mHostValid = true;
        const T* A = hostData();
        for (u64 i = 0; i < mRows; ++i)
            for (u64 j = 0; j < mCols; ++j) t.mHost[j * mRows + i] = A[i * mCols + j];
        return t;
    }
    void transposeInPlace() { *this = transpose(); }

    // Both share planes of a replicated sharing at once: (o0, o1) = (a0 op b0, a1 op b1) as ONE kernel
    // launch when the operands live on the device (latency-bound protocol loops count launches).
    // o may be the same object as a (in-place) but not as b.
    static void binary2(const eMatrix& a0, const eMatrix& b0, eMatrix& o0, const eMatrix& a1, const eMatrix& b1, eMatrix& o1, int op) {
        if (a0.mRows != b0.mRows || a0.mCols != b0.mCols || a1.mRows != b1.mRows || a1.mCols != b1.mCols ||
            a0.mRows != a1.mRows || a0.mCols != a1.mCols)
            throw std::runtime_error("eMatrix: shape mismatch " LOCATION);
        if constexpr (kDeviceOps) {
            if ((a0.mDevValid || b0.mDevValid) && (a1.mDevValid || b1.mDevValid) && a0.size()) {
                const int64_t* x0 = (const int64_t*)a0.dev();
                const int64_t* y0 = (const int64_t*)b0.dev();
                const int64_t* x1 = (const int64_t*)a1.dev();
                const int64_t* y1 = (const int64_t*)b1.dev();
                auto outPtr = [](eMatrix& o, const eMatrix& a) -> int64_t* {
                    if (&o == &a) return (int64_t*)o.devMut();
                    o.resize(a.mRows, a.mCols);
                    return (int64_t*)o.devOut();
                };
                int64_t* z0 = outPtr(o0, a0);
                int64_t* z1 = outPtr(o1, a1);
                gpu::check(aby3cu_share_op2(a0.ctx()->h(), op, x0, y0, z0, x1, y1, z1, a0.size()));
                return;
            }
        }
        if (&o0 == &a0) o0.inplace(b0, op); else o0 = a0.binary(b0, op);
        if (&o1 == &a1) o1.inplace(b1, op); else o1 = a1.binary(b1, op);
    }
    // (t0, t1) = (a0^T, a1^T) in one launch; t may be the same object as a
    static void transpose2(const eMatrix& a0, const eMatrix& a1, eMatrix& t0, eMatrix& t1) {
        if constexpr (kDeviceOps) {
            if (a0.mDevValid && a1.mDevValid && a0.size() && a0.mRows == a1.mRows && a0.mCols == a1.mCols) {
                eMatrix r0, r1;
                r0.mRows = a0.mCols; r0.mCols = a0.mRows;
                r1.mRows = a1.mCols; r1.mCols = a1.mRows;
                gpu::check(aby3cu_transpose_i64_2(a0.ctx()->h(), (const int64_t*)a0.dev(), (const int64_t*)a1.dev(), a0.mRows, a0.mCols,
                                                  (int64_t*)r0.devOut(), (int64_t*)r1.devOut()));
                t0 = std::move(r0);
                t1 = std::move(r1);
                return;
            }
        }
        eMatrix r0 = a0.transpose(), r1 = a1.transpose();
        t0 = std::move(r0);
        t1 = std::move(r1);
    }

    bool operator==(const eMatrix& b) const {
        if (mRows != b.mRows || mCols != b.mCols) return false;
        return size() == 0 || memcmp(hostData(), b.hostData(), size() * sizeof(T)) == 0;
    }
    bool operator!=(const eMatrix& b) const { return !(*this == b); }

    // ------------------------------------------ row / col / block proxies ----
    struct RowRef {
        eMatrix& m; u64 i;
        RowRef& operator=(const RowRef& o) { for (u64 j = 0; j < m.cols(); ++j) m(i, j) = o.m(o.i, j); return *this; }
        template <typename R, typename = decltype(std::declval<const R&>().m)>
        RowRef& operator=(const R& o) { for (u64 j = 0; j < m.cols(); ++j) m(i, j) = o.m(o.i, j); return *this; }
        T& operator()(u64 j) { return m(i, j); }
    };
    struct ColRef {
        eMatrix& m; u64 j;
        ColRef& operator=(const ColRef& o) { for (u64 i = 0; i < m.rows(); ++i) m(i, j) = o.m(i, o.j); return *this; }
        template <typename R, typename = decltype(std::declval<const R&>().m)>
        ColRef& operator=(const R& o) { for (u64 i = 0; i < m.rows(); ++i) m(i, j) = o.m(i, o.j); return *this; }
        T& operator()(u64 i) { return m(i, j); }
    };
    RowRef row(u64 i) { return RowRef{*this, i}; }
    ColRef col(u64 j) { return ColRef{*this, j}; }
    // read-only rows / columns of a const matrix (XX.row(i) = X.row(k), aby3-ML/Regression.h:55-56)
    struct ConstRowRef { const eMatrix& m; u64 i; const T& operator()(u64 j) const { return m(i, j); } };
    struct ConstColRef { const eMatrix& m; u64 j; const T& operator()(u64 i) const { return m(i, j); } };
    ConstRowRef row(u64 i) const { return ConstRowRef{*this, i}; }
    ConstColRef col(u64 j) const { return ConstColRef{*this, j}; }
    eMatrix block(u64 r, u64 c, u64 h, u64 w) const {
        eMatrix b(h, w);
        const T* A = hostData();
        T* B = b.data();
        for (u64 i = 0; i < h; ++i)
            for (u64 j = 0; j < w; ++j) B[i * w + j] = A[(r + i) * mCols + c + j];
        return b;
    }
    // a block of a non-const matrix is a writable view (`res.block(...) = data;`, aby3-Basic/Basic.cpp:48) that also
    // converts to a matrix (`i64Matrix part = m.block(...);`, :17)
    struct BlockRef {
        eMatrix& m; u64 r, c, h, w;
        BlockRef& operator=(const eMatrix& src) {
            if (src.rows() != h || src.cols() != w) throw std::runtime_error("eMatrix: block assignment shape mismatch " LOCATION);
            const T* S = src.hostData();
            T* D = m.data();
            for (u64 i = 0; i < h; ++i)
                for (u64 j = 0; j < w; ++j) D[(r + i) * m.mCols + c + j] = S[i * w + j];
            return *this;
        }
        BlockRef& operator=(const BlockRef& o) { return *this = eMatrix(o); }
        operator eMatrix() const { return static_cast<const eMatrix&>(m).block(r, c, h, w); }
        u64 rows() const { return h; }
        u64 cols() const { return w; }
        T& operator()(u64 i, u64 j) { return m(r + i, c + j); }
        BlockRef& setConstant(T v) {
            T* D = m.data();
            for (u64 i = 0; i < h; ++i)
                for (u64 j = 0; j < w; ++j) D[(r + i) * m.mCols + c + j] = v;
            return *this;
        }
        BlockRef& setZero() { return setConstant(T{}); }
    };
    BlockRef block(u64 r, u64 c, u64 h, u64 w) { return BlockRef{*this, r, c, h, w}; }

private:
    void ensureDevBuffer() {
        const size_t bytes = std::max<size_t>(size() * sizeof(T), 16);
        if (!mDev || mDev.bytes() < bytes) mDev.reset(gpu::current(), bytes);
    }
    static void dropEvent(void*& e, int device) {
        if (e) { aby3cu_event_sync(e); gpu::EventPool::put(device, e); e = nullptr; }
    }
    // a prefetch on a copy stream: order the calling party's stream behind it
    void orderBehindPrefetch() const {
        if (!mDevReady) return;
        gpu::check(aby3cu_event_wait(gpu::current()->h(), mDevReady));
        gpu::EventPool::put(mDevReadyDevice, mDevReady);
        mDevReady = nullptr;
    }
    void touchDev() const {
        orderBehindPrefetch();
        if (mDevValid) return;
        auto* self = const_cast<eMatrix*>(this);
        self->ensureDevBuffer();
        // cudaMemcpyAsync from pageable memory is staged before it returns, so the
        // host vector may be modified right after this call.
        if (mHostValid && size()) {
            gpu::check(aby3cu_h2d(mDev.ctx()->h(), mDev.ptr(), mHost.data(), size() * sizeof(T)));
            // large host copies are page-locked (gpu::HostAllocator): the upload is truly
            // asynchronous and the host copy must not be written before it has run
            if (size() * sizeof(T) >= gpu::HostAllocator<T>::kPinThreshold) noteUpload(mDev.ctx());
        } else if (size()) {
            // never written on either side: the host path reads such a matrix as zeros (touchHost zero-fills), so the
            // device path must too -- a recycled pool block holds someone else's old data
            gpu::check(aby3cu_memset(mDev.ctx()->h(), mDev.ptr(), 0, size() * sizeof(T)));
        }
        self->mDevValid = true;
    }
    void touchHost(bool willWrite) const {
        auto* self = const_cast<eMatrix*>(this);
        waitHost();
        if (!mHostValid) {
            if (mDevValid) self->mHost.resize(size());        // no fill: the d2h below overwrites it
            else self->mHost.assign(size(), T{});
            if (mDevValid && size()) {
                gpu::check(aby3cu_d2h(mDev.ctx()->h(), self->mHost.data(), mDev.ptr(), size() * sizeof(T)));
                mDev.ctx()->sync();
            }
            self->mHostValid = true;
        }
        if (willWrite) {
            dropEvent(self->mDevReady, mDevReadyDevice);          // a prefetch still reads the host copy
            waitUpload();
            self->mDevValid = false;
        }
    }
    // page-locked host blocks are recycled (gpu::PinnedPool): an asynchronous upload
    // reading this matrix's host copy must have run before that copy is released
    void settleUpload() {
        dropEvent(mDevReady, mDevReadyDevice);
        dropEvent(mHostReady, mHostReadyDevice);
        try { waitUpload(); } catch (...) {}
        mUploadInFlight = false;
    }
    // An asynchronous upload reads the page-locked host copy: remember an event right behind it, so that a later host WRITE
    // (or the release of the host block) waits for that copy alone.  (It used to synchronise the party's whole stream: in a
    // pipelined caller "touch the next step's input buffer" then meant "wait for the step that is running", and the next
    // step's uploads never overlapped anything -- tools/e2e_trace.py.)
    void noteUpload(gpu::Context* c) const {
        if (mUploadDone) { gpu::EventPool::put(mUploadDoneDevice, mUploadDone); mUploadDone = nullptr; }
        mUploadDone = c->recordEvent();
        mUploadDoneDevice = c->device();
        mUploadInFlight = true;
    }
    void waitUpload() const {
        if (mUploadDone) {
            gpu::check(aby3cu_event_sync(mUploadDone));
            gpu::EventPool::put(mUploadDoneDevice, mUploadDone);
            mUploadDone = nullptr;
        } else if (mUploadInFlight && mDev.ctx()) {
            mDev.ctx()->sync();
        }
        mUploadInFlight = false;
    }
    void copyFrom(const eMatrix& o) {
        settleUpload();
        mRows = o.mRows; mCols = o.mCols;
        if (o.mDevValid && !o.mHostValid) {
            mHost.clear(); mHostValid = false;
            mDev.reset(o.mDev.ctx(), std::max<size_t>(size() * sizeof(T), 16));
            if (size())
                gpu::check(aby3cu_d2d(mDev.ctx()->h(), mDev.ptr(), mDev.ctx()->device(), o.mDev.ptr(),
                                      o.mDev.ctx()->device(), size() * sizeof(T)));
            mDevValid = true;
        } else {
            o.waitHost();
            mHost = o.mHost; mHostValid = o.mHostValid; mDevValid = false;
            mDev.free();
        }
    }
    void moveFrom(eMatrix&& o) {
        settleUpload();
        mRows = o.mRows; mCols = o.mCols;
        mHost = std::move(o.mHost); mHostValid = o.mHostValid;
        mDev = std::move(o.mDev); mDevValid = o.mDevValid;
        mUploadInFlight = o.mUploadInFlight; o.mUploadInFlight = false;
        mUploadDone = o.mUploadDone; mUploadDoneDevice = o.mUploadDoneDevice; o.mUploadDone = nullptr;
        mDevReady = o.mDevReady; mDevReadyDevice = o.mDevReadyDevice; o.mDevReady = nullptr;
        mHostReady = o.mHostReady; mHostReadyDevice = o.mHostReadyDevice; o.mHostReady = nullptr;
        o.mRows = o.mCols = 0; o.mHostValid = false; o.mDevValid = false; o.mHost.clear();
    }
    eMatrix binary(const eMatrix& b, int op) const {
        if (mRows != b.mRows || mCols != b.mCols) throw std::runtime_error("eMatrix: shape mismatch " LOCATION);
        eMatrix c;
        c.mRows = mRows; c.mCols = mCols;
        if constexpr (kDeviceOps) {
            if (mDevValid || b.mDevValid) {
                const int64_t* x = (const int64_t*)dev();
                const int64_t* y = (const int64_t*)b.dev();
                gpu::check(aby3cu_share_op(ctx()->h(), op, x, y, (int64_t*)c.devOut(), size()));
                return c;
            }
        }
        c.mHost.resize(size());
        c.mHostValid = true;
        const T* A = hostData();
        const T* B = b.hostData();
        for (u64 i = 0; i < size(); ++i)
            c.mHost[i] = op == ABY3CU_OP_ADD ? (T)((u64)A[i] + (u64)B[i]) : (T)((u64)A[i] - (u64)B[i]);
        return c;
    }
    eMatrix& inplace(const eMatrix& b, int op) {
        if (mRows != b.mRows || mCols != b.mCols) throw std::runtime_error("eMatrix: shape mismatch " LOCATION);
        if constexpr (kDeviceOps) {
            if (mDevValid || b.mDevValid) {
                const int64_t* y = (const int64_t*)b.dev();
                int64_t* x = (int64_t*)devMut();
                gpu::check(aby3cu_share_op(ctx()->h(), op, x, y, x, size()));
                return *this;
            }
        }
        T* A = data();
        const T* B = b.hostData();
        for (u64 i = 0; i < size(); ++i)
            A[i] = op == ABY3CU_OP_ADD ? (T)((u64)A[i] + (u64)B[i]) : (T)((u64)A[i] - (u64)B[i]);
        return *this;
    }

    u64 mRows = 0, mCols = 0;
    mutable std::vector<T, gpu::HostAllocator<T>> mHost;
    mutable bool mHostValid = false;
    mutable gpu::Buffer mDev;
    mutable bool mDevValid = false;
    mutable bool mUploadInFlight = false;
    mutable void* mDevReady = nullptr;      // prefetch in flight on a copy stream
    mutable void* mHostReady = nullptr;     // download in flight on a copy stream
    mutable void* mUploadDone = nullptr;    // right behind the last asynchronous upload of the host copy (any stream)
    mutable int mDevReadyDevice = 0, mHostReadyDevice = 0, mUploadDoneDevice = 0;
};

template <typename T>
std::ostream& operator<<(std::ostream& o, const eMatrix<T>& m) {
    for (u64 i = 0; i < m.rows(); ++i) {
        for (u64 j = 0; j < m.cols(); ++j) o << m(i, j) << (j + 1 < m.cols() ? " " : "");
        o << "\n";
    }
    return o;
}

using i64Matrix = eMatrix<i64>;

}  // namespace aby3
