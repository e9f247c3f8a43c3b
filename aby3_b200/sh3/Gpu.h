// Gpu.h -- the facade's only door to the device: a per-party context over the
// C ABI (include/aby3cu.h) with a small stream-ordered buffer pool.  There is no
// host implementation behind it; without libaby3cu.so + a B200 every call throws.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "Defines.h"
#include "aby3cu.h"

namespace aby3 {
namespace gpu {

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("aby3cu: ") + aby3cu_last_error());
}

// Ordering-only CUDA events, recycled per device: a message hand-over between parties must not pay
// for cudaEventCreate / cudaEventDestroy.  An event may be re-recorded as soon as every wait on it
// has been ISSUED (cudaStreamWaitEvent captures the record that precedes the call).
class EventPool {
public:
    static void* get(aby3cu_ctx* ctx, int device) {
        auto& p = inst();
        {
            std::lock_guard<std::mutex> g(p.mMtx);
            auto& v = p.mFree[device];
            if (!v.empty()) { void* e = v.back(); v.pop_back(); return e; }
        }
        void* e = nullptr;
        check(aby3cu_event_create_sync(ctx, &e));
        return e;
    }
    static void put(int device, void* e) {
        if (!e) return;
        auto& p = inst();
        {
            std::lock_guard<std::mutex> g(p.mMtx);
            auto& v = p.mFree[device];
            if (v.size() < 8192) { v.push_back(e); return; }
        }
        aby3cu_event_destroy(e);
    }
private:
    static EventPool& inst() { static EventPool* p = new EventPool; return *p; }   // leaked on purpose: outlives every context
    std::mutex mMtx;
    std::map<int, std::vector<void*>> mFree;
};

// One party's device, stream and buffer pool.  Buffers released to the pool are
// reused in stream order; a buffer another party's stream may still be reading
// is parked together with the event that marks the end of that read.
class Context {
public:
    explicit Context(int device = 0) {
        check(aby3cu_ctx_create(device, &mCtx));
        mDevice = device;
    }
    Context(int device, void* stream) {
        check(aby3cu_ctx_create_on_stream(device, stream, &mCtx));
        mDevice = device;
    }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ~Context() {
        if (!mCtx) return;
        if (mAux) aby3cu_sync(mAux->h());
        if (mComm) aby3cu_sync(mComm->h());
        aby3cu_sync(mCtx);
        mAux.reset();
        mComm.reset();
        for (auto& kv : mFree)
            for (auto& e : kv.second) {
                for (auto& ev : e.events) EventPool::put(ev.first, ev.second);
                if (e.pos) EventPool::put(mDevice, e.pos);
                driverFree(e.ptr, kv.first, /*noexcept*/ true);
            }
        if (guardEnabled())
            std::fprintf(stderr, "[pool guard] device %d: %llu block checks, %llu damaged\n", mDevice,
                         (unsigned long long)mGuardChecks, (unsigned long long)mGuardBad);
        aby3cu_ctx_destroy(mCtx);
    }

    aby3cu_ctx* h() const { return mCtx; }
    int device() const { return mDevice; }
    void sync() {
        if (mAux) check(aby3cu_sync(mAux->h()));
        if (mComm) check(aby3cu_sync(mComm->h()));
        check(aby3cu_sync(mCtx));
    }

    // ---- the party's second compute stream ------------------------------------------------------------------
    // Input-independent work of a protocol step (the truncation pair of a fixed-point product: keystream only)
    // is issued here so that it runs under the PREVIOUS step's kernels instead of behind them.  Buffers it writes
    // come from allocEarly(); joinAux() orders the party's own stream behind everything issued on this one.
    Context* aux() {
        if (!mAux) {
            mAux.reset(new Context(mDevice));
            static const bool corun = [] { const char* e = std::getenv("ABY3_AUX_CORUN"); return !(e && e[0] == '0'); }();
            check(aby3cu_ctx_set_corun(mAux->h(), corun ? 1 : 0));
        }
        return mAux.get();
    }
    bool hasAux() const { return (bool)mAux; }
    // ---- the party's communication stream: transfers that overlap the party's own kernels (block-wise opens) --------
    Context* comm() {
        if (!mComm) mComm.reset(new Context(mDevice));
        return mComm.get();
    }
    void joinComm() {
        if (!mComm) return;
        void* e = mComm->recordEvent();
        check(aby3cu_event_wait(mCtx, e));
        mComm->recycleEvent(e);
    }
    void joinAux() {
        if (!mAux) return;
        void* e = mAux->recordEvent();
        check(aby3cu_event_wait(mCtx, e));
        mAux->recycleEvent(e);
    }
    // kernels launched on this context's stream and on its second stream
    u64 launchCount() const { return aby3cu_launch_count(mCtx) + (mAux ? aby3cu_launch_count(mAux->h()) : 0); }
    // no ordering events are recorded at release while the stream is being captured into a graph
    void setCapturing(bool on) { mCapturing = on; }

    void* alloc(size_t bytes) {
        if (bytes == 0) return nullptr;
        bytes = roundSize(bytes);
        Entry e{nullptr, {}, nullptr};
        {
            std::lock_guard<std::mutex> g(mMtx);
            auto it = mFree.find(bytes);
            if (it != mFree.end() && !it->second.empty()) {
                e = std::move(it->second.back());
                it->second.pop_back();
                mCached -= bytes;
            }
        }
        if (e.ptr) {
            for (auto& ev : e.events) {                      // readers on other streams: waited for only now, at reuse
                check(aby3cu_event_wait(mCtx, ev.second));
                EventPool::put(ev.first, ev.second);
            }
            if (e.pos) EventPool::put(mDevice, e.pos);       // this stream is behind its own earlier position anyway
            if (guardEnabled()) guardReuse(this, e.ptr, bytes);
            return e.ptr;
        }
        return fresh(bytes);
    }
    // A block whose first writer runs on the aux stream.  Blocks of kEarlyMin bytes and more are parked together
    // with the position of the party's stream at their release; the aux stream waits for that position (and for a
    // foreign reader, if any) instead of for everything the party has enqueued since.  The OLDEST parked block is
    // taken and the kEarlySpare newest ones are always left in place (the newest block's last reader typically ends the
    // step that is still executing), so the pool settles kEarlySpare blocks per size class in use above what in-order
    // recycling would need.  (Measured with tools/step_trace.py: the pairs become ready when the step before them ends --
    // they leave the gap between two products but do not yet run under the previous product's GEMMs; a deeper rotation,
    // kEarlySpare = 3, did not change that.)
    void* allocEarly(size_t bytes) {
        if (bytes == 0) return nullptr;
        bytes = roundSize(bytes);
        Entry e{nullptr, {}, nullptr};
        {
            std::lock_guard<std::mutex> g(mMtx);
            auto it = mFree.find(bytes);
            if (it != mFree.end() && it->second.size() > kEarlySpare) {
                e = std::move(it->second.front());
                it->second.erase(it->second.begin());
                mCached -= bytes;
            }
        }
        if (!e.ptr) return fresh(bytes);
        Context* a = aux();
        for (auto& ev : e.events) {
            check(aby3cu_event_wait(a->h(), ev.second));
            EventPool::put(ev.first, ev.second);
        }
        if (e.pos) {
            check(aby3cu_event_wait(a->h(), e.pos));
            EventPool::put(mDevice, e.pos);
        } else {
            // parked without a position (released under capture): order behind the whole stream
            void* now = recordEvent();
            check(aby3cu_event_wait(a->h(), now));
            EventPool::put(mDevice, now);
        }
        if (guardEnabled()) guardReuse(a, e.ptr, bytes);
        return e.ptr;
    }
    // give every cached block back to the driver (between workloads with different buffer sizes)
    void trim() {
        std::map<size_t, std::vector<Entry>> drop;
        {
            std::lock_guard<std::mutex> g(mMtx);
            drop.swap(mFree);
            mCached = 0;
        }
        if (mAux) aby3cu_sync(mAux->h());
        aby3cu_sync(mCtx);
        for (auto& kv : drop)
            for (auto& e : kv.second) {
                for (auto& ev : e.events) { aby3cu_event_sync(ev.second); EventPool::put(ev.first, ev.second); }
                if (e.pos) EventPool::put(mDevice, e.pos);
                driverFree(e.ptr, kv.first);
                ++mFrees;
            }
    }
    // pool misses (driver allocations) since creation: a steady-state loop should show none
    u64 mallocCount() const { return mMallocs; }
    u64 mallocBytes() const { return mMallocBytes; }
    u64 freeCount() const { return mFrees; }
    // `after` (may be null): a pooled event recorded on ANOTHER stream (of device `afterDevice`) that still
    // reads the buffer
    void release(void* p, size_t bytes, void* after = nullptr, int afterDevice = -1) {
        std::vector<std::pair<int, void*>> readers;
        if (after) readers.emplace_back(afterDevice < 0 ? mDevice : afterDevice, after);
        releaseShared(p, bytes, std::move(readers));
    }
    // the same for a block several other streams read (SharedBuffer): their events travel with the block and are waited
    // for when it is handed out again, not now -- the party's stream must not stall behind its neighbours' readers
    void releaseShared(void* p, size_t bytes, std::vector<std::pair<int, void*>>&& readers) {
        if (!p) { for (auto& ev : readers) EventPool::put(ev.first, ev.second); return; }
        bytes = roundSize(bytes);
        // where this party's stream stands now: everything the party itself enqueued on the block precedes it
        void* pos = (bytes >= kEarlyMin && !mCapturing && earlyEnabled()) ? recordEvent() : nullptr;
        {
            std::lock_guard<std::mutex> g(mMtx);
            if (mCached + bytes <= kCacheCap) {
                mFree[bytes].push_back(Entry{p, std::move(readers), pos});
                mCached += bytes;
                return;
            }
        }
        // the cache is full: give the block back to the driver (workloads whose buffer sizes keep
        // changing -- e.g. the shrinking stages of a merge network -- must not hoard HBM)
        for (auto& ev : readers) { aby3cu_event_sync(ev.second); EventPool::put(ev.first, ev.second); }
        if (pos) EventPool::put(mDevice, pos);
        if (guardEnabled()) { if (mAux) aby3cu_sync(mAux->h()); aby3cu_sync(mCtx); }
        driverFree(p, bytes);
        ++mFrees;
    }
    // a recycled ordering event for this context's device / recorded on this context's stream
    void* newEvent() { return EventPool::get(mCtx, mDevice); }
    void* recordEvent() { void* e = newEvent(); check(aby3cu_event_record(mCtx, e)); return e; }
    void recycleEvent(void* e) { EventPool::put(mDevice, e); }
    void* stream() const { return aby3cu_ctx_stream(mCtx); }

    // Size classes: multiples of 512 B up to 1 MiB, then 1/8-octave steps, so that buffers whose
    // sizes drift (stage after stage of a merge network) still recycle each other's memory.
    static size_t roundSize(size_t b) {
        if (b <= (size_t(1) << 20)) return (b + 511) & ~size_t(511);
        size_t step = size_t(1) << 17;
        while ((step << 4) < b) step <<= 1;          // step = 2^k with 8*step < b <= 16*step
        return (b + step - 1) / step * step;
    }

    static constexpr size_t kEarlyMin = size_t(4) << 20;
    static constexpr size_t kEarlySpare = 1;
    // ABY3_EARLY_TRUNCATION=0: nothing is issued ahead on the second stream, so releases record no positions either
    static bool guardEnabled() {
        static const bool on = [] { const char* e = std::getenv("ABY3_POOL_GUARD"); return e && e[0] == '1'; }();
        return on;
    }
    static bool earlyEnabled() {
        static const bool on = [] { const char* e = std::getenv("ABY3_EARLY_TRUNCATION"); return !(e && e[0] == '0'); }();
        return on;
    }

private:
    // events: pooled events (device, event) of OTHER streams that still read the block; pos: this party's stream at release
    struct Entry { void* ptr; std::vector<std::pair<int, void*>> events; void* pos; };
    void* fresh(size_t bytes) {
        void* p = nullptr;
        const size_t total = bytes + (guardEnabled() ? 2 * kGuard : 0);
        if (aby3cu_malloc(mCtx, &p, total) != 0) {
            // HBM is exhausted by blocks this pool keeps for reuse: hand them back and try once more
            trim();
            check(aby3cu_malloc(mCtx, &p, total));
        }
        ++mMallocs; mMallocBytes += bytes;
        if (guardEnabled()) {
            check(aby3cu_memset(mCtx, p, kCanary, kGuard));
            check(aby3cu_memset(mCtx, (char*)p + kGuard + bytes, kCanary, kGuard));
            check(aby3cu_memset(mCtx, (char*)p + kGuard, kPoison, bytes));
            p = (char*)p + kGuard;
        }
        return p;
    }
    // ---- ABY3_POOL_GUARD=1 (debug): what compute-sanitizer's memcheck / initcheck would look for, done by the pool ------
    // (the reference's own equivalent is the poison value mCheckBlock of Sh3BinaryEvaluator.cpp:578-621).
    // Every block carries kGuard canary bytes on both sides, verified whenever the block changes hands or goes back to
    // the driver: a kernel writing outside its buffer trips it.  A block handed out again is first filled with a poison
    // pattern on the stream that will write it -- AFTER the waits on every recorded reader -- so a reader the pool failed
    // to order (the early-free path hands blocks to the second stream while other parties may still read them) sees
    // poison instead of its data and the share-level parity tests fail.
    void checkGuards(void* p, size_t bytes, bool quiet = false) {
        unsigned char h[2 * kGuard];
        if (aby3cu_d2h(mCtx, h, (char*)p - kGuard, kGuard) || aby3cu_d2h(mCtx, h + kGuard, (char*)p + bytes, kGuard) || aby3cu_sync(mCtx)) {
            if (quiet) return;
            check(1);
        }
        ++mGuardChecks;
        for (size_t i = 0; i < 2 * kGuard; ++i)
            if (h[i] != kCanary) {
                ++mGuardBad;
                char msg[160];
                std::snprintf(msg, sizeof msg, "pool guard: %zu-byte block %p was written %s its bounds (offset %zd)", bytes, p,
                              i < kGuard ? "BEFORE" : "PAST", i < kGuard ? (ptrdiff_t)i - (ptrdiff_t)kGuard : (ptrdiff_t)(i - kGuard));
                std::fprintf(stderr, "%s\n", msg);
                if (!quiet) throw std::runtime_error(msg);
                return;
            }
    }
    // `on`: the context whose stream has just been ordered behind every recorded reader of the block
    void guardReuse(Context* on, void* p, size_t bytes) {
        check(aby3cu_sync(on->h()));
        checkGuards(p, bytes);
        check(aby3cu_memset(on->h(), p, kPoison, bytes));
    }
    void driverFree(void* p, size_t bytes, bool quiet = false) {
        if (guardEnabled()) {
            checkGuards(p, bytes, quiet);
            p = (char*)p - kGuard;
        }
        aby3cu_free(mCtx, p);
    }
    static constexpr size_t kGuard = 512;
    static constexpr int kCanary = 0xA5, kPoison = 0xCD;
    u64 mGuardChecks = 0, mGuardBad = 0;
    aby3cu_ctx* mCtx = nullptr;
    int mDevice = 0;
    std::unique_ptr<Context> mAux, mComm;
    bool mCapturing = false;
    std::mutex mMtx;
    std::map<size_t, std::vector<Entry>> mFree;
    size_t mCached = 0;
    u64 mMallocs = 0, mMallocBytes = 0, mFrees = 0;
    static constexpr size_t kCacheCap = size_t(24) << 30;
};

// The context of the calling thread (one thread per party, as in the reference:
// frontend/aby3Tutorial.cpp:393-394).  Sh3Runtime::init adopts the context of
// its CommPkg; tests may set it directly.
Context*& currentSlot();
inline Context* current() {
    Context* c = currentSlot();
    if (!c) throw std::runtime_error("aby3::gpu: no device context bound to this thread (call gpu::setCurrent)");
    return c;
}
inline void setCurrent(Context* c) { currentSlot() = c; }

struct Early {};

// RAII device allocation drawn from a context's pool
class Buffer {
public:
    Buffer() = default;
    Buffer(Context* c, size_t bytes) { reset(c, bytes); }
    // first written on c's aux stream (Context::allocEarly)
    Buffer(Context* c, size_t bytes, Early) : mCtx(c), mPtr(c->allocEarly(bytes)), mBytes(bytes) {}
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
    Buffer(Buffer&& o) noexcept { *this = std::move(o); }
    Buffer& operator=(Buffer&& o) noexcept {
        if (this != &o) {
            free();
            mCtx = o.mCtx; mPtr = o.mPtr; mBytes = o.mBytes;
            o.mCtx = nullptr; o.mPtr = nullptr; o.mBytes = 0;
        }
        return *this;
    }
    ~Buffer() { free(); }
    void reset(Context* c, size_t bytes) {
        if (mCtx == c && mBytes >= bytes && mBytes <= 2 * Context::roundSize(bytes)) return;
        free();
        mCtx = c; mBytes = bytes; mPtr = c->alloc(bytes);
    }
    // `after`: a pooled ordering event of device `afterDevice` marking the last foreign read
    void free(void* after = nullptr, int afterDevice = -1) {
        if (mPtr && mCtx) mCtx->release(mPtr, mBytes, after, afterDevice);
        else if (after) aby3cu_event_destroy(after);        // no owner to hand it to
        mPtr = nullptr; mBytes = 0; mCtx = nullptr;
    }
    // readers: (device, pooled event) pairs marking the last reads by other streams
    void freeShared(std::vector<std::pair<int, void*>>&& readers) {
        if (mPtr && mCtx) mCtx->releaseShared(mPtr, mBytes, std::move(readers));
        else for (auto& ev : readers) EventPool::put(ev.first, ev.second);
        mPtr = nullptr; mBytes = 0; mCtx = nullptr;
    }
    void* ptr() const { return mPtr; }
    size_t bytes() const { return mBytes; }
    Context* ctx() const { return mCtx; }
    explicit operator bool() const { return mPtr != nullptr; }
private:
    Context* mCtx = nullptr;
    void* mPtr = nullptr;
    size_t mBytes = 0;
};

// A device buffer several parties read: the producer sends it WITHOUT a staging copy, every reader
// reports the event after its last read, and the block returns to the producer's pool together with
// those events: whoever takes it next waits for them.  Contents are immutable once shared.
class SharedBuffer {
public:
    SharedBuffer(Context* c, size_t bytes) : mBuf(c, bytes) {}
    SharedBuffer(Context* c, size_t bytes, Early e) : mBuf(c, bytes, e) {}
    SharedBuffer(const SharedBuffer&) = delete;
    SharedBuffer& operator=(const SharedBuffer&) = delete;
    ~SharedBuffer() {
        std::lock_guard<std::mutex> g(mMtx);
        mBuf.freeShared(std::move(mReaders));        // the readers' events go to the pool with the block
    }
    void* ptr() const { return mBuf.ptr(); }
    size_t bytes() const { return mBuf.bytes(); }
    Context* ctx() const { return mBuf.ctx(); }
    // `event` (from reader's device pool) marks the end of a reader's last access
    void addReader(int device, void* event) {
        std::lock_guard<std::mutex> g(mMtx);
        mReaders.emplace_back(device, event);
    }
private:
    Buffer mBuf;
    std::mutex mMtx;
    std::vector<std::pair<int, void*>> mReaders;
};

// Page-locked host blocks are expensive to create (cudaHostAlloc is a syscall-heavy
// path), so freed blocks are kept for reuse, up to a cap.
class PinnedPool {
public:
    static PinnedPool& get() { static PinnedPool p; return p; }
    void* alloc(size_t bytes) {
        bytes = (bytes + 4095) & ~size_t(4095);
        {
            std::lock_guard<std::mutex> g(mMtx);
            auto it = mFree.find(bytes);
            if (it != mFree.end() && !it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                mCached -= bytes;
                return p;
            }
        }
        void* p = nullptr;
        if (aby3cu_host_alloc(&p, bytes) != 0) throw std::bad_alloc();
        return p;
    }
    void release(void* p, size_t bytes) {
        bytes = (bytes + 4095) & ~size_t(4095);
        {
            std::lock_guard<std::mutex> g(mMtx);
            if (mCached + bytes <= kCap) {
                mFree[bytes].push_back(p);
                mCached += bytes;
                return;
            }
        }
        aby3cu_host_free(p);
    }
private:
    static constexpr size_t kCap = size_t(4) << 30;
    std::mutex mMtx;
    std::map<size_t, std::vector<void*>> mFree;
    size_t mCached = 0;
};

// Host storage of matrices: page-locked once it is big enough to matter, so the
// h2d / d2h copies behind eMatrix run at full PCIe rate and truly asynchronously.
template <typename T>
struct HostAllocator {
    using value_type = T;
    static constexpr size_t kPinThreshold = 1u << 16;
    HostAllocator() = default;
    template <typename U>
    HostAllocator(const HostAllocator<U>&) {}
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes >= kPinThreshold) return static_cast<T*>(PinnedPool::get().alloc(bytes));
        void* p = ::operator new(bytes);
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t n) {
        if (n * sizeof(T) >= kPinThreshold) PinnedPool::get().release(p, n * sizeof(T));
        else ::operator delete(p);
    }
    // value-initialisation is skipped for trivial types: resize() must not sweep
    // hundreds of MiB that a d2h copy is about to overwrite
    template <typename U>
    void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename... Args>
    void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
    template <typename U>
    bool operator==(const HostAllocator<U>&) const { return true; }
    template <typename U>
    bool operator!=(const HostAllocator<U>&) const { return false; }
};

}  // namespace gpu
}  // namespace aby3
