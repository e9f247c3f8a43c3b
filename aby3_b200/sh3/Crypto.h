// Crypto.h -- oc::PRNG / oc::AES as the sh3 code uses them (cryptoTools, absent
// from the reference tree; restated per SURVEY section 8c):
//   PRNG(seed): one contiguous AES-128-CTR keystream AES_seed(0)||AES_seed(1)||...
//   consumed through a byte cursor; get<T>() / get(ptr, n) copy the next bytes.
// Small draws (keys, scalars) are produced on the host by the C ABI's bounded
// key-draw helper; bulk draws go to the device (aby3cu_aes_ctr_fill) -- the host
// never runs the hot loop.
#pragma once
#include "Gpu.h"
#include "BitVector.h"

namespace oc {

class PRNG {
public:
    PRNG() = default;
    explicit PRNG(const block& seed, u64 /*bufferSize*/ = 256) { SetSeed(seed); }
    void SetSeed(const block& seed, u64 /*bufferSize*/ = 256) { mSeed = seed; mByteIdx = 0; mSet = true; }
    block getSeed() const { return mSeed; }
    u64 byteCursor() const { return mByteIdx; }
    // advance without producing output (the bytes were consumed by a fused device kernel)
    void skip(u64 bytes) { mByteIdx += bytes; }

    template <typename T>
    T get() {
        T v;
        getBytes(reinterpret_cast<u8*>(&v), sizeof(T));
        return v;
    }
    block get() { return get<block>(); }
    template <typename T>
    void get(T* dst, u64 n) { getBytes(reinterpret_cast<u8*>(dst), n * sizeof(T)); }
    template <typename T>
    void get(span<T> s) { getBytes(reinterpret_cast<u8*>(s.data()), s.size() * sizeof(T)); }

    // usable as a standard random source and as std::random_shuffle's functor (aby3-ML/Regression.h:36)
    typedef u32 result_type;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return (result_type)-1; }
    result_type operator()() { return get<result_type>(); }
    u64 operator()(u64 mod) { return get<u64>() % mod; }

    // bulk draw straight into device memory (offset and size multiples of 8)
    void getDevice(aby3::gpu::Context* ctx, void* d_dst, u64 bytes) {
        requireSeed();
        aby3::gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mSeed.data(), mByteIdx, d_dst, bytes));
        mByteIdx += bytes;
    }

private:
    void requireSeed() const { if (!mSet) throw std::runtime_error("PRNG used before SetSeed " LOCATION); }
    void getBytes(u8* dst, u64 n) {
        requireSeed();
        aby3::gpu::Context* ctx = aby3::gpu::currentSlot();
        if (n >= (1u << 16) && ctx && mByteIdx % 8 == 0 && n % 8 == 0) {
            // large host-side draw (plaintext test inputs): generate on the device, copy back
            aby3::gpu::Buffer tmp(ctx, n);
            aby3::gpu::check(aby3cu_aes_ctr_fill(ctx->h(), mSeed.data(), mByteIdx, tmp.ptr(), n));
            aby3::gpu::check(aby3cu_d2h(ctx->h(), dst, tmp.ptr(), n));
            ctx->sync();
            mByteIdx += n;
            return;
        }
        while (n) {
            const u64 step = n < 4096 ? n : 4096;
            aby3::gpu::check(aby3cu_host_keystream(mSeed.data(), mByteIdx, step, dst));
            dst += step; n -= step; mByteIdx += step;
        }
    }
    block mSeed;
    u64 mByteIdx = 0;
    bool mSet = false;
};

// oc::AES as used by Sh3ShareGen / Sh3BinaryEvaluator: a key holder; the
// encryption itself happens inside the device kernels.
class AES {
public:
    AES() = default;
    explicit AES(const block& k) { setKey(k); }
    void setKey(const block& k) { mKey = k; }
    const block& key() const { return mKey; }
    // host-side counter-mode blocks for the scalar paths (<= 256 blocks per call)
    void ecbEncCounterMode(u64 baseIdx, u64 nblocks, block* out) const {
        while (nblocks) {
            const u64 step = nblocks < 256 ? nblocks : 256;
            aby3::gpu::check(aby3cu_host_keystream(mKey.data(), baseIdx * 16, step * 16, reinterpret_cast<u8*>(out)));
            out += step; baseIdx += step; nblocks -= step;
        }
    }
private:
    block mKey;
};

// a fresh non-deterministic seed (cryptoTools sysRandomSeed(); aby3-Basic/BuildingBlocks.cpp:64)
block sysRandomSeed();

}  // namespace oc
