// Defines.h -- basic types of the sh3 facade.
// Mirrors aby3/Common/Defines.h:1-70 and the handful of cryptoTools primitives
// the hot path touches (oc::block / toBlock / span / divCeil / roundUpTo), which
// live in libOTe @ cf537295 and are absent from the reference tree.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#define ABY3_STRINGIZE_DETAIL(x) #x
#define ABY3_STRINGIZE(x) ABY3_STRINGIZE_DETAIL(x)
#ifndef LOCATION
#define LOCATION __FILE__ ":" ABY3_STRINGIZE(__LINE__)
#endif
#ifndef RTE_LOC
#define RTE_LOC std::runtime_error(LOCATION)
#endif

namespace oc {

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint16_t u16;
typedef uint8_t u8;
typedef int8_t i8;

// 16 raw bytes.  toBlock(hi, lo) follows _mm_set_epi64x(hi, lo): bytes 0..7 are
// lo (little endian), bytes 8..15 are hi.
struct block {
    u64 lo = 0, hi = 0;
    block() = default;
    block(u64 hi_, u64 lo_) : lo(lo_), hi(hi_) {}
    bool operator==(const block& o) const { return lo == o.lo && hi == o.hi; }
    bool operator!=(const block& o) const { return !(*this == o); }
    block operator^(const block& o) const { return block(hi ^ o.hi, lo ^ o.lo); }
    const u8* data() const { return reinterpret_cast<const u8*>(this); }
    u8* data() { return reinterpret_cast<u8*>(this); }
};
static_assert(sizeof(block) == 16, "block must be 16 bytes");

inline block toBlock(u64 hi, u64 lo) { return block(hi, lo); }
inline block toBlock(u64 lo) { return block(0, lo); }
static const block ZeroBlock = block(0, 0);
static const block OneBlock = block(0, 1);
static const block AllOneBlock = block(~0ull, ~0ull);

inline std::ostream& operator<<(std::ostream& o, const block& b) {
    char buf[40];
    snprintf(buf, sizeof(buf), "%016llx%016llx", (unsigned long long)b.hi, (unsigned long long)b.lo);
    return o << buf;
}

inline u64 divCeil(u64 a, u64 b) { return (a + b - 1) / b; }
inline u64 roundUpTo(u64 a, u64 b) { return divCeil(a, b) * b; }

// minimal span (gsl::span in cryptoTools)
template <typename T>
class span {
public:
    span() = default;
    // any integer length (gsl::span's index type is signed: callers write {ptr, i64(n)})
    template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type>
    span(T* p, I n) : mPtr(p), mSize((u64)n) {}
    template <typename C>
    span(C& c) : mPtr(c.data()), mSize(c.size()) {}
    T* data() const { return mPtr; }
    u64 size() const { return mSize; }
    T& operator[](u64 i) const { return mPtr[i]; }
    T* begin() const { return mPtr; }
    T* end() const { return mPtr + mSize; }
private:
    T* mPtr = nullptr;
    u64 mSize = 0;
};

}  // namespace oc

namespace aby3 {
using oc::u64; using oc::i64; using oc::u32; using oc::i32; using oc::u16; using oc::u8; using oc::i8;
using oc::block; using oc::span;

#define ABY3_ASSERT(cond) do { if (!(cond)) throw RTE_LOC; } while (0)
}  // namespace aby3
