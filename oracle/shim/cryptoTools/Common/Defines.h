#pragma once
// shim of cryptoTools/Common/Defines.h (absent third-party header; see ../../README.md)
#include <algorithm>
#include <array>
#include <cassert>
#include <functional>
#include <future>
#include <list>
#include <utility>
#include <cstdint>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#include <emmintrin.h>
#include <immintrin.h>

#define OC_STRINGIZE_DETAIL(x) #x
#define OC_STRINGIZE(x) OC_STRINGIZE_DETAIL(x)
#define LOCATION __FILE__ ":" OC_STRINGIZE(__LINE__)
#define RTE_LOC std::runtime_error(LOCATION)
#define TODO(x)

namespace osuCrypto {
typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef uint16_t u16;
typedef int16_t i16;
typedef uint8_t u8;
typedef int8_t i8;

// contiguous view (gsl::span in the original)
template <typename T>
class span {
public:
    typedef T value_type;
    typedef T* iterator;
    span() = default;
    span(T* p, u64 n) : mP(p), mN(n) {}
    span(T* b, T* e) : mP(b), mN(u64(e - b)) {}
    template <typename C, typename = decltype(std::declval<C&>().data()),
              typename = typename std::enable_if<std::is_convertible<decltype(std::declval<C&>().data()), T*>::value>::type>
    span(C& c) : mP(c.data()), mN(c.size()) {}
    template <typename C, typename = decltype(std::declval<const C&>().data()),
              typename = typename std::enable_if<std::is_convertible<decltype(std::declval<const C&>().data()), T*>::value>::type>
    span(const C& c) : mP(c.data()), mN(c.size()) {}
    T* data() const { return mP; }
    u64 size() const { return mN; }
    u64 size_bytes() const { return mN * sizeof(T); }
    bool empty() const { return mN == 0; }
    T* begin() const { return mP; }
    T* end() const { return mP + mN; }
    T& operator[](u64 i) const { return mP[i]; }
    T& front() const { return mP[0]; }
    T& back() const { return mP[mN - 1]; }
    span subspan(u64 off, u64 n = ~0ull) const { return span(mP + off, n == ~0ull ? mN - off : n); }
private:
    T* mP = nullptr;
    u64 mN = 0;
};

// 128-bit block; bytes 0..7 = low word (little endian), 8..15 = high word
struct alignas(16) block {
    __m128i mData;
    block() = default;
    block(const __m128i& x) : mData(x) {}
    block(u64 hi, u64 lo) : mData(_mm_set_epi64x((long long)hi, (long long)lo)) {}
    operator const __m128i&() const { return mData; }
    operator __m128i&() { return mData; }
    block operator^(const block& o) const { return _mm_xor_si128(mData, o.mData); }
    block operator&(const block& o) const { return _mm_and_si128(mData, o.mData); }
    block operator|(const block& o) const { return _mm_or_si128(mData, o.mData); }
    block& operator^=(const block& o) { mData = _mm_xor_si128(mData, o.mData); return *this; }
    bool operator==(const block& o) const { return std::memcmp(this, &o, 16) == 0; }
    bool operator!=(const block& o) const { return !(*this == o); }
    bool operator<(const block& o) const { return std::memcmp(this, &o, 16) < 0; }
    template <typename T>
    T as() const { T t; static_assert(sizeof(T) == 16, ""); std::memcpy(&t, this, 16); return t; }
};
static_assert(sizeof(block) == 16, "");

inline block toBlock(u64 hi, u64 lo) { return block(hi, lo); }
inline block toBlock(u64 lo) { return block(0, lo); }
inline block toBlock(const u8* p) { block b; std::memcpy(&b, p, 16); return b; }
static const block ZeroBlock = toBlock(0, 0);
static const block OneBlock = toBlock(0, 1);
static const block AllOneBlock = toBlock(~0ull, ~0ull);
static const block CCBlock = toBlock(0xccccccccccccccccull, 0xccccccccccccccccull);

inline std::ostream& operator<<(std::ostream& o, const block& b) {
    const u8* p = (const u8*)&b;
    std::ios_base::fmtflags f(o.flags());
    for (int i = 15; i >= 0; --i) o << std::hex << std::setw(2) << std::setfill('0') << int(p[i]);
    o.flags(f);
    return o;
}

inline u64 roundUpTo(u64 v, u64 d) { return (v + d - 1) / d * d; }
inline u64 divCeil(u64 v, u64 d) { return (v + d - 1) / d; }
inline u64 log2ceil(u64 v) { u64 r = 0; while ((1ull << r) < v) ++r; return r; }
inline u64 log2floor(u64 v) { u64 r = 0; while ((2ull << r) <= v) ++r; return r; }

block sysRandomSeed();

enum class AllocType { Uninitialized, Zeroed };

template <typename T, typename = void>
struct is_container : std::false_type {};
template <typename T>
struct is_container<T, typename std::enable_if<
    std::is_convertible<decltype(std::declval<T&>().data()), const void*>::value &&
    std::is_convertible<decltype(std::declval<T&>().size()), u64>::value>::type> : std::true_type {};

template <typename T, typename = void>
struct is_resizable_container : std::false_type {};
template <typename T>
struct is_resizable_container<T, typename std::enable_if<
    is_container<T>::value && std::is_same<decltype(std::declval<T&>().resize(0)), void>::value>::type> : std::true_type {};
}  // namespace osuCrypto
namespace oc = osuCrypto;
