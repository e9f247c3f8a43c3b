#pragma once
// compat_surface.h -- the odds and ends of cryptoTools / Eigen that unchanged reference APPLICATION code names
// (aby3-ML, aby3-Basic), none of it on the data path: oc::lout / Color / ostreamLock (aby3-ML/Regression.h:182,292),
// namespace osuCrypto, Eigen::internal::Packet4i (aby3-Basic/debug.h:43).
#include <cassert>
#include <cmath>
#include <iostream>
#include <mutex>

#include "aby3_b200/sh3/Defines.h"

namespace osuCrypto = oc;

namespace oc {
enum class Color { LightGreen = 2, LightGrey = 3, LightRed = 4, OffWhite1 = 5, OffWhite2 = 6, Grey = 8, Green = 10, Blue = 11,
                   Red = 12, Pink = 13, Yellow = 14, White = 15, Default };
inline std::ostream& operator<<(std::ostream& o, Color) { return o; }
inline std::mutex& ioStreamMutex() { static std::mutex m; return m; }
// a std::cout whose statements do not interleave between party threads
struct ostreamLock {
    std::ostream& out;
    std::unique_lock<std::mutex> mLock;
    explicit ostreamLock(std::ostream& o) : out(o), mLock(ioStreamMutex()) {}
    template <typename T>
    ostreamLock& operator<<(T&& v) { out << std::forward<T>(v); return *this; }
    ostreamLock& operator<<(std::ostream& (*v)(std::ostream&)) { out << v; return *this; }
};
struct ostreamLocker {
    std::ostream& out;
    explicit ostreamLocker(std::ostream& o) : out(o) {}
    template <typename T>
    ostreamLocker& operator<<(T&& v) { std::lock_guard<std::mutex> g(ioStreamMutex()); out << std::forward<T>(v); return *this; }
    ostreamLocker& operator<<(std::ostream& (*v)(std::ostream&)) { std::lock_guard<std::mutex> g(ioStreamMutex()); out << v; return *this; }
};
static ostreamLocker lout(std::cout);
}  // namespace oc

namespace Eigen {
namespace internal {
struct Packet4i { int v[4]; };
inline std::ostream& operator<<(std::ostream& o, const Packet4i& p) { return o << p.v[0] << " " << p.v[1] << " " << p.v[2] << " " << p.v[3]; }
}  // namespace internal
}  // namespace Eigen
