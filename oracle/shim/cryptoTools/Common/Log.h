#pragma once
// shim of cryptoTools/Common/Log.h: a locked cout, colours ignored
#include <mutex>
#include "cryptoTools/Common/Defines.h"
namespace osuCrypto {
enum class Color { LightGreen = 2, LightGrey = 3, LightRed = 4, OffWhite1 = 5, OffWhite2 = 6, Grey = 8, Green = 10, Blue = 11,
                   Red = 12, Pink = 13, Yellow = 14, White = 15, Default };
inline std::ostream& operator<<(std::ostream& o, Color) { return o; }
enum class IoStream { lock, unlock };
extern std::mutex gIoStreamMtx;
inline std::ostream& operator<<(std::ostream& o, IoStream) { return o; }
inline std::ostream& lock(std::ostream& o) { return o; }
inline std::ostream& unlock(std::ostream& o) { return o; }
struct ostreamLock {
    std::ostream& out;
    std::unique_lock<std::mutex> mLock;
    explicit ostreamLock(std::ostream& o) : out(o), mLock(gIoStreamMtx) {}
    template <typename T>
    ostreamLock& operator<<(T&& v) { out << std::forward<T>(v); return *this; }
    ostreamLock& operator<<(std::ostream& (*v)(std::ostream&)) { out << v; return *this; }
};
struct ostreamLocker {
    std::ostream& out;
    explicit ostreamLocker(std::ostream& o) : out(o) {}
    template <typename T>
    ostreamLocker& operator<<(T&& v) { std::lock_guard<std::mutex> g(gIoStreamMtx); out << std::forward<T>(v); return *this; }
    ostreamLocker& operator<<(std::ostream& (*v)(std::ostream&)) { std::lock_guard<std::mutex> g(gIoStreamMtx); out << v; return *this; }
};
extern ostreamLocker lout;
struct LogAdapter {
    template <typename T>
    LogAdapter& operator<<(const T&) { return *this; }
    LogAdapter& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
extern LogAdapter gLog;
struct Log {
    std::mutex mMtx;
    std::vector<std::string> mMsgs;
    void push(const std::string& s) { std::lock_guard<std::mutex> g(mMtx); mMsgs.push_back(s); }
};
inline std::ostream& operator<<(std::ostream& o, Log& l) { for (auto& m : l.mMsgs) o << m << "\n"; return o; }
void setThreadName(const std::string&);
}  // namespace osuCrypto
