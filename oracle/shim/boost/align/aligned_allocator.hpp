#pragma once
// shim of boost/align/aligned_allocator.hpp (Sh3BinaryEvaluator.h:137)
#include <cstddef>
#include <new>
namespace boost { namespace alignment {
template <typename T, std::size_t Align = alignof(T)>
struct aligned_allocator {
    typedef T value_type;
    aligned_allocator() = default;
    template <typename U> aligned_allocator(const aligned_allocator<U, Align>&) {}
    template <typename U> struct rebind { typedef aligned_allocator<U, Align> other; };
    T* allocate(std::size_t n) { return static_cast<T*>(::operator new(n * sizeof(T), std::align_val_t(Align < 64 ? 64 : Align))); }
    void deallocate(T* p, std::size_t) { ::operator delete(p, std::align_val_t(Align < 64 ? 64 : Align)); }
    template <typename U> bool operator==(const aligned_allocator<U, Align>&) const { return true; }
    template <typename U> bool operator!=(const aligned_allocator<U, Align>&) const { return false; }
};
} }
