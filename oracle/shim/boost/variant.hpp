#pragma once
// shim of boost/variant.hpp: a move-only tagged holder with boost::get<T>(variant*) -> T* (nullptr when
// another alternative is active), as used by Sh3Runtime.h:85-91 / Sh3Runtime.cpp:94,121,190,288-289.
#include <tuple>
#include <type_traits>
#include <utility>
namespace boost {
namespace shim_detail {
template <typename T, typename... Ts> struct index_of;
template <typename T, typename... Ts> struct index_of<T, T, Ts...> : std::integral_constant<int, 0> {};
template <typename T, typename U, typename... Ts> struct index_of<T, U, Ts...> : std::integral_constant<int, 1 + index_of<T, Ts...>::value> {};
template <typename T, typename... Ts> struct contains : std::false_type {};
template <typename T, typename U, typename... Ts> struct contains<T, U, Ts...> : std::integral_constant<bool, std::is_same<T, U>::value || contains<T, Ts...>::value> {};
}
template <typename... Ts>
class variant {
    int mWhich = -1;
    std::tuple<Ts...> mVals;       // every alternative is default-constructible and cheap when empty
public:
    variant() = default;
    variant(variant&&) = default;
    variant& operator=(variant&&) = default;
    template <typename T, typename D = typename std::decay<T>::type, typename = typename std::enable_if<shim_detail::contains<D, Ts...>::value>::type>
    variant(T&& v) { *this = std::forward<T>(v); }
    template <typename T, typename D = typename std::decay<T>::type, typename = typename std::enable_if<shim_detail::contains<D, Ts...>::value>::type>
    variant& operator=(T&& v) {
        mWhich = shim_detail::index_of<D, Ts...>::value;
        std::get<shim_detail::index_of<D, Ts...>::value>(mVals) = std::forward<T>(v);
        return *this;
    }
    int which() const { return mWhich; }
    template <typename T>
    T* ptr() { return mWhich == shim_detail::index_of<T, Ts...>::value ? &std::get<shim_detail::index_of<T, Ts...>::value>(mVals) : nullptr; }
};
template <typename T, typename... Ts>
T* get(variant<Ts...>* v) { return v->template ptr<T>(); }
}  // namespace boost
