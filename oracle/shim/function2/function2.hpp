#pragma once
// shim of function2: fu2::unique_function<Sig> as a move-only type-erased callable
#include <memory>
#include <utility>
namespace fu2 {
template <typename Sig>
class unique_function;
template <typename R, typename... A>
class unique_function<R(A...)> {
    struct Base { virtual ~Base() {} virtual R call(A... a) = 0; };
    template <typename F>
    struct Impl : Base {
        F f;
        explicit Impl(F&& x) : f(std::move(x)) {}
        R call(A... a) override { return f(std::forward<A>(a)...); }
    };
    std::unique_ptr<Base> mP;
public:
    unique_function() = default;
    unique_function(std::nullptr_t) {}
    unique_function(unique_function&&) = default;
    unique_function& operator=(unique_function&&) = default;
    unique_function(const unique_function&) = delete;
    unique_function& operator=(const unique_function&) = delete;
    template <typename F, typename D = typename std::decay<F>::type,
              typename = typename std::enable_if<!std::is_same<D, unique_function>::value>::type,
              typename = decltype(std::declval<D&>()(std::declval<A>()...))>
    unique_function(F&& f) : mP(new Impl<D>(D(std::forward<F>(f)))) {}
    explicit operator bool() const { return (bool)mP; }
    R operator()(A... a) { return mP->call(std::forward<A>(a)...); }
    R operator()(A... a) const { return mP->call(std::forward<A>(a)...); }
};
}  // namespace fu2
