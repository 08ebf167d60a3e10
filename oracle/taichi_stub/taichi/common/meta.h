// TEST INFRASTRUCTURE — see taichi/util.h.  Compile-time branches: TC_STATIC_IF(c) { A } TC_STATIC_ELSE { B }
// TC_STATIC_END_IF instantiates only the taken branch (generic lambdas), as the core's meta.h does.
#pragma once
#include <utility>
namespace taichi {
namespace stub_meta {
struct Identity { template <class T> decltype(auto) operator()(T &&x) const { return std::forward<T>(x); } };
template <bool C> struct StaticIf;
template <> struct StaticIf<true> {
  template <class F> explicit StaticIf(F &&f) { f(Identity()); }
  template <class F> void else_(F &&) {}
};
template <> struct StaticIf<false> {
  template <class F> explicit StaticIf(F &&) {}
  template <class F> void else_(F &&f) { f(Identity()); }
};
}  // namespace stub_meta
}  // namespace taichi
// C++17: the untaken branch of a value-dependent condition is not instantiated, which is what the call sites
// rely on (e.g. a 2-argument Linear_Offset inside TC_STATIC_IF(dim == 2), src/mpm.cpp:836-846)
#define TC_STATIC_IF(x) if constexpr (x) {
#define TC_STATIC_ELSE } else {
#define TC_STATIC_END_IF }
#define TC_REPEAT27(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16) M(17) M(18) M(19) M(20) M(21) M(22) M(23) M(24) M(25) M(26)
