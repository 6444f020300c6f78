// TEST INFRASTRUCTURE ONLY (oracle/): scalar stand-in for libsimdpp.
// The reference includes "simdpp/simd.h" (include/config.hxx:15-16) and uses a handful of
// AVX2 wrappers in include/vector.hxx:192-194,286,302-307,330-341,355-422,583-590,868-943.
// The submodule is not vendored; this shim implements those wrappers with plain lane loops
// (min/max/add are exact per lane, so results are identical to the SIMD library's).
// It is our own code, not a copy of libsimdpp.
#pragma once
#include <algorithm>
#include <cstddef>

#ifndef SIMDPP_ARCH_X86_AVX2
#define SIMDPP_ARCH_X86_AVX2 1
#endif

namespace simdpp {

template <typename T, unsigned N>
struct lanes {
  T v[N];
  T vec(unsigned i) const { return v[i]; }
  T& vec(unsigned i) { return v[i]; }
};

template <unsigned N> struct float64 : lanes<double, N> {
  float64() {}
  template <typename P, typename = decltype(std::declval<P>().template get<double, N>())>
  float64(const P& p) { *this = p.template get<double, N>(); }
};
template <unsigned N> struct float32 : lanes<float, N> {
  float32() {}
  template <typename P, typename = decltype(std::declval<P>().template get<float, N>())>
  float32(const P& p) { *this = p.template get<float, N>(); }
};

namespace detail {
template <typename T, unsigned N> struct vec_of;
template <unsigned N> struct vec_of<double, N> { using type = float64<N>; };
template <unsigned N> struct vec_of<float, N> { using type = float32<N>; };

template <typename S>
struct load_proxy {  // deferred load: lane count and type come from the destination
  const S* p;
  template <typename T, unsigned N> typename vec_of<T, N>::type get() const {
    typename vec_of<T, N>::type r;
    for (unsigned i = 0; i < N; ++i) r.v[i] = static_cast<T>(p[i]);
    return r;
  }
};
template <typename S>
struct splat_proxy {
  S x;
  template <typename T, unsigned N> typename vec_of<T, N>::type get() const {
    typename vec_of<T, N>::type r;
    for (unsigned i = 0; i < N; ++i) r.v[i] = static_cast<T>(x);
    return r;
  }
};
}  // namespace detail

template <typename S> inline detail::load_proxy<S> load(const S* p) { return {p}; }
template <typename S> inline detail::load_proxy<S> load(S* p) { return {p}; }
template <typename S> inline detail::splat_proxy<S> load_splat(const S* p) { return {*p}; }
template <typename S> inline detail::splat_proxy<S> load_splat(S* p) { return {*p}; }
template <typename S> inline detail::splat_proxy<S> make_float(S x) { return {x}; }

template <typename S, typename T, unsigned N> inline void store_impl(S* p, const lanes<T, N>& a) {
  for (unsigned i = 0; i < N; ++i) p[i] = a.v[i];
}
template <typename S, unsigned N> inline void store(S* p, const float64<N>& a) { store_impl(p, a); }
template <typename S, unsigned N> inline void store(S* p, const float32<N>& a) { store_impl(p, a); }

#define LPMP_SHIM_BINOP(NAME, EXPR)                                                     \
  template <unsigned N> inline float64<N> NAME(const float64<N>& a, const float64<N>& b) { \
    float64<N> r; for (unsigned i = 0; i < N; ++i) { const double x = a.v[i], y = b.v[i]; r.v[i] = (EXPR); } return r; } \
  template <unsigned N> inline float32<N> NAME(const float32<N>& a, const float32<N>& b) { \
    float32<N> r; for (unsigned i = 0; i < N; ++i) { const float x = a.v[i], y = b.v[i]; r.v[i] = (EXPR); } return r; }
LPMP_SHIM_BINOP(min, (y < x ? y : x))
LPMP_SHIM_BINOP(max, (x < y ? y : x))
LPMP_SHIM_BINOP(operator+, x + y)
#undef LPMP_SHIM_BINOP

template <typename T, unsigned N> inline T reduce_min_impl(const lanes<T, N>& a) {
  T m = a.v[0]; for (unsigned i = 1; i < N; ++i) m = std::min(m, a.v[i]); return m;
}
template <typename T, unsigned N> inline T reduce_max_impl(const lanes<T, N>& a) {
  T m = a.v[0]; for (unsigned i = 1; i < N; ++i) m = std::max(m, a.v[i]); return m;
}
template <unsigned N> inline double reduce_min(const float64<N>& a) { return reduce_min_impl(a); }
template <unsigned N> inline float reduce_min(const float32<N>& a) { return reduce_min_impl(a); }
template <unsigned N> inline double reduce_max(const float64<N>& a) { return reduce_max_impl(a); }
template <unsigned N> inline float reduce_max(const float32<N>& a) { return reduce_max_impl(a); }

template <unsigned I, unsigned N> inline double extract(const float64<N>& a) { return a.v[I]; }
template <unsigned I, unsigned N> inline float extract(const float32<N>& a) { return a.v[I]; }

template <unsigned N> inline void split(const float32<N>& a, float32<N / 2>& lo, float32<N / 2>& hi) {
  for (unsigned i = 0; i < N / 2; ++i) { lo.v[i] = a.v[i]; hi.v[i] = a.v[N / 2 + i]; }
}
template <unsigned N> inline void split(const float64<N>& a, float64<N / 2>& lo, float64<N / 2>& hi) {
  for (unsigned i = 0; i < N / 2; ++i) { lo.v[i] = a.v[i]; hi.v[i] = a.v[N / 2 + i]; }
}

template <typename P> inline void prefetch_read(const P* p) { __builtin_prefetch(p); }

}  // namespace simdpp
