// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for ericniebler/meta.
// The reference includes "meta/meta.hpp" (include/factors_messages.hxx:24,
// include/template_utilities.hxx:10) for type-list metaprogramming; call sites:
// include/LP_MP.h:233,261,488-494, include/factors_messages.hxx:87,606-607,655,932-935,2164,
// 2202,2393-2401,2516,3269,3381-3428, include/solver.hxx:54,215,356-357.
// The submodule is not vendored; this shim provides only the operations those sites use.
// It is our own code, not a copy of the meta library.
#pragma once
#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace meta {

template <typename... Ts>
struct list {
  using type = list;
  static constexpr std::size_t size() noexcept { return sizeof...(Ts); }
};

template <std::size_t N> using size_t = std::integral_constant<std::size_t, N>;
template <bool B> using bool_ = std::integral_constant<bool, B>;
using npos = size_t<std::size_t(-1)>;

template <typename L> struct size_impl;
template <typename... Ts> struct size_impl<list<Ts...>> : size_t<sizeof...(Ts)> {};
template <typename L> using size = size_impl<L>;

// at_c
template <typename L, std::size_t N> struct at_impl;
template <typename T, typename... Ts> struct at_impl<list<T, Ts...>, 0> { using type = T; };
template <typename T, typename... Ts, std::size_t N> struct at_impl<list<T, Ts...>, N> : at_impl<list<Ts...>, N - 1> {};
template <typename L, std::size_t N> using at_c = typename at_impl<L, N>::type;

// find_index: index of first occurrence, npos when absent
template <typename L, typename T, std::size_t I> struct find_index_impl;
template <typename T, std::size_t I> struct find_index_impl<list<>, T, I> : npos {};
template <typename T, typename... Ts, std::size_t I> struct find_index_impl<list<T, Ts...>, T, I> : size_t<I> {};
template <typename U, typename... Ts, typename T, std::size_t I>
struct find_index_impl<list<U, Ts...>, T, I> : find_index_impl<list<Ts...>, T, I + 1> {};
template <typename L, typename T> using find_index = find_index_impl<L, T, 0>;

// invoke a "callable" (class with nested template alias `invoke`)
template <typename F, typename... Args> using invoke = typename F::template invoke<Args...>;

// quote
template <template <typename...> class C>
struct quote {
  template <typename... Ts> using invoke = C<Ts...>;
};

// apply<F, list<Ts...>> = F::invoke<Ts...>
template <typename F, typename L> struct apply_impl;
template <typename F, typename... Ts> struct apply_impl<F, list<Ts...>> { using type = typename F::template invoke<Ts...>; };
template <typename F, typename L> using apply = typename apply_impl<F, L>::type;

// transform
template <typename L, typename F> struct transform_impl;
template <typename... Ts, typename F> struct transform_impl<list<Ts...>, F> { using type = list<typename F::template invoke<Ts>...>; };
template <typename L, typename F> using transform = typename transform_impl<L, F>::type;

// concat
template <typename... Ls> struct concat_impl;
template <> struct concat_impl<> { using type = list<>; };
template <typename... Ts> struct concat_impl<list<Ts...>> { using type = list<Ts...>; };
template <typename... Ts, typename... Us, typename... Rest>
struct concat_impl<list<Ts...>, list<Us...>, Rest...> : concat_impl<list<Ts..., Us...>, Rest...> {};
template <typename... Ls> using concat = typename concat_impl<Ls...>::type;

// filter
template <typename L, typename P> struct filter_impl;
template <typename P> struct filter_impl<list<>, P> { using type = list<>; };
template <typename T, typename... Ts, typename P>
struct filter_impl<list<T, Ts...>, P> {
  using rest = typename filter_impl<list<Ts...>, P>::type;
  using type = typename std::conditional<P::template invoke<T>::value, concat<list<T>, rest>, rest>::type;
};
template <typename L, typename P> using filter = typename filter_impl<L, P>::type;

// unique (keeps first occurrences)
template <typename Acc, typename L> struct unique_impl;
template <typename Acc> struct unique_impl<Acc, list<>> { using type = Acc; };
template <typename... As, typename T, typename... Ts>
struct unique_impl<list<As...>, list<T, Ts...>>
    : unique_impl<typename std::conditional<(find_index<list<As...>, T>::value == npos::value), list<As..., T>, list<As...>>::type,
                  list<Ts...>> {};
template <typename L> using unique = typename unique_impl<list<>, L>::type;

// any_of
template <typename L, typename P> struct any_of_impl;
template <typename P> struct any_of_impl<list<>, P> : std::false_type {};
template <typename T, typename... Ts, typename P>
struct any_of_impl<list<T, Ts...>, P> : bool_<(P::template invoke<T>::value || any_of_impl<list<Ts...>, P>::value)> {};
template <typename L, typename P> using any_of = any_of_impl<L, P>;

template <bool B, typename T, typename E> using if_c = typename std::conditional<B, T, E>::type;

// runtime for_each over the types of a list: calls f(T{}) for each T in order
template <typename... Ts, typename F>
constexpr F for_each(list<Ts...>, F f) {
  using swallow = int[];
  (void)swallow{0, ((void)f(Ts{}), 0)...};
  return f;
}

}  // namespace meta
