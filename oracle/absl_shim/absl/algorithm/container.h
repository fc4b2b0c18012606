#pragma once
#include <algorithm>
#include <iterator>
#include <numeric>
namespace absl {
template <class C, class T> void c_fill(C&& c, const T& v) { std::fill(std::begin(c), std::end(c), v); }
template <class C> void c_sort(C& c) { std::sort(std::begin(c), std::end(c)); }
template <class C, class F> void c_sort(C& c, F f) { std::sort(std::begin(c), std::end(c), f); }
template <class C> void c_stable_sort(C& c) { std::stable_sort(std::begin(c), std::end(c)); }
template <class C, class F> void c_stable_sort(C& c, F f) { std::stable_sort(std::begin(c), std::end(c), f); }
template <class C> bool c_is_sorted(const C& c) { return std::is_sorted(std::begin(c), std::end(c)); }
template <class C, class F> bool c_is_sorted(const C& c, F f) { return std::is_sorted(std::begin(c), std::end(c), f); }
template <class C, class T> auto c_find(C& c, const T& v) { return std::find(std::begin(c), std::end(c), v); }
template <class C, class F> auto c_find_if(C& c, F f) { return std::find_if(std::begin(c), std::end(c), f); }
template <class C, class T> auto c_count(const C& c, const T& v) { return std::count(std::begin(c), std::end(c), v); }
template <class C, class F> auto c_count_if(const C& c, F f) { return std::count_if(std::begin(c), std::end(c), f); }
template <class C, class F> bool c_all_of(const C& c, F f) { return std::all_of(std::begin(c), std::end(c), f); }
template <class C, class F> bool c_any_of(const C& c, F f) { return std::any_of(std::begin(c), std::end(c), f); }
template <class C, class F> bool c_none_of(const C& c, F f) { return std::none_of(std::begin(c), std::end(c), f); }
template <class C, class O> auto c_copy(const C& c, O o) { return std::copy(std::begin(c), std::end(c), o); }
template <class C, class T> auto c_lower_bound(C& c, const T& v) { return std::lower_bound(std::begin(c), std::end(c), v); }
template <class C, class T, class F> auto c_lower_bound(C& c, const T& v, F f) { return std::lower_bound(std::begin(c), std::end(c), v, f); }
template <class C, class T> auto c_upper_bound(C& c, const T& v) { return std::upper_bound(std::begin(c), std::end(c), v); }
template <class C, class T> bool c_binary_search(const C& c, const T& v) { return std::binary_search(std::begin(c), std::end(c), v); }
template <class C> void c_reverse(C& c) { std::reverse(std::begin(c), std::end(c)); }
template <class C> auto c_max_element(C& c) { return std::max_element(std::begin(c), std::end(c)); }
template <class C> auto c_min_element(C& c) { return std::min_element(std::begin(c), std::end(c)); }
template <class C1, class C2> bool c_equal(const C1& a, const C2& b) { return std::equal(std::begin(a), std::end(a), std::begin(b), std::end(b)); }
template <class C, class T> T c_accumulate(const C& c, T init) { return std::accumulate(std::begin(c), std::end(c), init); }
template <class C, class T> void c_iota(C& c, T v) { std::iota(std::begin(c), std::end(c), v); }
template <class C, class T> bool c_linear_search(const C& c, const T& v) { return std::find(std::begin(c), std::end(c), v) != std::end(c); }
template <class C, class F> auto c_for_each(C&& c, F f) { return std::for_each(std::begin(c), std::end(c), f); }
}
