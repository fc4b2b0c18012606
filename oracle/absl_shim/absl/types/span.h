#pragma once
#include <cstddef>
#include <initializer_list>
#include <type_traits>
#include <vector>
namespace absl {
template <class T> class Span {
 public:
  using value_type = std::remove_cv_t<T>; using pointer = T*; using const_pointer = const T*; using reference = T&; using const_reference = const T&;
  using iterator = T*; using const_iterator = const T*; using size_type = size_t; using element_type = T; using difference_type = ptrdiff_t;
  static constexpr size_type npos = ~size_type(0);
  constexpr Span() : p_(nullptr), n_(0) {}
  constexpr Span(T* p, size_t n) : p_(p), n_(n) {}
  template <size_t N> constexpr Span(T (&a)[N]) : p_(a), n_(N) {}
  template <class C, class = std::enable_if_t<!std::is_same_v<std::decay_t<C>, Span> && std::is_convertible_v<decltype(std::declval<C&>().data()), T*>>>
  constexpr Span(C& c) : p_(c.data()), n_(c.size()) {}
  template <class C, class = std::enable_if_t<std::is_const_v<T> && !std::is_same_v<std::decay_t<C>, Span> && std::is_convertible_v<decltype(std::declval<const C&>().data()), T*>>, class = void>
  constexpr Span(const C& c) : p_(c.data()), n_(c.size()) {}
  template <class U = T, class = std::enable_if_t<std::is_const_v<U>>>
  constexpr Span(std::initializer_list<value_type> l) : p_(l.begin()), n_(l.size()) {}
  constexpr T* data() const { return p_; } constexpr size_t size() const { return n_; } constexpr size_t length() const { return n_; }
  constexpr bool empty() const { return n_ == 0; }
  constexpr T& operator[](size_t i) const { return p_[i]; } constexpr T& at(size_t i) const { return p_[i]; }
  constexpr T& front() const { return p_[0]; } constexpr T& back() const { return p_[n_ - 1]; }
  constexpr T* begin() const { return p_; } constexpr T* end() const { return p_ + n_; }
  constexpr const T* cbegin() const { return p_; } constexpr const T* cend() const { return p_ + n_; }
  auto rbegin() const { return std::reverse_iterator<T*>(end()); } auto rend() const { return std::reverse_iterator<T*>(begin()); }
  void remove_prefix(size_t n) { p_ += n; n_ -= n; } void remove_suffix(size_t n) { n_ -= n; }
  constexpr Span subspan(size_t pos = 0, size_t len = npos) const { return Span(p_ + pos, len < n_ - pos ? len : n_ - pos); }
  constexpr Span first(size_t n) const { return Span(p_, n); } constexpr Span last(size_t n) const { return Span(p_ + n_ - n, n); }
  template <class H> friend H AbslHashValue(H h, Span s) { return H::combine(H::combine_contiguous(std::move(h), s.data(), s.size()), s.size()); }
 private:
  T* p_; size_t n_;
};
template <class T> bool operator==(Span<T> a, Span<T> b) { if (a.size() != b.size()) return false; for (size_t i = 0; i < a.size(); ++i) if (!(a[i] == b[i])) return false; return true; }
template <class T> bool operator!=(Span<T> a, Span<T> b) { return !(a == b); }
template <class T> Span<T> MakeSpan(T* p, size_t n) { return Span<T>(p, n); }
template <class T> Span<T> MakeSpan(T* b, T* e) { return Span<T>(b, (size_t)(e - b)); }
template <class C> auto MakeSpan(C& c) { return Span<std::remove_pointer_t<decltype(c.data())>>(c.data(), c.size()); }
template <class T> Span<const T> MakeConstSpan(const T* p, size_t n) { return Span<const T>(p, n); }
template <class C> auto MakeConstSpan(const C& c) { return Span<const std::remove_pointer_t<decltype(c.data())>>(c.data(), c.size()); }
}
