#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
#include <vector>
namespace absl {
struct HashState {
  uint64_t h = 0xcbf29ce484222325ull;
  void mix(uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
  template <class T> static void one(HashState& s, const T& v);
  template <class... A> static HashState combine(HashState s, const A&... a) { (one(s, a), ...); return s; }
  template <class T> static HashState combine_contiguous(HashState s, const T* p, size_t n) { for (size_t i = 0; i < n; ++i) one(s, p[i]); return s; }
};
template <class H, class T> H AbslHashValue(H h, const std::vector<T>& v) { return H::combine(H::combine_contiguous(std::move(h), v.data(), v.size()), v.size()); }
template <class H, class A, class B> H AbslHashValue(H h, const std::pair<A, B>& p) { return H::combine(std::move(h), p.first, p.second); }
template <class T> void HashState::one(HashState& s, const T& v) {
  if constexpr (std::is_floating_point_v<T>) s.mix(std::hash<T>{}(v));
  else if constexpr (std::is_integral_v<T> || std::is_enum_v<T> || std::is_pointer_v<T>) s.mix((uint64_t)v);
  else if constexpr (std::is_same_v<T, std::string> || std::is_same_v<T, std::string_view>) s.mix(std::hash<std::string_view>{}(v));
  else s = AbslHashValue(std::move(s), v);
}
template <class T> struct Hash { size_t operator()(const T& v) const { HashState s; HashState::one(s, v); return (size_t)s.h; } };
template <class... A> size_t HashOf(const A&... a) { return (size_t)HashState::combine(HashState{}, a...).h; }
}
