#pragma once
#include <bit>
#include <cstdint>
namespace absl {
using std::countr_zero; using std::countl_zero; using std::countr_one; using std::countl_one;
using std::bit_width; using std::rotr; using std::rotl; using std::popcount; using std::has_single_bit;
using std::bit_ceil; using std::bit_floor; using std::endian;
template <class T> constexpr T byteswap(T v) {
  if constexpr (sizeof(T) == 1) return v;
  else if constexpr (sizeof(T) == 2) return (T)__builtin_bswap16((uint16_t)v);
  else if constexpr (sizeof(T) == 4) return (T)__builtin_bswap32((uint32_t)v);
  else return (T)__builtin_bswap64((uint64_t)v);
}
}
