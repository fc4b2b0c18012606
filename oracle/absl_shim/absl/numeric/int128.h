#pragma once
#include <cstdint>
#include <ostream>
namespace absl {
using uint128 = unsigned __int128;
using int128 = __int128;
constexpr uint64_t Uint128Low64(uint128 v) { return (uint64_t)v; }
constexpr uint64_t Uint128High64(uint128 v) { return (uint64_t)(v >> 64); }
constexpr uint128 MakeUint128(uint64_t hi, uint64_t lo) { return ((uint128)hi << 64) | lo; }
}
