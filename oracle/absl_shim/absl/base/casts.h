#pragma once
#include <bit>
#include <type_traits>
namespace absl {
template <typename To> constexpr To implicit_cast(std::type_identity_t<To> to) { return to; }
using std::bit_cast;
}
