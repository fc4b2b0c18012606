#pragma once
#include <cstddef>
#include <memory>
#include <string_view>
#include <vector>
#include "absl/strings/string_view.h"
namespace absl { template <class T, size_t N, class A = std::allocator<T>> using InlinedVector = std::vector<T, A>; }
