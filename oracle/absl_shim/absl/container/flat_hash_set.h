#pragma once
#include <unordered_set>
#include "absl/hash/hash.h"
namespace absl { template <class K, class H = absl::Hash<K>, class E = std::equal_to<K>> using flat_hash_set = std::unordered_set<K, H, E>;
template <class K, class H = absl::Hash<K>, class E = std::equal_to<K>> using node_hash_set = std::unordered_set<K, H, E>; }
