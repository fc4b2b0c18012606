#pragma once
#include <unordered_map>
#include "absl/hash/hash.h"
namespace absl { template <class K, class V, class H = absl::Hash<K>, class E = std::equal_to<K>> using flat_hash_map = std::unordered_map<K, V, H, E>;
template <class K, class V, class H = absl::Hash<K>, class E = std::equal_to<K>> using node_hash_map = std::unordered_map<K, V, H, E>; }
