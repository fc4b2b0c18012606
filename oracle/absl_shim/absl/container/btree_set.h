#pragma once
#include <set>
namespace absl { template <class K, class C = std::less<K>, class A = std::allocator<K>> using btree_set = std::set<K, C, A>;
template <class K, class C = std::less<K>, class A = std::allocator<K>> using btree_multiset = std::multiset<K, C, A>; }
