#pragma once
#include <cstddef>
#include <map>
namespace absl {
// std::map plus the tree_.bytes_used() member the reference's memory accounting pokes at.
template <class K, class V, class C = std::less<K>, class A = std::allocator<std::pair<const K, V>>>
class btree_map : public std::map<K, V, C, A> {
 public:
  using std::map<K, V, C, A>::map;
  struct Tree { const btree_map* m; size_t bytes_used() const { return m->size() * (sizeof(std::pair<const K, V>) + 32); } };
  Tree tree_{this};
  btree_map() : std::map<K, V, C, A>() {}
  btree_map(const btree_map& o) : std::map<K, V, C, A>(o), tree_{this} {}
  btree_map(btree_map&& o) noexcept : std::map<K, V, C, A>(std::move(o)), tree_{this} {}
  btree_map& operator=(const btree_map& o) { std::map<K, V, C, A>::operator=(o); return *this; }
  btree_map& operator=(btree_map&& o) noexcept { std::map<K, V, C, A>::operator=(std::move(o)); return *this; }
};
template <class K, class V, class C = std::less<K>, class A = std::allocator<std::pair<const K, V>>>
using btree_multimap = std::multimap<K, V, C, A>;
}
