// Stand-in for absl::FixedArray (test infrastructure, see oracle/build_ref.sh): a fixed-size array with a run-time length.
#pragma once
#include <cstddef>
#include <vector>
namespace absl {
template <class T, size_t N = 0>
class FixedArray {
 public:
  using value_type = T;
  using iterator = typename std::vector<T>::iterator;
  using const_iterator = typename std::vector<T>::const_iterator;
  explicit FixedArray(size_t n) : v_(n) {}
  FixedArray(size_t n, const T& value) : v_(n, value) {}
  template <class It> FixedArray(It first, It last) : v_(first, last) {}
  size_t size() const { return v_.size(); }
  bool empty() const { return v_.empty(); }
  T* data() { return v_.data(); }
  const T* data() const { return v_.data(); }
  T& operator[](size_t i) { return v_[i]; }
  const T& operator[](size_t i) const { return v_[i]; }
  T& back() { return v_.back(); }
  const T& back() const { return v_.back(); }
  T& front() { return v_.front(); }
  iterator begin() { return v_.begin(); }
  iterator end() { return v_.end(); }
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  void fill(const T& value) { for (auto& x : v_) x = value; }
 private:
  std::vector<T> v_;
};
}  // namespace absl
