#pragma once
#include <type_traits>
#include <utility>
namespace absl {
template <class Sig> class FunctionRef;
template <class R, class... A> class FunctionRef<R(A...)> {
  void* obj_; R (*call_)(void*, A...);
 public:
  template <class F, class = std::enable_if_t<!std::is_same_v<std::decay_t<F>, FunctionRef>>>
  FunctionRef(F&& f) : obj_((void*)&f), call_([](void* o, A... a) -> R { return (*(std::remove_reference_t<F>*)o)(std::forward<A>(a)...); }) {}
  R operator()(A... a) const { return call_(obj_, std::forward<A>(a)...); }
};
}
