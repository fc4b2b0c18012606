#pragma once
#include <type_traits>
namespace absl {
using std::is_trivially_default_constructible; using std::is_trivially_copy_constructible; using std::is_trivially_destructible;
using std::is_trivially_copy_assignable; using std::enable_if_t; using std::conditional_t; using std::decay_t; using std::remove_cv_t;
using std::remove_reference_t; using std::void_t; using std::conjunction; using std::disjunction; using std::negation; using std::underlying_type_t;
using std::result_of_t; using std::make_unsigned_t; using std::make_signed_t; using std::remove_const_t; using std::remove_extent_t; using std::remove_pointer_t; using std::add_const_t;
using std::common_type_t; using std::is_trivially_copyable;
}
