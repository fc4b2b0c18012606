#pragma once
#include <string>
namespace absl { template <class T> struct Flag { T value; }; 
template <class T> const T& GetFlag(const Flag<T>& f) { return f.value; }
template <class T, class V> void SetFlag(Flag<T>* f, const V& v) { f->value = T(v); } }
#define ABSL_FLAG(type, name, def, help) absl::Flag<type> FLAGS_##name{def}
#define ABSL_RETIRED_FLAG(type, name, def, help)
