#pragma once
#include <cstdio>
#include <ostream>
#include <string>
#include <string_view>
#include <type_traits>
namespace absl {
namespace shim_fmt {
template <class T> auto Arg(const T& v) {
  if constexpr (std::is_same_v<T, std::string>) return v.c_str();
  else if constexpr (std::is_same_v<T, std::string_view>) return std::string(v);
  else return v;
}
template <class T> auto CArg(const T& v) { if constexpr (std::is_same_v<T, std::string>) return v.c_str(); else return v; }
}
template <class... A> struct FormatSpec { const char* f; template <class S> FormatSpec(const S& s) : f(s) {} FormatSpec(const char* s) : f(s) {} };
template <class... A> std::string StrFormat(const char* f, const A&... a) {
  // The reference uses absl's type-safe "%d" for 64-bit values; widen integers to long long and rewrite the spec.
  std::string fmt; for (const char* p = f; *p; ++p) { if (*p == '%' && p[1] != '%') { fmt += '%'; ++p; while (*p && !std::isalpha((unsigned char)*p)) fmt += *p++;
      if (*p == 'd' || *p == 'i') fmt += "lld"; else if (*p == 'u') fmt += "llu"; else if (*p == 'x') fmt += "llx"; else fmt += *p; } else { fmt += *p; if (*p == '%') { fmt += *++p; } } }
  auto widen = [](const auto& v) { using T = std::decay_t<decltype(v)>;
    if constexpr (std::is_integral_v<T> && std::is_signed_v<T>) return (long long)v;
    else if constexpr (std::is_integral_v<T>) return (unsigned long long)v;
    else if constexpr (std::is_same_v<T, std::string>) return v.c_str();
    else return v; };
  int n = std::snprintf(nullptr, 0, fmt.c_str(), widen(a)...);
  std::string out(n > 0 ? n : 0, '\0'); if (n > 0) std::snprintf(out.data(), n + 1, fmt.c_str(), widen(a)...);
  return out;
}
template <class... S, class... A> std::string StrFormat(const FormatSpec<S...>& f, const A&... a) { return StrFormat(f.f, a...); }
template <class... A> void StrAppendFormat(std::string* d, const char* f, const A&... a) { d->append(StrFormat(f, a...)); }
struct StreamFormatted { std::string s; };
inline std::ostream& operator<<(std::ostream& os, const StreamFormatted& f) { return os << f.s; }
template <class... A> StreamFormatted StreamFormat(const char* f, const A&... a) { return {StrFormat(f, a...)}; }
template <class Sink, class... A> bool Format(Sink* sink, const char* f, const A&... a) { sink->Append(StrFormat(f, a...)); return true; }
}
