#pragma once
#include <cstdint>
#include <cstdio>
#include <sstream>
#include <string>
#include <string_view>
namespace absl {
enum PadSpec { kNoPad = 1, kZeroPad2, kZeroPad3, kZeroPad4, kZeroPad5, kZeroPad6, kZeroPad7, kZeroPad8, kZeroPad9,
  kZeroPad10, kZeroPad11, kZeroPad12, kZeroPad13, kZeroPad14, kZeroPad15, kZeroPad16 };
struct Hex { uint64_t v; int width; Hex(uint64_t x, PadSpec p = kNoPad) : v(x), width((int)p) {} };
inline std::ostream& operator<<(std::ostream& os, const Hex& h) {
  char buf[32]; std::snprintf(buf, sizeof buf, "%0*llx", h.width, (unsigned long long)h.v); return os << buf;
}
template <class... A> std::string StrCat(const A&... a) { std::ostringstream os; os.precision(17); (void)(os << ... << a); return os.str(); }
template <class... A> void StrAppend(std::string* d, const A&... a) { d->append(StrCat(a...)); }
}
