// Streaming CHECK/DCHECK stand-ins. CHECK failure aborts; DCHECK compiles away (NDEBUG oracle build).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include "absl/base/attributes.h"
#include "absl/base/casts.h"
#include "absl/base/macros.h"
#include "absl/base/optimization.h"
namespace absl_shim {
struct FatalStream {
  std::ostringstream os;
  FatalStream(const char* f, int l, const char* c) { os << f << ":" << l << " CHECK failed: " << c << " "; }
  [[noreturn]] ~FatalStream() { std::cerr << os.str() << std::endl; std::abort(); }
  template <class T> FatalStream& operator<<(const T& v) { os << v; return *this; }
  FatalStream& operator<<(std::ostream& (*m)(std::ostream&)) { os << m; return *this; }
};
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct Voidify { void operator&(const FatalStream&) {} void operator&(const NullStream&) {} };
}
#define ABSL_SHIM_CHECK_(c, txt) (c) ? (void)0 : ::absl_shim::Voidify() & ::absl_shim::FatalStream(__FILE__, __LINE__, txt)
#define ABSL_SHIM_NULL_(c) true ? (void)0 : ::absl_shim::Voidify() & ::absl_shim::NullStream()
#define ABSL_CHECK(c) ABSL_SHIM_CHECK_((c), #c)
#define ABSL_QCHECK(c) ABSL_SHIM_CHECK_((c), #c)
#define ABSL_CHECK_OK(s) ABSL_SHIM_CHECK_((s).ok(), #s)
#define ABSL_CHECK_EQ(a, b) ABSL_SHIM_CHECK_((a) == (b), #a " == " #b)
#define ABSL_CHECK_NE(a, b) ABSL_SHIM_CHECK_((a) != (b), #a " != " #b)
#define ABSL_CHECK_LT(a, b) ABSL_SHIM_CHECK_((a) < (b), #a " < " #b)
#define ABSL_CHECK_LE(a, b) ABSL_SHIM_CHECK_((a) <= (b), #a " <= " #b)
#define ABSL_CHECK_GT(a, b) ABSL_SHIM_CHECK_((a) > (b), #a " > " #b)
#define ABSL_CHECK_GE(a, b) ABSL_SHIM_CHECK_((a) >= (b), #a " >= " #b)
#ifdef NDEBUG
#define ABSL_DCHECK(c) ABSL_SHIM_NULL_(c)
#define ABSL_DCHECK_OK(s) ABSL_SHIM_NULL_(s)
#define ABSL_DCHECK_EQ(a, b) ABSL_SHIM_NULL_((a) == (b))
#define ABSL_DCHECK_NE(a, b) ABSL_SHIM_NULL_((a) != (b))
#define ABSL_DCHECK_LT(a, b) ABSL_SHIM_NULL_((a) < (b))
#define ABSL_DCHECK_LE(a, b) ABSL_SHIM_NULL_((a) <= (b))
#define ABSL_DCHECK_GT(a, b) ABSL_SHIM_NULL_((a) > (b))
#define ABSL_DCHECK_GE(a, b) ABSL_SHIM_NULL_((a) >= (b))
#else
#define ABSL_DCHECK(c) ABSL_CHECK(c)
#define ABSL_DCHECK_OK(s) ABSL_CHECK_OK(s)
#define ABSL_DCHECK_EQ(a, b) ABSL_CHECK_EQ(a, b)
#define ABSL_DCHECK_NE(a, b) ABSL_CHECK_NE(a, b)
#define ABSL_DCHECK_LT(a, b) ABSL_CHECK_LT(a, b)
#define ABSL_DCHECK_LE(a, b) ABSL_CHECK_LE(a, b)
#define ABSL_DCHECK_GT(a, b) ABSL_CHECK_GT(a, b)
#define ABSL_DCHECK_GE(a, b) ABSL_CHECK_GE(a, b)
#endif
