#pragma once
#include "absl/log/absl_check.h"
#include "absl/base/log_severity.h"
namespace absl_shim {
struct LogStream {
  bool fatal; std::ostringstream os;
  explicit LogStream(bool f) : fatal(f) {}
  ~LogStream() { std::cerr << os.str() << std::endl; if (fatal) std::abort(); }
  template <class T> LogStream& operator<<(const T& v) { os << v; return *this; }
  LogStream& operator<<(std::ostream& (*m)(std::ostream&)) { os << m; return *this; }
};
struct VoidifyL { void operator&(const LogStream&) {} void operator&(const NullStream&) {} };
constexpr bool kSev_INFO = false, kSev_WARNING = false, kSev_ERROR = false, kSev_FATAL = true, kSev_DFATAL = false;
constexpr bool kShow_INFO = false, kShow_WARNING = true, kShow_ERROR = true, kShow_FATAL = true, kShow_DFATAL = true;
}
#define ABSL_LOG(sev) !::absl_shim::kShow_##sev ? (void)0 : ::absl_shim::VoidifyL() & ::absl_shim::LogStream(::absl_shim::kSev_##sev)
#define ABSL_LOG_IF(sev, c) !(::absl_shim::kShow_##sev && (c)) ? (void)0 : ::absl_shim::VoidifyL() & ::absl_shim::LogStream(::absl_shim::kSev_##sev)
#define ABSL_LOG_EVERY_N(sev, n) ABSL_LOG(sev)
#define ABSL_LOG_FIRST_N(sev, n) ABSL_LOG(sev)
#define ABSL_DLOG(sev) true ? (void)0 : ::absl_shim::VoidifyL() & ::absl_shim::NullStream()
#define ABSL_DLOG_IF(sev, c) true ? (void)0 : ::absl_shim::VoidifyL() & ::absl_shim::NullStream()
#define ABSL_VLOG(n) true ? (void)0 : ::absl_shim::VoidifyL() & ::absl_shim::NullStream()
#define ABSL_DVLOG(n) true ? (void)0 : ::absl_shim::VoidifyL() & ::absl_shim::NullStream()
#define ABSL_VLOG_IS_ON(n) false
