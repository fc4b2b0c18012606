#pragma once
#include <ostream>
#include <string>
#include <string_view>
namespace absl {
enum class StatusCode : int { kOk = 0, kCancelled = 1, kUnknown = 2, kInvalidArgument = 3, kDeadlineExceeded = 4, kNotFound = 5,
  kAlreadyExists = 6, kPermissionDenied = 7, kResourceExhausted = 8, kFailedPrecondition = 9, kAborted = 10, kOutOfRange = 11,
  kUnimplemented = 12, kInternal = 13, kUnavailable = 14, kDataLoss = 15, kUnauthenticated = 16 };
inline std::string StatusCodeToString(StatusCode c) { return "StatusCode(" + std::to_string((int)c) + ")"; }
inline std::ostream& operator<<(std::ostream& os, StatusCode c) { return os << StatusCodeToString(c); }
class Status {
 public:
  Status() = default;
  Status(StatusCode c, std::string_view m) : code_(c), msg_(m) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  std::string_view message() const { return msg_; }
  std::string ToString() const { return ok() ? "OK" : StatusCodeToString(code_) + ": " + msg_; }
  friend bool operator==(const Status& a, const Status& b) { return a.code_ == b.code_ && a.msg_ == b.msg_; }
  friend bool operator!=(const Status& a, const Status& b) { return !(a == b); }
 private:
  StatusCode code_ = StatusCode::kOk; std::string msg_;
};
inline std::ostream& operator<<(std::ostream& os, const Status& s) { return os << s.ToString(); }
inline Status OkStatus() { return Status(); }
#define ABSL_SHIM_ERR_(Name, kc) inline Status Name##Error(std::string_view m) { return Status(StatusCode::kc, m); } \
  inline bool Is##Name(const Status& s) { return s.code() == StatusCode::kc; }
ABSL_SHIM_ERR_(Cancelled, kCancelled) ABSL_SHIM_ERR_(Unknown, kUnknown) ABSL_SHIM_ERR_(InvalidArgument, kInvalidArgument)
ABSL_SHIM_ERR_(DeadlineExceeded, kDeadlineExceeded) ABSL_SHIM_ERR_(NotFound, kNotFound) ABSL_SHIM_ERR_(AlreadyExists, kAlreadyExists)
ABSL_SHIM_ERR_(PermissionDenied, kPermissionDenied) ABSL_SHIM_ERR_(ResourceExhausted, kResourceExhausted)
ABSL_SHIM_ERR_(FailedPrecondition, kFailedPrecondition) ABSL_SHIM_ERR_(Aborted, kAborted) ABSL_SHIM_ERR_(OutOfRange, kOutOfRange)
ABSL_SHIM_ERR_(Unimplemented, kUnimplemented) ABSL_SHIM_ERR_(Internal, kInternal) ABSL_SHIM_ERR_(Unavailable, kUnavailable)
ABSL_SHIM_ERR_(DataLoss, kDataLoss) ABSL_SHIM_ERR_(Unauthenticated, kUnauthenticated)
}
