#pragma once
#include <stdexcept>
#include <string>
namespace absl {
[[noreturn]] inline void ThrowStdLengthError(const char* m) { throw std::length_error(m); }
[[noreturn]] inline void ThrowStdOutOfRange(const char* m) { throw std::out_of_range(m); }
[[noreturn]] inline void ThrowStdLengthError(const std::string& m) { throw std::length_error(m); }
[[noreturn]] inline void ThrowStdOutOfRange(const std::string& m) { throw std::out_of_range(m); }
namespace base_internal { using absl::ThrowStdLengthError; using absl::ThrowStdOutOfRange; }
}
