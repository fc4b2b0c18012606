#pragma once
#include <cctype>
#include <string>
#include <string_view>
namespace absl {
inline bool ascii_isspace(unsigned char c) { return std::isspace(c); }
inline bool ascii_isdigit(unsigned char c) { return std::isdigit(c); }
inline bool ascii_isxdigit(unsigned char c) { return std::isxdigit(c); }
inline bool ascii_isalpha(unsigned char c) { return std::isalpha(c); }
inline char ascii_tolower(unsigned char c) { return (char)std::tolower(c); }
inline char ascii_toupper(unsigned char c) { return (char)std::toupper(c); }
inline std::string_view StripAsciiWhitespace(std::string_view s) {
  while (!s.empty() && std::isspace((unsigned char)s.front())) s.remove_prefix(1);
  while (!s.empty() && std::isspace((unsigned char)s.back())) s.remove_suffix(1);
  return s;
}
}
