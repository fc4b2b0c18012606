#pragma once
#include <string_view>
namespace absl { using std::string_view; }
