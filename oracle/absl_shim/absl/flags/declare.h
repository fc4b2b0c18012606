#pragma once
#include "absl/flags/flag.h"
#define ABSL_DECLARE_FLAG(type, name) extern absl::Flag<type> FLAGS_##name
