#pragma once
#include <cassert>
#include <cstddef>
#include "absl/base/attributes.h"
#include "absl/base/optimization.h"
#define ABSL_DEPRECATED(msg)
#define ABSL_DEPRECATE_AND_INLINE()
#define ABSL_HARDENING_ASSERT(x) assert(x)
#define ABSL_ASSERT(x) assert(x)
#define ABSL_ARRAYSIZE(a) (sizeof(a) / sizeof((a)[0]))
#define ABSL_INTERNAL_TRY try
#define ABSL_INTERNAL_CATCH_ANY catch (...)
#define ABSL_INTERNAL_RETHROW do { throw; } while (false)
