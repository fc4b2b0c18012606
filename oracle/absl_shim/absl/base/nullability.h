#pragma once
#define absl_nonnull
#define absl_nullable
#define absl_nullability_unknown
namespace absl { template <typename T> using Nonnull = T; template <typename T> using Nullable = T; }
