// Stand-in for absl/base/config.h (test infrastructure, see oracle/build_ref.sh): only the feature macros the reference reads.
#pragma once
#define ABSL_HAVE_BUILTIN(x) 0
