// Stand-in for abseil (absent from this image): attribute macros only. Test infrastructure.
#pragma once
#define ABSL_ATTRIBUTE_ALWAYS_INLINE __attribute__((always_inline))
#define ABSL_ATTRIBUTE_NOINLINE __attribute__((noinline))
#define ABSL_ATTRIBUTE_LIFETIME_BOUND
#define ABSL_ATTRIBUTE_WEAK __attribute__((weak))
#define ABSL_ATTRIBUTE_WARN_UNUSED
#define ABSL_ATTRIBUTE_UNUSED __attribute__((unused))
#define ABSL_ATTRIBUTE_NORETURN __attribute__((noreturn))
#define ABSL_ATTRIBUTE_PACKED __attribute__((packed))
#define ABSL_MUST_USE_RESULT
#define ABSL_ATTRIBUTE_PURE_FUNCTION
#define ABSL_ATTRIBUTE_CONST_FUNCTION
#define ABSL_ATTRIBUTE_COLD
#define ABSL_ATTRIBUTE_HOT
#define ABSL_ATTRIBUTE_REINITIALIZES
#define ABSL_ATTRIBUTE_VIEW
#define ABSL_ATTRIBUTE_OWNER
#define ABSL_CONST_INIT
#define ABSL_FALLTHROUGH_INTENDED [[fallthrough]]
#define ABSL_HAVE_ATTRIBUTE(x) 0
#define ABSL_INTERNAL_ATTRIBUTE_VIEW
#define ABSL_INTERNAL_ATTRIBUTE_OWNER
