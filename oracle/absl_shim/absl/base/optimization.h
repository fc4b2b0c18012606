#pragma once
#define ABSL_PREDICT_TRUE(x) (__builtin_expect(false || (x), true))
#define ABSL_PREDICT_FALSE(x) (__builtin_expect(false || (x), false))
#define ABSL_ASSUME(c) do { if (!(c)) __builtin_unreachable(); } while (0)
#define ABSL_UNREACHABLE() __builtin_unreachable()
#define ABSL_BLOCK_TAIL_CALL_OPTIMIZATION() __asm__ __volatile__("")
#define ABSL_CACHELINE_SIZE 64
#define ABSL_CACHELINE_ALIGNED __attribute__((aligned(64)))
