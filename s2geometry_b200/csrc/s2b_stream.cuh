// s2b_stream.cuh — device helpers shared by the streaming pre-filter kernels (s2b_index.cu phase A0, s2b_cellindex.cu):
// the single-precision leaf-coordinate estimate and its error radius, warp-aggregated list appends in reserved chunks,
// and the mbarrier + 1-D bulk copy (TMA) primitives that stage the point stream through shared memory.
#pragma once
#include <stdint.h>

namespace s2b {

// Work-list slots are reserved per warp in chunks of kSlotChunk with ONE global atomicAdd per chunk (a same-address
// atomic per warp iteration serialises at the L2: ~600k of them per 20M points when most warps see a hit). Unused slots
// of a chunk are filled with a sentinel that the consumer skips.
constexpr uint32_t kSlotChunk = 128;

constexpr int kFastRadius = 2048;

struct U32Slots {   // like WarpSlots, for a list of point indices (sentinel 0xFFFFFFFF)
  uint32_t next = 0, end = 0;
  __device__ __forceinline__ void append(bool want, uint32_t value, uint32_t* list, unsigned int* counter, unsigned lane) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const uint32_t k = (uint32_t)__popc(m);
    if (next + k > end) {
      for (uint32_t s = next + lane; s < end; s += 32) list[s] = 0xFFFFFFFFu;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(counter, kSlotChunk);
      next = __shfl_sync(0xffffffffu, base, 0);
      end = next + kSlotChunk;
    }
    if (want) list[next + (uint32_t)__popc(m & ((1u << lane) - 1u))] = value;
    next += k;
  }
  __device__ __forceinline__ void finish(uint32_t* list, unsigned lane) {
    for (uint32_t s = next + lane; s < end; s += 32) list[s] = 0xFFFFFFFFu;
  }
};

// The same for 64-bit work items (sentinel: all ones).
struct U64Slots {
  uint32_t next = 0, end = 0;
  __device__ __forceinline__ void append(bool want, unsigned long long value, unsigned long long* list, unsigned int* counter, unsigned lane) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const uint32_t k = (uint32_t)__popc(m);
    if (next + k > end) {
      for (uint32_t s = next + lane; s < end; s += 32) list[s] = ~0ull;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(counter, kSlotChunk);
      next = __shfl_sync(0xffffffffu, base, 0);
      end = next + kSlotChunk;
    }
    if (want) list[next + (uint32_t)__popc(m & ((1u << lane) - 1u))] = value;
    next += k;
  }
  __device__ __forceinline__ void finish(unsigned long long* list, unsigned lane) {
    for (uint32_t s = next + lane; s < end; s += 32) list[s] = ~0ull;
  }
};

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ int approx_leaf_coord(float w) {   // w = u or v  ->  approximate leaf index (may be off by < 800)
  const float r = 0.5f * sqrt_approx(fmaf(3.0f, fabsf(w), 1.0f));
  const float s = w >= 0.0f ? r : 1.0f - r;
  return __float2int_rz(1073741824.0f * s);                     // saturating; NaN -> 0
}

// mbarrier + 1-D bulk copy (TMA) helpers: one thread moves a whole tile with a single instruction.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
               "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
  }
}


}  // namespace s2b
