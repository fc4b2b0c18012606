// s2b_region.cu — batched S2ShapeIndexRegion cell tests: the calls an S2RegionCoverer or a
// cell-vs-polygon classification join makes on an index region, one target cell per thread.
//   S2ShapeIndexRegion::MayIntersect(S2Cell)        s2shape_index_region.h:344-371
//   S2ShapeIndexRegion::Contains(S2Cell)            s2shape_index_region.h:313-342
//   S2ShapeIndexRegion::VisitIntersectingShapeIds   s2shape_index_region.h:373-424
// Per-target arithmetic is in s2b_region.cuh. The visitor becomes a CSR of (shape id, contains_target) pairs: raw
// (target, shape, not_contains) tuples are emitted, radix-sorted and reduced per (target, shape), which is what the
// reference's hash map does for a target that spans several index cells.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_hilbert.h"
#include "s2b_index_handle.h"
#include "s2b_internal.h"
#include "s2b_region.cuh"
#include "s2b_stream.cuh"

using namespace s2b;

namespace s2b {
int set_error(int code, const char* fmt, ...);

// This translation unit's copy of the exact-stage counter (s2b_predicates.cuh), read by s2b_exact_stage_count.
cudaError_t region_exact_stage_count(unsigned long long* count, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(count, g_exact_count, sizeof *count);
  if (e == cudaSuccess && reset) {
    const unsigned long long zero = 0;
    e = cudaMemcpyToSymbol(g_exact_count, &zero, sizeof zero);
  }
  return e;
}
int cuda_fail(cudaError_t e, const char* what);
}


namespace {

constexpr int kThreads = 128;   // long per-thread edge loops with ~100 live registers: small CTAs spread them over the SMs
__device__ const HilbertTables g_hilbert_rg = kHilbert;

__device__ __forceinline__ void stage_ij(uint16_t* s_ij) {
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_ij[k] = g_hilbert_rg.ij[k];
  __syncthreads();
}

// kMode 0: MayIntersect(S2Cell), 1: Contains(S2Cell). Two launches per batch, like K3's Locate / process split:
// `region_classify_kernel` answers every target that needs no edge test — DISJOINT, SUBDIVIDED, a target that IS an index
// cell, and a target inside a pure interior cell (no edges, every clipped shape contains the centre: AnyEdgeIntersects is
// false and ShapeContains returns contains_center) — and appends the others to a work list; `region_process_kernel` runs
// the edge tests with one work item per thread, so warps are not left with a quarter of their lanes in the edge loops.
// Work item: target position (32 bits) | index cell (32 bits).
template <int kMode>
__global__ void __launch_bounds__(256)
region_classify_kernel(const FlatIndexView ix, const uint64_t* __restrict__ targets, uint32_t n, uint8_t* __restrict__ out,
                       uint64_t* __restrict__ work, unsigned int* __restrict__ work_count) {
  const uint32_t stride = gridDim.x * 256u;
  const unsigned lane = threadIdx.x & 31u;
  U64Slots slots;
  for (uint64_t base = (uint64_t)blockIdx.x * 256u + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint64_t t = base + lane;
    bool hard = false;
    int32_t at = -1;
    if (t < n) {
      const uint64_t target = __ldcs(targets + t);
      const int rel = locate_target(ix, target, &at);
      uint8_t r = 0;
      if (rel == 1) r = kMode == 0;
      else if (rel == 0) {
        const bool same = ix.cell_ids[at] == target;
        if (same && kMode == 0) r = 1;
        else if (ix.interior_info[at] != 0) r = 1;      // pure interior cell: contained and intersecting, same or not
        else if (same) r = 0;                            // Contains(S2Cell): an index cell with edges (or a clip missing the centre) ...
        else hard = true;
        if (same && kMode == 1 && ix.interior_info[at] == 0) {
          // ... unless one of its clips has no edges and contains the centre (s2shape_index_region.h:330-331)
          const FlatCell fc = ix.cells[at];
          for (uint32_t k = 0; k < fc.clip_count; ++k) {
            const FlatClip clip = ix.clips[fc.clip_begin + k];
            if (clip.num_edges == 0 && (clip.flags & 1u)) r = 1;
          }
        }
      }
      if (!hard) out[t] = r;
    }
    // slots are reserved per warp in chunks (one same-address atomic per warp iteration serialises at the L2: 625k of them
    // per 20M targets were most of this kernel's time); unused slots of a chunk hold the sentinel ~0
    slots.append(hard, (unsigned long long)((t << 32) | (uint32_t)at), (unsigned long long*)work, work_count, lane);
  }
  slots.finish((unsigned long long*)work, lane);
}

template <int kMode>
__global__ void __launch_bounds__(kThreads)
region_process_kernel(const FlatIndexView ix, const uint64_t* __restrict__ targets, const uint64_t* __restrict__ work,
                      const unsigned int* __restrict__ work_count, uint8_t* __restrict__ out) {
  __shared__ uint16_t s_ij[1024];
  stage_ij(s_ij);
  const unsigned count = *work_count;
  const unsigned stride = gridDim.x * kThreads;
  for (unsigned w = blockIdx.x * kThreads + threadIdx.x; w < count; w += stride) {
    const uint64_t item = work[w];
    if (item == ~0ull) continue;                                 // unused slot of a reservation chunk
    const uint64_t t = item >> 32;
    const uint32_t at = (uint32_t)item;
    const uint64_t target = targets[t];
    const FlatCell fc = ix.cells[at];
    const TargetCell tc = target_cell(target, s_ij);
    const P3 a{fc.cx, fc.cy, fc.cz};
    const P3 centre = cellid_to_point(target, s_ij);
    const P3 a_cross_b = cross(a, centre);
    bool r = false;
    for (uint32_t k = 0; k < fc.clip_count && !r; ++k) {
      const FlatClip clip = ix.clips[fc.clip_begin + k];
      if (kMode == 0) {
        r = any_edge_intersects(ix, clip, tc) || clip_contains_centre(ix, clip, a, centre, a_cross_b);
      } else {
        r = ((clip.flags >> 1) & 3u) == 2 && !any_edge_intersects(ix, clip, tc) && clip_contains_centre(ix, clip, a, centre, a_cross_b);
      }
    }
    out[t] = r ? 1 : 0;
  }
}

// Index cells [at, end) lie inside a SUBDIVIDED target: end = first cell past target.range_max().
__device__ __forceinline__ uint32_t subdivided_end(const FlatIndexView& ix, uint64_t target, uint32_t at) {
  const uint64_t tmax = cellid_range_max(target);
  uint32_t lo = at, hi = ix.num_cells;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ix.cell_ids[mid] <= tmax) lo = mid + 1; else hi = mid; }
  return lo;
}

// Tuple key: target index (31 bits) | shape id (32 bits) | not_contains (1 bit). After sorting, the last tuple of every
// (target, shape) run carries the OR of the run's not_contains bits.
__device__ __forceinline__ uint64_t tuple_key(size_t t, int32_t shape, bool not_contains) {
  return ((uint64_t)t << 33) | ((uint64_t)(uint32_t)shape << 1) | (not_contains ? 1u : 0u);
}

// kFill = false: raw tuple count per target; true: tuples written at raw_offsets[t].
template <bool kFill>
__global__ void __launch_bounds__(kThreads)
region_shapes_kernel(const FlatIndexView ix, const uint64_t* __restrict__ targets, size_t n, uint64_t* __restrict__ counts,
                     const uint64_t* __restrict__ raw_offsets, uint64_t* __restrict__ tuples) {
  __shared__ uint16_t s_ij[1024];
  stage_ij(s_ij);
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t target = __ldcs(targets + t);
    int32_t at;
    const int rel = locate_target(ix, target, &at);
    uint32_t cnt = 0;
    uint64_t* dst = kFill ? tuples + raw_offsets[t] : nullptr;
    if (rel == 1) {
      const uint32_t end = subdivided_end(ix, target, (uint32_t)at);
      const uint32_t k0 = ix.cells[at].clip_begin, k1 = ix.cells[end - 1].clip_begin + ix.cells[end - 1].clip_count;
      cnt = k1 - k0;
      if (kFill) {
        for (uint32_t k = k0; k < k1; ++k) {
          const FlatClip clip = ix.clips[k];
          dst[k - k0] = tuple_key(t, clip.shape_id, clip.num_edges > 0 || !(clip.flags & 1u));
        }
      }
    } else if (rel == 0) {
      const FlatCell fc = ix.cells[at];
      const bool same = ix.cell_ids[at] == target;
      TargetCell tc{};
      const P3 a{fc.cx, fc.cy, fc.cz};
      P3 centre{0, 0, 0}, a_cross_b{0, 0, 0};
      if (!same) {
        tc = target_cell(target, s_ij);
        centre = cellid_to_point(target, s_ij);
        a_cross_b = cross(a, centre);
      }
      for (uint32_t k = 0; k < fc.clip_count; ++k) {
        const FlatClip clip = ix.clips[fc.clip_begin + k];
        const int r = region_clip_relation(ix, clip, same, tc, a, centre, a_cross_b);
        if (r == 0) continue;
        if (kFill) dst[cnt] = tuple_key(t, clip.shape_id, r != 2);
        ++cnt;
      }
    }
    if (!kFill) counts[t] = cnt;
  }
}

__global__ void run_tail_flags_kernel(const uint64_t* __restrict__ keys, size_t m, uint8_t* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) flags[i] = (i + 1 == m || (keys[i] >> 1) != (keys[i + 1] >> 1)) ? 1 : 0;
}

// offsets[t] = first compacted pair of target t (lower bound of t << 33); pairs unpacked in the same launch.
__global__ void region_pairs_kernel(const uint64_t* __restrict__ keys, size_t m, size_t n, uint32_t* __restrict__ offsets,
                                    int32_t* __restrict__ shape_ids, uint8_t* __restrict__ contains) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    const uint64_t k = keys[i];
    shape_ids[i] = (int32_t)(uint32_t)(k >> 1);
    contains[i] = (k & 1) ? 0 : 1;
  }
  if (i <= n) {
    const uint64_t want = (uint64_t)i << 33;
    size_t lo = 0, hi = m;
    while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
    offsets[i] = (uint32_t)lo;
  }
}

// One step of the fixed-level covering: kept cells (MayIntersect) at the target level go to `done`, shallower ones hand
// their four children (S2CellId::child, s2cell_id.h:670-676) to the next frontier. One atomicAdd per warp and list.
__global__ void __launch_bounds__(256)
covering_expand_kernel(const uint64_t* __restrict__ ids, const uint8_t* __restrict__ keep, uint32_t n, int level,
                       uint64_t* __restrict__ next, uint64_t* __restrict__ done, unsigned int* __restrict__ counters) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * 256u;
  for (uint64_t base = (uint64_t)blockIdx.x * 256u + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint64_t t = base + lane;
    uint64_t id = 0;
    bool to_next = false, to_done = false;
    if (t < n && keep[t]) {
      id = ids[t];
      to_done = cellid_level(id) >= level;
      to_next = !to_done;
    }
    const unsigned mn = __ballot_sync(0xffffffffu, to_next), md = __ballot_sync(0xffffffffu, to_done);
    unsigned bn = 0, bd = 0;
    if (lane == 0) {
      if (mn) bn = atomicAdd(counters, (unsigned)__popc(mn));
      if (md) bd = atomicAdd(counters + 1, (unsigned)__popc(md));
    }
    bn = __shfl_sync(0xffffffffu, bn, 0);
    bd = __shfl_sync(0xffffffffu, bd, 0);
    const unsigned below = (1u << lane) - 1u;
    if (to_done) done[bd + __popc(md & below)] = id;
    if (to_next) {
      const uint64_t child_lsb = cellid_lsb(id) >> 2;
      uint64_t* dst = next + 4ull * (bn + __popc(mn & below));
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = id + (uint64_t)(int64_t)(2 * k + 1 - 4) * child_lsb;
    }
  }
}

int grid_for(size_t n, const void* kernel) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  const size_t want = (n + kThreads - 1) / kThreads, cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

// Stream-ordered scratch stays cached across synchronisations (the default release threshold of 0 hands freed blocks back
// to the driver at every sync, which costs milliseconds per call); same setting as the point queries make.
void keep_pool_memory_once() {
  static thread_local int done_for_device = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for_device) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_for_device = dev;
}

int check_region_index(const S2bIndex* ix) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the index lives on device %d but the current device is %d", ix->device, dev);
  return S2B_OK;
}

// Targets per classify/process round: bounds the work list (8 B per target in the worst case) at 1 GB and keeps target
// positions in 32 bits.
constexpr size_t kRegionRound = size_t(1) << 27;
constexpr unsigned int kRegionSortMin = 1u << 16;   // below this the sort costs more than the divergence it removes

template <int kMode>
int region_cells_dev(const S2bIndex* ix, const uint64_t* ids_dev, size_t n, uint8_t* out_dev, void* stream) {
  int rc = check_region_index(ix);
  if (rc != S2B_OK) return rc;
  if (n == 0) return S2B_OK;
  if (!ids_dev || !out_dev) return set_error(S2B_EINVAL, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  keep_pool_memory_once();
  const size_t round = n < kRegionRound ? n : kRegionRound;
  void* scratch = nullptr;
  // slots: a chunk of kSlotChunk holds at least kSlotChunk - 31 items, every warp leaves at most one partly used chunk
  int per_sm_c = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, (const void*)region_classify_kernel<kMode>, 256, 0);
  const size_t warps = (size_t)sm_count() * (per_sm_c < 1 ? 1 : per_sm_c) * 8;
  const size_t slots_cap = kSlotChunk * (round / (kSlotChunk - 31) + 2 + warps);
  cudaError_t e = cudaMallocAsync(&scratch, 8 * slots_cap + 256, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(region work list)");
  uint64_t* work = (uint64_t*)((char*)scratch + 256);
  unsigned int* work_count = (unsigned int*)scratch;
  auto kc = region_classify_kernel<kMode>;
  auto kp = region_process_kernel<kMode>;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kc, 256, 0);
  if (per_sm < 1) per_sm = 1;
  // The work list is sorted by index cell before the edge tests (radix sort on the low log2(num_cells) bits of the items),
  // so the lanes of a warp read the same cell's clips and edges and run the same trip counts. Needs the item count on the
  // host: one 4-byte copy + stream synchronisation per round.
  int cell_bits = 1;
  while ((1ull << cell_bits) < ix->view.num_cells) ++cell_bits;
  void* sort_tmp = nullptr;
  uint64_t* sorted = nullptr;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, work, work, slots_cap, 0, cell_bits, st);
  for (size_t off = 0; off < n && e == cudaSuccess; off += round) {
    const size_t cnt = n - off < round ? n - off : round;
    e = cudaMemsetAsync(work_count, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) break;
    const size_t want = (cnt + 255) / 256, cap = (size_t)sm_count() * per_sm;
    kc<<<(int)(want < cap ? want : cap), 256, 0, st>>>(ix->view, ids_dev + off, (uint32_t)cnt, out_dev + off, work, work_count);
    count_launch();
    unsigned int hard = 0;
    e = cudaMemcpyAsync(&hard, work_count, sizeof hard, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess || hard == 0) continue;
    const uint64_t* items = work;
    if (hard >= kRegionSortMin) {
      if (!sorted) {
        e = cudaMallocAsync(&sort_tmp, sort_bytes + 8 * slots_cap + 256, st);
        if (e != cudaSuccess) break;
        sorted = (uint64_t*)((char*)sort_tmp + ((sort_bytes + 255) / 256) * 256);
      }
      e = cub::DeviceRadixSort::SortKeys(sort_tmp, sort_bytes, work, sorted, (size_t)hard, 0, cell_bits, st);
      if (e != cudaSuccess) break;
      count_launch();
      items = sorted;
    }
    kp<<<grid_for(hard, (const void*)kp), kThreads, 0, st>>>(ix->view, ids_dev + off, items, work_count, out_dev + off);
    count_launch();
    e = cudaGetLastError();
  }
  if (sort_tmp) cudaFreeAsync(sort_tmp, st);
  cudaFreeAsync(scratch, st);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "region cells launch");
}

template <int kMode>
int region_cells_host(const S2bIndex* ix, const uint64_t* ids, size_t n, uint8_t* out) {
  int rc = check_region_index(ix);
  if (rc != S2B_OK) return rc;
  if (n == 0) return S2B_OK;
  if (!ids || !out) return set_error(S2B_EINVAL, "null pointer");
  void *d_ids = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_ids, 8 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, n);
  if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids, 8 * n, cudaMemcpyHostToDevice);
  rc = e == cudaSuccess ? region_cells_dev<kMode>(ix, (const uint64_t*)d_ids, n, (uint8_t*)d_out, nullptr) : cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK && (e = cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  cudaFree(d_ids); cudaFree(d_out);
  return rc;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T> T* as() const { return (T*)p; }
};

}  // namespace

extern "C" {

int s2b_region_may_intersect_cells_dev(const S2bIndex* ix, const uint64_t* cell_ids_dev, size_t n, uint8_t* out_dev, void* stream) {
  return region_cells_dev<0>(ix, cell_ids_dev, n, out_dev, stream);
}
int s2b_region_may_intersect_cells(const S2bIndex* ix, const uint64_t* cell_ids, size_t n, uint8_t* out) {
  return region_cells_host<0>(ix, cell_ids, n, out);
}
int s2b_region_contains_cells_dev(const S2bIndex* ix, const uint64_t* cell_ids_dev, size_t n, uint8_t* out_dev, void* stream) {
  return region_cells_dev<1>(ix, cell_ids_dev, n, out_dev, stream);
}
int s2b_region_contains_cells(const S2bIndex* ix, const uint64_t* cell_ids, size_t n, uint8_t* out) {
  return region_cells_host<1>(ix, cell_ids, n, out);
}

int s2b_region_intersecting_shapes(const S2bIndex* ix, const uint64_t* cell_ids, size_t n, uint32_t* offsets_out, int32_t* shape_ids_out,
                                   uint8_t* contains_out, size_t capacity, size_t* total_out) {
  int rc = check_region_index(ix);
  if (rc != S2B_OK) return rc;
  if (!offsets_out) return set_error(S2B_EINVAL, "null offsets_out");
  if (total_out) *total_out = 0;
  if (n == 0) { offsets_out[0] = 0; return S2B_OK; }
  if (!cell_ids) return set_error(S2B_EINVAL, "null pointer");
  if (n >= (size_t(1) << 31)) return set_error(S2B_EINVAL, "at most 2^31 - 1 targets per call");
  DevBuf ids, counts, raw_off, temp, tuples, sorted, flags, compact, n_sel, off, shp, con;
  cudaError_t e = ids.alloc(8 * n);
  if (e == cudaSuccess) e = counts.alloc(8 * (n + 1));
  if (e == cudaSuccess) e = raw_off.alloc(8 * (n + 1));
  if (e == cudaSuccess) e = cudaMemcpy(ids.p, cell_ids, 8 * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(counts.p, 0, 8 * (n + 1));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc/cudaMemcpy");
  auto kc = region_shapes_kernel<false>;
  auto kf = region_shapes_kernel<true>;
  kc<<<grid_for(n, (const void*)kc), kThreads>>>(ix->view, ids.as<uint64_t>(), n, counts.as<uint64_t>(), nullptr, nullptr);
  count_launch();
  size_t temp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts.as<uint64_t>(), raw_off.as<uint64_t>(), n + 1);
  if ((e = temp.alloc(temp_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cub::DeviceScan::ExclusiveSum(temp.p, temp_bytes, counts.as<uint64_t>(), raw_off.as<uint64_t>(), n + 1);
  count_launch();
  uint64_t raw_total = 0;
  if ((e = cudaMemcpy(&raw_total, raw_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "region shapes count");
  if (raw_total >= (uint64_t(1) << 31)) return set_error(S2B_ECAPACITY, "%llu raw (target, shape) tuples: split the batch", (unsigned long long)raw_total);
  const size_t m = (size_t)raw_total;
  size_t total = 0;
  if (m > 0) {
    e = tuples.alloc(8 * m);
    if (e == cudaSuccess) e = sorted.alloc(8 * m);
    if (e == cudaSuccess) e = flags.alloc(m);
    if (e == cudaSuccess) e = compact.alloc(8 * m);
    if (e == cudaSuccess) e = n_sel.alloc(8);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    kf<<<grid_for(n, (const void*)kf), kThreads>>>(ix->view, ids.as<uint64_t>(), n, nullptr, raw_off.as<uint64_t>(), tuples.as<uint64_t>());
    count_launch();
    DevBuf temp2;
    size_t sort_bytes = 0, sel_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, tuples.as<uint64_t>(), sorted.as<uint64_t>(), m);
    cub::DeviceSelect::Flagged(nullptr, sel_bytes, sorted.as<uint64_t>(), flags.as<uint8_t>(), compact.as<uint64_t>(), n_sel.as<uint64_t>(), m);
    if ((e = temp2.alloc(sort_bytes > sel_bytes ? sort_bytes : sel_bytes)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cub::DeviceRadixSort::SortKeys(temp2.p, sort_bytes, tuples.as<uint64_t>(), sorted.as<uint64_t>(), m);
    run_tail_flags_kernel<<<(unsigned)((m + 255) / 256), 256>>>(sorted.as<uint64_t>(), m, flags.as<uint8_t>());
    cub::DeviceSelect::Flagged(temp2.p, sel_bytes, sorted.as<uint64_t>(), flags.as<uint8_t>(), compact.as<uint64_t>(), n_sel.as<uint64_t>(), m);
    count_launch(); count_launch(); count_launch();
    uint64_t sel = 0;
    if ((e = cudaMemcpy(&sel, n_sel.p, 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "region shapes reduce");
    total = (size_t)sel;
  }
  if (total_out) *total_out = total;
  e = off.alloc(4 * (n + 1));
  if (e == cudaSuccess) e = shp.alloc(4 * total);
  if (e == cudaSuccess) e = con.alloc(total);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  const size_t work = (total > n + 1 ? total : n + 1);
  region_pairs_kernel<<<(unsigned)((work + 255) / 256), 256>>>(compact.as<uint64_t>(), total, n, off.as<uint32_t>(), shp.as<int32_t>(), con.as<uint8_t>());
  count_launch();
  if ((e = cudaMemcpy(offsets_out, off.p, 4 * (n + 1), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "region shapes offsets");
  if (total > capacity) return set_error(S2B_ECAPACITY, "%zu (cell, shape) pairs but capacity is %zu", total, capacity);
  if (total > 0) {
    if (!shape_ids_out || !contains_out) return set_error(S2B_EINVAL, "null output arrays");
    if ((e = cudaMemcpy(shape_ids_out, shp.p, 4 * total, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    if ((e = cudaMemcpy(contains_out, con.p, total, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  }
  return S2B_OK;
}

int s2b_region_cell_union_bound(const S2bIndex* ix, uint64_t cell_ids_out[6], int* count_out) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (!cell_ids_out || !count_out) return set_error(S2B_EINVAL, "null pointer");
  *count_out = 0;
  const size_t C = ix->view.num_cells;
  if (C == 0) return S2B_OK;
  std::vector<uint64_t> ids(C);
  cudaError_t e = cudaMemcpy(ids.data(), ix->view.cell_ids, 8 * C, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(cell ids)");
  *count_out = region_cell_union_bound(ids.data(), C, cell_ids_out);
  return S2B_OK;
}

int s2b_region_covering_at_level(const S2bIndex* ix, int level, uint64_t* cell_ids_out, size_t capacity, size_t* total_out) {
  int rc = check_region_index(ix);
  if (rc != S2B_OK) return rc;
  if (level < 0 || level > 30) return set_error(S2B_EINVAL, "level must be in [0, 30]");
  if (total_out) *total_out = 0;
  uint64_t start[6];
  int num_start = 0;
  if ((rc = s2b_region_cell_union_bound(ix, start, &num_start)) != S2B_OK) return rc;
  std::vector<uint64_t> frontier;
  for (int k = 0; k < num_start; ++k) frontier.push_back(cellid_level(start[k]) > level ? cellid_parent(start[k], level) : start[k]);
  std::sort(frontier.begin(), frontier.end());
  frontier.erase(std::unique(frontier.begin(), frontier.end()), frontier.end());
  std::vector<uint64_t> result;
  DevBuf cur, counters;
  cudaError_t e = counters.alloc(8);
  size_t n = frontier.size();
  if (e == cudaSuccess) e = cur.alloc(8 * n);
  if (e == cudaSuccess && n) e = cudaMemcpy(cur.p, frontier.data(), 8 * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc/cudaMemcpy");
  while (n > 0) {
    if (n >= (size_t(1) << 29)) return set_error(S2B_ECAPACITY, "covering frontier of %zu cells: choose a coarser level", n);
    DevBuf next_buf, done_buf, keep_buf;
    e = next_buf.alloc(32 * n);
    if (e == cudaSuccess) e = done_buf.alloc(8 * n);
    if (e == cudaSuccess) e = keep_buf.alloc(n);
    if (e == cudaSuccess) e = cudaMemset(counters.p, 0, 8);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(covering frontier)");
    if ((rc = region_cells_dev<0>(ix, cur.as<uint64_t>(), n, keep_buf.as<uint8_t>(), nullptr)) != S2B_OK) return rc;
    const size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    covering_expand_kernel<<<(int)(want < cap ? want : cap), 256>>>(cur.as<uint64_t>(), keep_buf.as<uint8_t>(), (uint32_t)n, level,
                                                                   next_buf.as<uint64_t>(), done_buf.as<uint64_t>(), counters.as<unsigned int>());
    count_launch();
    unsigned int cnt[2] = {0, 0};
    if ((e = cudaMemcpy(cnt, counters.p, 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "covering step");
    if (cnt[1]) {
      const size_t old = result.size();
      result.resize(old + cnt[1]);
      if ((e = cudaMemcpy(result.data() + old, done_buf.p, 8ull * cnt[1], cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    }
    n = 4ull * cnt[0];
    std::swap(cur.p, next_buf.p);   // next_buf's destructor frees the old frontier
  }
  std::sort(result.begin(), result.end());
  if (total_out) *total_out = result.size();
  if (result.size() > capacity) return set_error(S2B_ECAPACITY, "%zu covering cells but capacity is %zu", result.size(), capacity);
  if (!result.empty()) {
    if (!cell_ids_out) return set_error(S2B_EINVAL, "null cell_ids_out");
    std::memcpy(cell_ids_out, result.data(), 8 * result.size());
  }
  return S2B_OK;
}

}  // extern "C"
