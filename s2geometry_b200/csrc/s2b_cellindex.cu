// s2b_cellindex.cu — batched S2CellIndex joins:
//   S2CellIndex::Add / Build                     s2cell_index.h:143-160, s2cell_index.cc:71-146
//   S2CellIndex::VisitIntersectingCells          s2cell_index.h:619-647
//   S2CellIndex::GetIntersectingLabels           s2cell_index.cc:148-163
//   S2RegionSharder::GetMostIntersectingShard    s2region_sharder.cc:112-140 (regions given as S2CellUnions)
// plus the fused point form: leaf cell of every point (K1 arithmetic) -> per-label hit histogram.
// One thread per target; data layout and traversal in s2b_cellindex.cuh. Integer work, bit-exact with the reference.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <new>

#include "../../include/s2b_capi.h"
#include "s2b_cellindex.cuh"
#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_internal.h"
#include "s2b_sincos.cuh"

struct S2bCellIndex {
  int device = 0;
  void* blob = nullptr;
  size_t blob_bytes = 0;
  int32_t num_labels = 0;
  s2b::CellIndexView view{};
};

namespace s2b {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

namespace {

constexpr int kThreads = 256;
__device__ const HilbertTables g_hilbert_ci = kHilbert;

int grid_for_ci(size_t n, const void* kernel) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  size_t want = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

struct Targets {   // T target unions: cells ids[offsets[t] .. offsets[t+1]), or one cell per target when offsets == nullptr
  const uint64_t* ids;
  const uint32_t* offsets;
  size_t T;
  __device__ __forceinline__ void range(size_t t, uint32_t* b, uint32_t* e) const {
    if (offsets) { *b = offsets[t]; *e = offsets[t + 1]; } else { *b = (uint32_t)t; *e = (uint32_t)t + 1; }
  }
};

// Pass 1: number of pairs every target reports; the grand total is accumulated in 64 bits. all_pairs: report a pair once
// per target CELL it intersects instead of once per target (the sharder scores every (shard cell, bound cell) overlap).
__global__ void __launch_bounds__(kThreads)
ci_count_kernel(CellIndexView ix, Targets tg, bool all_pairs, uint32_t* __restrict__ counts, unsigned long long* __restrict__ total) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  unsigned long long mine = 0;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < tg.T; t += stride) {
    uint32_t b, e;
    tg.range(t, &b, &e);
    CiUnionState st;
    unsigned long long c = 0;
    for (uint32_t k = b; k < e; ++k) {
      if (all_pairs) st.cutoff = -1;
      c += ci_visit_cell(ix, tg.ids[k], st, [](const CiNode&) {});
    }
    counts[t] = (uint32_t)(c > 0xFFFFFFFFull ? 0xFFFFFFFFull : c);
    mine += c;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

// Pass 2: the pairs themselves. weight (nullable) = lsb of (pair cell ∩ the target cell that reported it), or 0 when a
// filter union is given for the target (flt.ids != nullptr) and that intersection cell misses it.
__global__ void __launch_bounds__(kThreads)
ci_fill_kernel(CellIndexView ix, Targets tg, Targets flt, const uint32_t* __restrict__ offsets, uint64_t* __restrict__ cells_out,
               int32_t* __restrict__ labels_out, uint64_t* __restrict__ weight_out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < tg.T; t += stride) {
    uint32_t b, e;
    tg.range(t, &b, &e);
    CiUnionState st;
    size_t w = offsets[t];
    uint32_t fb = 0, fe = 0;
    if (flt.ids) flt.range(t, &fb, &fe);
    for (uint32_t k = b; k < e; ++k) {
      const uint64_t target = tg.ids[k];
      if (weight_out) st.cutoff = -1;   // all_pairs mode of the count pass
      ci_visit_cell(ix, target, st, [&](const CiNode& nd) {
        if (cells_out) cells_out[w] = nd.id;
        if (labels_out) labels_out[w] = nd.label;
        if (weight_out) {
          const uint64_t c = ci_overlap_cell(nd.id, target);
          weight_out[w] = (!flt.ids || ci_union_intersects(flt.ids + fb, fe - fb, c)) ? cellid_lsb(c) : 0;
        }
        ++w;
      });
    }
  }
}

// Sorted label segments -> number of distinct labels per target.
__global__ void __launch_bounds__(kThreads)
ci_unique_count_kernel(const int32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets, size_t T, uint32_t* __restrict__ counts,
                       unsigned long long* __restrict__ total) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  unsigned long long mine = 0;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < T; t += stride) {
    uint32_t c = 0;
    for (uint32_t k = offsets[t]; k < offsets[t + 1]; ++k) c += (k == offsets[t] || sorted[k] != sorted[k - 1]);
    counts[t] = c;
    mine += c;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

__global__ void __launch_bounds__(kThreads)
ci_unique_fill_kernel(const int32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets, size_t T, const uint32_t* __restrict__ out_offsets,
                      int32_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < T; t += stride) {
    size_t w = out_offsets[t];
    for (uint32_t k = offsets[t]; k < offsets[t + 1]; ++k)
      if (k == offsets[t] || sorted[k] != sorted[k - 1]) out[w++] = sorted[k];
  }
}

// GetMostIntersectingShard over (label, overlap) pairs sorted by label.
__global__ void __launch_bounds__(kThreads)
ci_best_label_kernel(const int32_t* __restrict__ sorted, const uint64_t* __restrict__ weight, const uint32_t* __restrict__ offsets, size_t T,
                     int32_t fallback, int32_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < T; t += stride) {
    out[t] = ci_best_label(sorted, weight, offsets[t], offsets[t + 1], fallback);
  }
}

// Fused point form: S2CellId(point) with K1's arithmetic, then the pairs containing that leaf cell. Every lookup is a
// dependent, scattered load (bucket table -> range record -> overflow labels), so the kernel is bound by load latency:
// each thread carries kPts points through the stages together to keep that many loads in flight.
// kSharedHist: the per-label counters of a CTA live in shared memory (num_labels 32-bit slots after the Hilbert table) and
// are added to the global 64-bit histogram once at the end — hot labels (large cells) would otherwise serialise on L2 atomics.
constexpr int kPts = 2;
template <bool kSharedHist>
__global__ void __launch_bounds__(kThreads)
ci_point_histogram_kernel(CellIndexView ix, const double* __restrict__ xyz, size_t n, int num_labels, unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint16_t* s_pos = reinterpret_cast<uint16_t*>(s_raw);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_raw + 2048);
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_ci.pos[k];
  if (kSharedHist) for (int k = threadIdx.x; k < num_labels; k += kThreads) s_hist[k] = 0;
  __syncthreads();
  auto hit = [&](int32_t label) {
    if (kSharedHist) atomicAdd(s_hist + label, 1u); else atomicAdd(counts + label, 1ull);
  };
  const size_t stride = (size_t)gridDim.x * kThreads * kPts;
  auto load_points = [&](size_t base, P3* p) {   // a dead slot (past n) re-reads the thread's first point; it is never used
#pragma unroll
    for (int q = 0; q < kPts; ++q) {
      const size_t t = base + (size_t)q * kThreads;
      const size_t tt = t < n ? t : base;
      p[q] = P3{__ldcs(xyz + 3 * tt), __ldcs(xyz + 3 * tt + 1), __ldcs(xyz + 3 * tt + 2)};
    }
  };
  size_t base = (size_t)blockIdx.x * kThreads * kPts + threadIdx.x;
  P3 p[kPts], p_next[kPts];
  if (base < n) load_points(base, p);
  for (; base < n; base += stride) {
    const size_t next = base + stride;
    if (next < n) load_points(next, p_next);   // the DRAM latency of the next points hides behind this iteration's lookups
    bool live[kPts];
#pragma unroll
    for (int q = 0; q < kPts; ++q) live[q] = base + (size_t)q * kThreads < n;
    // (face, i, j) with K1's filtered projection (IEEE division / square root only near a leaf boundary), then just the three
    // leading steps of FromFaceIJ: enough to name the bucket. The id is finished only if the bucket has to be searched.
    uint32_t pi[kPts], pj[kPts], lo[kPts], hi[kPts];
    HilbertState hs[kPts];
#pragma unroll
    for (int q = 0; q < kPts; ++q) {
      double fi, fj;
      const int face = point_to_face_fij_checked(p[q], &fi, &fj);
      pi[q] = fij_to_ij(fi); pj[q] = fij_to_ij(fj);
      hs[q] = hilbert_steps<7, 5>(hilbert_begin(face), pi[q], pj[q], s_pos);
      const uint64_t b = hs[q].n >> (ix.rshift - 1);   // id = 2n + 1, so id >> rshift == n >> (rshift - 1)
      lo[q] = __ldg(ix.rfirst + b);
      hi[q] = __ldg(ix.rfirst + b + 1);
    }
    CiRangeRec rec[kPts];
#pragma unroll
    for (int q = 0; q < kPts; ++q) {
      if (lo[q] < hi[q]) {   // upper_bound inside the bucket (rare: most buckets hold no range start)
        const uint64_t leaf = hilbert_steps<4, 0>(hs[q], pi[q], pj[q], s_pos).n * 2 + 1;
        do {
          const uint32_t mid = lo[q] + ((hi[q] - lo[q]) >> 1);
          if (__ldg(ix.rstart + mid) <= leaf) lo[q] = mid + 1; else hi[q] = mid;
        } while (lo[q] < hi[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < kPts; ++q) rec[q] = ci_load_rec(ix.rrec + (lo[q] - 1));
#pragma unroll
    for (int q = 0; q < kPts; ++q) if (live[q]) ci_emit_rec_labels(ix, rec[q], hit);
#pragma unroll
    for (int q = 0; q < kPts; ++q) p[q] = p_next[q];
  }
  if (kSharedHist) {
    __syncthreads();
    for (int k = threadIdx.x; k < num_labels; k += kThreads) {
      const uint32_t c = s_hist[k];
      if (c) atomicAdd(counts + k, (unsigned long long)c);
    }
  }
}

struct StreamBuf {   // stream-ordered scratch
  void* p = nullptr;
  cudaStream_t st;
  explicit StreamBuf(cudaStream_t s) : st(s) {}
  ~StreamBuf() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, st); }
};

int check_index(const S2bCellIndex* ix) {
  if (!ix) return set_error(S2B_EINVAL, "null cell index");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the cell index lives on device %d but the current device is %d", ix->device, dev);
  return S2B_OK;
}

// counts -> offsets[T + 1] (exclusive scan in place) and the 64-bit total on the host. Synchronises the stream.
int scan_and_total(uint32_t* offsets_dev, size_t T, unsigned long long* total_dev, cudaStream_t st, unsigned long long* total_host) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, offsets_dev, offsets_dev, (int)(T + 1), st);
  StreamBuf tmp(st);
  cudaError_t e = tmp.alloc(bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(offsets_dev + T, 0, 4, st);
  if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp.p, bytes, offsets_dev, offsets_dev, (int)(T + 1), st);
  count_launch();
  if (e == cudaSuccess) e = cudaMemcpyAsync(total_host, total_dev, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cell index scan");
}

// Count + scan (+ fill when the total fits): the common front half of every CSR query. offsets_dev has T + 1 slots.
int visit_dev(const S2bCellIndex* ix, const Targets& tg, bool all_pairs, uint32_t* offsets_dev, uint64_t* cells_dev, int32_t* labels_dev,
              size_t capacity, bool fill, unsigned long long* total_out, cudaStream_t st) {
  StreamBuf total(st);
  cudaError_t e = total.alloc(8);
  if (e == cudaSuccess) e = cudaMemsetAsync(total.p, 0, 8, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  ci_count_kernel<<<grid_for_ci(tg.T, (const void*)ci_count_kernel), kThreads, 0, st>>>(ix->view, tg, all_pairs, offsets_dev, (unsigned long long*)total.p);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "ci_count launch");
  int rc = scan_and_total(offsets_dev, tg.T, (unsigned long long*)total.p, st, total_out);
  if (rc != S2B_OK) return rc;
  if (*total_out > 0xFFFFFFFFull) return set_error(S2B_ECAPACITY, "more than 2^32 intersecting pairs: split the batch");
  if (!fill || *total_out == 0) return S2B_OK;
  if (*total_out > capacity) return set_error(S2B_ECAPACITY, "capacity %zu < %llu intersecting pairs", capacity, *total_out);
  ci_fill_kernel<<<grid_for_ci(tg.T, (const void*)ci_fill_kernel), kThreads, 0, st>>>(ix->view, tg, Targets{nullptr, nullptr, 0}, offsets_dev, cells_dev, labels_dev, nullptr);
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "ci_fill launch");
}

int check_targets(const uint64_t* ids, const uint32_t* offsets_host_or_null, size_t T) {
  (void)offsets_host_or_null;
  if (T > 0x7FFFFFF0ull) return set_error(S2B_EINVAL, "too many targets (max 2^31)");
  if (T && !ids) return set_error(S2B_EINVAL, "null target ids");
  return S2B_OK;
}

// labels of every target sorted within the target: fill into scratch, then a segmented sort. `sorted` / `weights_sorted`
// receive stream-ordered buffers of *total_out elements; offsets_dev (T + 1) is filled.
int sorted_labels_dev(const S2bCellIndex* ix, const Targets& tg, const Targets& flt, uint32_t* offsets_dev, StreamBuf& sorted, StreamBuf* weights_sorted,
                      unsigned long long* total_out, cudaStream_t st) {
  int rc = visit_dev(ix, tg, weights_sorted != nullptr, offsets_dev, nullptr, nullptr, 0, false, total_out, st);
  if (rc != S2B_OK || *total_out == 0) return rc;
  const size_t total = (size_t)*total_out;
  StreamBuf raw(st), wraw(st), tmp(st);
  cudaError_t e = raw.alloc(4 * total);
  if (e == cudaSuccess) e = sorted.alloc(4 * total);
  if (e == cudaSuccess && weights_sorted) { e = wraw.alloc(8 * total); if (e == cudaSuccess) e = weights_sorted->alloc(8 * total); }
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  ci_fill_kernel<<<grid_for_ci(tg.T, (const void*)ci_fill_kernel), kThreads, 0, st>>>(ix->view, tg, flt, offsets_dev, nullptr, (int32_t*)raw.p,
                                                                                         weights_sorted ? (uint64_t*)wraw.p : nullptr);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "ci_fill launch");
  size_t bytes = 0;
  if (weights_sorted) {
    cub::DeviceSegmentedSort::SortPairs(nullptr, bytes, (const int32_t*)raw.p, (int32_t*)sorted.p, (const uint64_t*)wraw.p, (uint64_t*)weights_sorted->p,
                                        (int64_t)total, (int64_t)tg.T, offsets_dev, offsets_dev + 1, st);
    if ((e = tmp.alloc(bytes)) != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
    e = cub::DeviceSegmentedSort::SortPairs(tmp.p, bytes, (const int32_t*)raw.p, (int32_t*)sorted.p, (const uint64_t*)wraw.p, (uint64_t*)weights_sorted->p,
                                            (int64_t)total, (int64_t)tg.T, offsets_dev, offsets_dev + 1, st);
  } else {
    cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, (const int32_t*)raw.p, (int32_t*)sorted.p, (int64_t)total, (int64_t)tg.T, offsets_dev,
                                       offsets_dev + 1, st);
    if ((e = tmp.alloc(bytes)) != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
    e = cub::DeviceSegmentedSort::SortKeys(tmp.p, bytes, (const int32_t*)raw.p, (int32_t*)sorted.p, (int64_t)total, (int64_t)tg.T, offsets_dev,
                                           offsets_dev + 1, st);
  }
  count_launch();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "segmented sort");
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

// Host-pointer targets -> device copies. n_ids = offsets[T] (or T).
int upload_targets(const uint64_t* ids, const uint32_t* offsets, size_t T, DevBuf& d_ids, DevBuf& d_off, Targets* tg) {
  const size_t n_ids = offsets ? offsets[T] : T;
  if (offsets) {
    if (offsets[0] != 0) return set_error(S2B_EINVAL, "target_offsets[0] must be 0");
    for (size_t t = 0; t < T; ++t) if (offsets[t + 1] < offsets[t]) return set_error(S2B_EINVAL, "target_offsets must not decrease");
  }
  if (n_ids && !ids) return set_error(S2B_EINVAL, "null target ids");
  cudaError_t e = d_ids.alloc(8 * n_ids);
  if (e == cudaSuccess && n_ids) e = cudaMemcpy(d_ids.p, ids, 8 * n_ids, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && offsets) {
    e = d_off.alloc(4 * (T + 1));
    if (e == cudaSuccess) e = cudaMemcpy(d_off.p, offsets, 4 * (T + 1), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc/cudaMemcpy(targets)");
  tg->ids = (const uint64_t*)d_ids.p;
  tg->offsets = offsets ? (const uint32_t*)d_off.p : nullptr;
  tg->T = T;
  return S2B_OK;
}

}  // namespace
}  // namespace s2b

using namespace s2b;

extern "C" {

int s2b_cell_index_create(const uint64_t* cell_ids, const int32_t* labels, size_t m, S2bCellIndex** out) {
  if (!out) return set_error(S2B_EINVAL, "null output");
  *out = nullptr;
  if (m && (!cell_ids || !labels)) return set_error(S2B_EINVAL, "null pointer");
  if (m > 0x7FFFFFF0ull) return set_error(S2B_EINVAL, "cell index too large (max 2^31 pairs)");
  CellIndexHost h;
  try {
    if (!build_cell_index_host(cell_ids, labels, m, &h)) return set_error(S2B_EINVAL, "invalid cell id or negative label (S2CellIndex::Add requires both)");
  } catch (const std::bad_alloc&) {
    return set_error(S2B_ENOMEM, "host allocation failed");
  }
  S2bCellIndex* ix = new (std::nothrow) S2bCellIndex();
  if (!ix) return set_error(S2B_ENOMEM, "host allocation failed");
  cudaError_t e = cudaGetDevice(&ix->device);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaGetDevice"); }
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t R = h.rstart.size();
  const size_t sizes[6] = {sizeof(CiNode) * (m + 1), 4 * h.first.size(), 8 * R, 4 * h.rfirst.size(), sizeof(CiRangeRec) * R, 4 * (h.llabels.size() + 1)};
  const void* src[6] = {h.nodes.data(), h.first.data(), h.rstart.data(), h.rfirst.data(), h.rrec.data(), h.llabels.data()};
  const size_t copy[6] = {sizeof(CiNode) * m, 4 * h.first.size(), 8 * R, 4 * h.rfirst.size(), sizeof(CiRangeRec) * R, 4 * h.llabels.size()};
  size_t off[7] = {0};
  for (int k = 0; k < 6; ++k) off[k + 1] = off[k] + up(sizes[k]);
  ix->blob_bytes = off[6];
  e = cudaMalloc(&ix->blob, ix->blob_bytes ? ix->blob_bytes : 256);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc(cell index)"); }
  char* base = (char*)ix->blob;
  for (int k = 0; k < 6 && e == cudaSuccess; ++k)
    if (copy[k]) e = cudaMemcpy(base + off[k], src[k], copy[k], cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(ix->blob); delete ix; return cuda_fail(e, "cudaMemcpy(cell index)"); }
  ix->num_labels = h.num_labels;
  ix->view = CellIndexView{(const CiNode*)base, (const uint32_t*)(base + off[1]), (uint32_t)m, 61 - 2 * h.table_level,
                           (uint32_t)(6u << (2 * h.table_level)), (const uint64_t*)(base + off[2]), (const uint32_t*)(base + off[3]),
                           (const CiRangeRec*)(base + off[4]), (const int32_t*)(base + off[5]), (uint32_t)R, 61 - 2 * h.range_table_level};
  *out = ix;
  return S2B_OK;
}

void s2b_cell_index_destroy(S2bCellIndex* ix) {
  if (!ix) return;
  if (ix->blob) cudaFree(ix->blob);
  delete ix;
}

int s2b_cell_index_info(const S2bCellIndex* ix, uint64_t* num_cells, int32_t* num_labels, uint64_t* device_bytes) {
  if (!ix) return set_error(S2B_EINVAL, "null cell index");
  if (num_cells) *num_cells = ix->view.m;
  if (num_labels) *num_labels = ix->num_labels;
  if (device_bytes) *device_bytes = ix->blob_bytes;
  return S2B_OK;
}

int s2b_cell_index_intersecting_cells_dev(const S2bCellIndex* ix, const uint64_t* target_ids_dev, const uint32_t* target_offsets_dev, size_t T,
                                          uint32_t* offsets_out_dev, uint64_t* cell_ids_out_dev, int32_t* labels_out_dev, size_t capacity,
                                          size_t* total_out, void* stream) {
  int rc = check_index(ix);
  if (rc == S2B_OK) rc = check_targets(target_ids_dev, nullptr, T);
  if (rc != S2B_OK) return rc;
  if (!offsets_out_dev) return set_error(S2B_EINVAL, "null offsets_out");
  if (total_out) *total_out = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {
    cudaError_t e = cudaMemsetAsync(offsets_out_dev, 0, 4, st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cudaMemsetAsync");
  }
  unsigned long long total = 0;
  rc = visit_dev(ix, Targets{target_ids_dev, target_offsets_dev, T}, false, offsets_out_dev, cell_ids_out_dev, labels_out_dev, capacity, true, &total, st);
  if (total_out) *total_out = (size_t)total;
  return rc;
}

int s2b_cell_index_intersecting_labels_dev(const S2bCellIndex* ix, const uint64_t* target_ids_dev, const uint32_t* target_offsets_dev, size_t T,
                                           uint32_t* offsets_out_dev, int32_t* labels_out_dev, size_t capacity, size_t* total_out, void* stream) {
  int rc = check_index(ix);
  if (rc == S2B_OK) rc = check_targets(target_ids_dev, nullptr, T);
  if (rc != S2B_OK) return rc;
  if (!offsets_out_dev) return set_error(S2B_EINVAL, "null offsets_out");
  if (total_out) *total_out = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {
    cudaError_t e = cudaMemsetAsync(offsets_out_dev, 0, 4, st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cudaMemsetAsync");
  }
  const Targets tg{target_ids_dev, target_offsets_dev, T};
  StreamBuf pair_offsets(st), sorted(st), total_dev(st);
  cudaError_t e = pair_offsets.alloc(4 * (T + 1));
  if (e == cudaSuccess) e = total_dev.alloc(8);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  unsigned long long pairs = 0, total = 0;
  rc = sorted_labels_dev(ix, tg, Targets{nullptr, nullptr, 0}, (uint32_t*)pair_offsets.p, sorted, nullptr, &pairs, st);
  if (rc != S2B_OK) return rc;
  if (pairs == 0) {
    e = cudaMemsetAsync(offsets_out_dev, 0, 4 * (T + 1), st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cudaMemsetAsync");
  }
  if ((e = cudaMemsetAsync(total_dev.p, 0, 8, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  ci_unique_count_kernel<<<grid_for_ci(T, (const void*)ci_unique_count_kernel), kThreads, 0, st>>>((const int32_t*)sorted.p, (const uint32_t*)pair_offsets.p, T,
                                                                                                    offsets_out_dev, (unsigned long long*)total_dev.p);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "ci_unique_count launch");
  rc = scan_and_total(offsets_out_dev, T, (unsigned long long*)total_dev.p, st, &total);
  if (rc != S2B_OK) return rc;
  if (total_out) *total_out = (size_t)total;
  if (total > capacity) return set_error(S2B_ECAPACITY, "capacity %zu < %llu labels", capacity, total);
  if (!labels_out_dev) return set_error(S2B_EINVAL, "null labels_out");
  ci_unique_fill_kernel<<<grid_for_ci(T, (const void*)ci_unique_fill_kernel), kThreads, 0, st>>>((const int32_t*)sorted.p, (const uint32_t*)pair_offsets.p, T,
                                                                                                  offsets_out_dev, labels_out_dev);
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "ci_unique_fill launch");
}

int s2b_cell_index_most_intersecting_label_dev(const S2bCellIndex* ix, const uint64_t* bound_ids_dev, const uint32_t* bound_offsets_dev,
                                               const uint64_t* region_ids_dev, const uint32_t* region_offsets_dev, size_t T,
                                               int32_t default_label, int32_t* out_dev, void* stream) {
  int rc = check_index(ix);
  if (rc == S2B_OK) rc = check_targets(bound_ids_dev, nullptr, T);
  if (rc != S2B_OK) return rc;
  if (T == 0) return S2B_OK;
  if (!out_dev) return set_error(S2B_EINVAL, "null output");
  cudaStream_t st = (cudaStream_t)stream;
  const Targets tg{bound_ids_dev, bound_offsets_dev, T};
  const Targets flt{region_ids_dev, region_offsets_dev, T};
  StreamBuf pair_offsets(st), sorted(st), weights(st);
  cudaError_t e = pair_offsets.alloc(4 * (T + 1));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  unsigned long long pairs = 0;
  rc = sorted_labels_dev(ix, tg, flt, (uint32_t*)pair_offsets.p, sorted, &weights, &pairs, st);
  if (rc != S2B_OK) return rc;
  ci_best_label_kernel<<<grid_for_ci(T, (const void*)ci_best_label_kernel), kThreads, 0, st>>>((const int32_t*)sorted.p, (const uint64_t*)weights.p,
                                                                                                (const uint32_t*)pair_offsets.p, T, default_label, out_dev);
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "ci_best_label launch");
}

int s2b_cell_index_point_histogram_dev(const S2bCellIndex* ix, const double* xyz_dev, size_t n, uint64_t* counts_dev, void* stream) {
  int rc = check_index(ix);
  if (rc != S2B_OK) return rc;
  if (n == 0) return S2B_OK;
  if (!xyz_dev || (ix->num_labels && !counts_dev)) return set_error(S2B_EINVAL, "null pointer");
  // up to 16K labels the CTA-private histogram fits next to the Hilbert table (<= 66 KB, 3+ CTAs per SM)
  const bool shared_hist = ix->num_labels <= 16384;
  const size_t smem = 2048 + (shared_hist ? 4 * (size_t)ix->num_labels : 0);
  const void* kernel = shared_hist ? (const void*)ci_point_histogram_kernel<true> : (const void*)ci_point_histogram_kernel<false>;
  cudaError_t ea = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ea != cudaSuccess) return cuda_fail(ea, "cudaFuncSetAttribute");
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
  if (per_sm < 1) per_sm = 1;
  const size_t want = (n + kThreads * kPts - 1) / (kThreads * kPts), cap = (size_t)sm_count() * per_sm;
  const int grid = (int)(want < cap ? want : cap);
  if (shared_hist)
    ci_point_histogram_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(ix->view, xyz_dev, n, ix->num_labels, (unsigned long long*)counts_dev);
  else
    ci_point_histogram_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(ix->view, xyz_dev, n, ix->num_labels, (unsigned long long*)counts_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "ci_point_histogram launch");
}

// ---- host-pointer forms (utility-sized inputs: whole-array copies)

int s2b_cell_index_intersecting_cells(const S2bCellIndex* ix, const uint64_t* target_ids, const uint32_t* target_offsets, size_t T,
                                      uint32_t* offsets_out, uint64_t* cell_ids_out, int32_t* labels_out, size_t capacity, size_t* total_out) {
  int rc = check_index(ix);
  if (rc == S2B_OK) rc = check_targets(target_ids, target_offsets, T);
  if (rc != S2B_OK) return rc;
  if (!offsets_out) return set_error(S2B_EINVAL, "null offsets_out");
  if (total_out) *total_out = 0;
  offsets_out[0] = 0;
  if (T == 0) return S2B_OK;
  DevBuf d_ids, d_off, d_out_off, d_cells, d_labels;
  Targets tg{};
  rc = upload_targets(target_ids, target_offsets, T, d_ids, d_off, &tg);
  if (rc != S2B_OK) return rc;
  cudaError_t e = d_out_off.alloc(4 * (T + 1));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  unsigned long long total = 0;
  rc = visit_dev(ix, tg, false, (uint32_t*)d_out_off.p, nullptr, nullptr, 0, false, &total, nullptr);
  if (total_out) *total_out = (size_t)total;
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(offsets_out, d_out_off.p, 4 * (T + 1), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (total == 0) return S2B_OK;
  if (total > capacity) return set_error(S2B_ECAPACITY, "capacity %zu < %llu intersecting pairs", capacity, total);
  if (cell_ids_out && (e = d_cells.alloc(8 * total)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if (labels_out && (e = d_labels.alloc(4 * total)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  ci_fill_kernel<<<grid_for_ci(T, (const void*)ci_fill_kernel), kThreads>>>(ix->view, tg, Targets{nullptr, nullptr, 0}, (const uint32_t*)d_out_off.p, (uint64_t*)d_cells.p,
                                                                             (int32_t*)d_labels.p, nullptr);
  count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "ci_fill launch");
  if (cell_ids_out && (e = cudaMemcpy(cell_ids_out, d_cells.p, 8 * total, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (labels_out && (e = cudaMemcpy(labels_out, d_labels.p, 4 * total, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cell_index_intersecting_labels(const S2bCellIndex* ix, const uint64_t* target_ids, const uint32_t* target_offsets, size_t T,
                                       uint32_t* offsets_out, int32_t* labels_out, size_t capacity, size_t* total_out) {
  int rc = check_index(ix);
  if (rc == S2B_OK) rc = check_targets(target_ids, target_offsets, T);
  if (rc != S2B_OK) return rc;
  if (!offsets_out) return set_error(S2B_EINVAL, "null offsets_out");
  if (total_out) *total_out = 0;
  offsets_out[0] = 0;
  if (T == 0) return S2B_OK;
  DevBuf d_ids, d_off, d_out_off, d_labels;
  Targets tg{};
  rc = upload_targets(target_ids, target_offsets, T, d_ids, d_off, &tg);
  if (rc != S2B_OK) return rc;
  cudaError_t e = d_out_off.alloc(4 * (T + 1));
  if (capacity && !labels_out) return set_error(S2B_EINVAL, "null labels_out");
  if (e == cudaSuccess) e = d_labels.alloc(4 * capacity);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  size_t total = 0;
  rc = s2b_cell_index_intersecting_labels_dev(ix, tg.ids, tg.offsets, T, (uint32_t*)d_out_off.p, (int32_t*)d_labels.p, capacity, &total, nullptr);
  if (total_out) *total_out = total;
  if (rc != S2B_OK && rc != S2B_ECAPACITY) return rc;
  if ((e = cudaMemcpy(offsets_out, d_out_off.p, 4 * (T + 1), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_ECAPACITY) return rc;
  if (total && (e = cudaMemcpy(labels_out, d_labels.p, 4 * total, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cell_index_most_intersecting_label(const S2bCellIndex* ix, const uint64_t* bound_ids, const uint32_t* bound_offsets,
                                           const uint64_t* region_ids, const uint32_t* region_offsets, size_t T, int32_t default_label, int32_t* out) {
  int rc = check_index(ix);
  if (rc == S2B_OK) rc = check_targets(bound_ids, bound_offsets, T);
  if (rc != S2B_OK) return rc;
  if (T == 0) return S2B_OK;
  if (!out) return set_error(S2B_EINVAL, "null output");
  if (region_ids && !region_offsets) return set_error(S2B_EINVAL, "region_ids needs region_offsets");
  DevBuf d_ids, d_off, d_rids, d_roff, d_out;
  Targets tg{}, flt{nullptr, nullptr, T};
  rc = upload_targets(bound_ids, bound_offsets, T, d_ids, d_off, &tg);
  if (rc == S2B_OK && region_ids) rc = upload_targets(region_ids, region_offsets, T, d_rids, d_roff, &flt);
  if (rc != S2B_OK) return rc;
  cudaError_t e = d_out.alloc(4 * T);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  rc = s2b_cell_index_most_intersecting_label_dev(ix, tg.ids, tg.offsets, flt.ids, flt.offsets, T, default_label, (int32_t*)d_out.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(out, d_out.p, 4 * T, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cell_index_point_histogram(const S2bCellIndex* ix, const double* xyz, size_t n, uint64_t* counts) {
  int rc = check_index(ix);
  if (rc != S2B_OK) return rc;
  if (ix->num_labels && !counts) return set_error(S2B_EINVAL, "null counts");
  for (int32_t l = 0; l < ix->num_labels; ++l) counts[l] = 0;
  if (n == 0 || ix->num_labels == 0) return S2B_OK;
  if (!xyz) return set_error(S2B_EINVAL, "null points");
  DevBuf d_counts, d_xyz;
  const size_t chunk = (size_t)1 << 26;   // 64M points (1.5 GB) per copy
  cudaError_t e = d_counts.alloc(8 * (size_t)ix->num_labels);
  if (e == cudaSuccess) e = d_xyz.alloc(24 * (n < chunk ? n : chunk));
  if (e == cudaSuccess) e = cudaMemset(d_counts.p, 0, 8 * (size_t)ix->num_labels);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  for (size_t b = 0; b < n; b += chunk) {
    const size_t k = n - b < chunk ? n - b : chunk;
    if ((e = cudaMemcpy(d_xyz.p, xyz + 3 * b, 24 * k, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
    rc = s2b_cell_index_point_histogram_dev(ix, (const double*)d_xyz.p, k, (uint64_t*)d_counts.p, nullptr);
    if (rc != S2B_OK) return rc;
  }
  if ((e = cudaMemcpy(counts, d_counts.p, 8 * (size_t)ix->num_labels, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

}  // extern "C"
