// s2b_cellops.cu — batched S2CellId / S2CellUnion algebra that callers run next to the encode path:
//   S2CellId::AppendVertexNeighbors   s2cell_id.cc:514-556
//   S2CellId::AppendAllNeighbors      s2cell_id.cc:558-598
//   S2CellUnion::Normalize            s2cell_union.cc:171-199   (+ AreSiblings :127-143)
//   S2CellUnion::Contains/Intersects(S2CellId)   s2cell_union.cc:285-309
// Integer work, bit-exact with the reference. One thread per id; the Hilbert tables live in shared memory.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "../../include/s2b_capi.h"
#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_internal.h"

namespace s2b {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

namespace {

constexpr int kThreads = 256;
__device__ const HilbertTables g_hilbert_ops = kHilbert;

int grid_for_ops(size_t n, const void* kernel) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  size_t want = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

// S2CellId::FromFaceIJSame (s2cell_id.h): on the same face the plain encoding, otherwise wrap onto the adjacent face.
__device__ __forceinline__ uint64_t from_face_ij_same(int face, int i, int j, bool same, const uint16_t* pos) {
  return same ? from_face_ij(face, (uint32_t)i, (uint32_t)j, pos) : from_face_ij_wrap(face, i, j, pos);
}

__device__ __forceinline__ bool cellid_is_valid(uint64_t id) {   // s2cell_id.h:583-585
  return (id >> 61) < 6 && (cellid_lsb(id) & 0x1555555555555555ull) != 0;
}

__global__ void __launch_bounds__(kThreads)
vertex_neighbors_kernel(const uint64_t* __restrict__ ids, size_t n, int level, uint64_t* __restrict__ out, uint8_t* __restrict__ count) {
  __shared__ uint16_t s_pos[1024], s_ij[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) { s_pos[k] = g_hilbert_ops.pos[k]; s_ij[k] = g_hilbert_ops.ij[k]; }
  __syncthreads();
  const int kMaxSize = 1 << 30;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t id = __ldcs(ids + t);
    uint64_t nb[4] = {0, 0, 0, 0};
    int cnt = 0;
    if (cellid_is_valid(id) && level < cellid_level(id)) {   // the reference requires level < this->level()
      int i, j;
      const int face = cellid_to_face_ij(id, s_ij, &i, &j);
      const int halfsize = 1 << (30 - (level + 1));
      const int size = halfsize << 1;
      bool isame, jsame;
      int ioffset, joffset;
      if (i & halfsize) { ioffset = size; isame = (i + size) < kMaxSize; } else { ioffset = -size; isame = (i - size) >= 0; }
      if (j & halfsize) { joffset = size; jsame = (j + size) < kMaxSize; } else { joffset = -size; jsame = (j - size) >= 0; }
      nb[0] = cellid_parent(id, level);
      nb[1] = cellid_parent(from_face_ij_same(face, i + ioffset, j, isame, s_pos), level);
      nb[2] = cellid_parent(from_face_ij_same(face, i, j + joffset, jsame, s_pos), level);
      cnt = 3;
      if (isame || jsame) nb[cnt++] = cellid_parent(from_face_ij_same(face, i + ioffset, j + joffset, isame && jsame, s_pos), level);
    }
    ulonglong2* o = (ulonglong2*)(out + 4 * t);
    __stcs(o, make_ulonglong2(nb[0], nb[1]));
    __stcs(o + 1, make_ulonglong2(nb[2], nb[3]));
    if (count) count[t] = (uint8_t)cnt;
  }
}

__global__ void __launch_bounds__(kThreads)
all_neighbors_kernel(const uint64_t* __restrict__ ids, size_t n, int nbr_level, uint64_t* __restrict__ out, size_t out_stride,
                     uint32_t* __restrict__ count) {
  __shared__ uint16_t s_pos[1024], s_ij[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) { s_pos[k] = g_hilbert_ops.pos[k]; s_ij[k] = g_hilbert_ops.ij[k]; }
  __syncthreads();
  const int kMaxSize = 1 << 30;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t id = __ldcs(ids + t);
    uint64_t* o = out + t * out_stride;
    uint64_t cnt = 0, total = 0;
    auto push = [&](uint64_t v) { if (cnt < out_stride) o[cnt] = v; ++cnt; };
    if (cellid_is_valid(id) && nbr_level >= cellid_level(id)) {
      int i, j;
      const int face = cellid_to_face_ij(id, s_ij, &i, &j);
      const int size = 1 << (30 - cellid_level(id));
      i &= -size;
      j &= -size;
      const int nbr_size = 1 << (30 - nbr_level);
      total = 4ull * (uint64_t)(size / nbr_size) + 4;
      for (int k = -nbr_size; cnt < out_stride; k += nbr_size) {     // (stops early once the caller's slots are full)
        bool same_face;
        if (k < 0) {
          same_face = (j + k >= 0);
        } else if (k >= size) {
          same_face = (j + k < kMaxSize);
        } else {
          same_face = true;
          push(cellid_parent(from_face_ij_same(face, i + k, j - nbr_size, j - size >= 0, s_pos), nbr_level));
          push(cellid_parent(from_face_ij_same(face, i + k, j + size, j + size < kMaxSize, s_pos), nbr_level));
        }
        push(cellid_parent(from_face_ij_same(face, i - nbr_size, j + k, same_face && i - size >= 0, s_pos), nbr_level));
        push(cellid_parent(from_face_ij_same(face, i + size, j + k, same_face && i + size < kMaxSize, s_pos), nbr_level));
        if (k >= size) break;
      }
      if (cnt > out_stride) cnt = out_stride;
    }
    for (uint64_t k = cnt; k < out_stride; ++k) o[k] = 0;
    if (count) count[t] = total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)total;
  }
}

// out[t] bit0 = union.Contains(id), bit1 = union.Intersects(id): lower_bound with EntirelyPrecedes (s2cell_union.cc:280-309).
__global__ void __launch_bounds__(kThreads)
union_cells_kernel(const uint64_t* __restrict__ u, size_t m, const uint64_t* __restrict__ ids, size_t n, uint8_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t id = __ldcs(ids + t);
    const uint64_t lo_id = cellid_range_min(id), hi_id = cellid_range_max(id);
    size_t lo = 0, hi = m;                       // first cell with range_max >= range_min(id)
    while (lo < hi) {
      const size_t mid = (lo + hi) >> 1;
      if (cellid_range_max(u[mid]) < lo_id) lo = mid + 1; else hi = mid;
    }
    uint8_t r = 0;
    if (lo < m) {
      const uint64_t c = u[lo], cmin = cellid_range_min(c), cmax = cellid_range_max(c);
      if (id >= cmin && id <= cmax) r |= 1;                  // S2CellId::contains (s2cell_id.h:648-652)
      if (lo_id <= cmax && hi_id >= cmin) r |= 2;            // S2CellId::intersects (s2cell_id.h:656-660)
    }
    out[t] = r;
  }
}

// S2CellId::is_valid / level / range_min / range_max (s2cell_id.h:583-603, 642-652), any output nullable. level is -1 for
// invalid ids (the reference's level() is undefined there).
__global__ void __launch_bounds__(kThreads)
cellid_info_kernel(const uint64_t* __restrict__ ids, size_t n, uint8_t* __restrict__ valid, int8_t* __restrict__ level,
                   uint64_t* __restrict__ rmin, uint64_t* __restrict__ rmax) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t id = __ldcs(ids + t);
    const bool ok = cellid_is_valid(id);
    if (valid) valid[t] = ok ? 1 : 0;
    if (level) level[t] = ok ? (int8_t)cellid_level(id) : (int8_t)-1;
    if (rmin) rmin[t] = cellid_range_min(id);
    if (rmax) rmax[t] = cellid_range_max(id);
  }
}

// S2CellId::child(0..3), advance / advance_wrap, GetCommonAncestorLevel element-wise (any output may be null).
__global__ void __launch_bounds__(kThreads)
cellid_traverse_kernel(const uint64_t* __restrict__ ids, const uint64_t* __restrict__ others, const long long* __restrict__ steps, size_t n, int wrap,
                       uint64_t* __restrict__ children4, uint64_t* __restrict__ advanced, int8_t* __restrict__ ancestor_level) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t id = __ldcs(ids + t);
    const bool ok = cellid_is_valid(id);
    if (children4) {
      const bool parent = ok && !(id & 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) children4[4 * t + k] = parent ? cellid_child(id, k) : 0;
    }
    if (advanced) advanced[t] = ok ? cellid_advance(id, steps[t], wrap != 0) : 0;
    if (ancestor_level) {
      const uint64_t other = others[t];
      ancestor_level[t] = (ok && cellid_is_valid(other)) ? (int8_t)cellid_common_ancestor_level(id, other) : (int8_t)-1;
    }
  }
}

// ---- Normalize -----------------------------------------------------------------------------------------------------
struct MaxOp { __host__ __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; } };
struct MinOp { __host__ __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a < b ? a : b; } };

// rmax[i] = range_max(ids[i]); rmin_rev[m-1-i] = range_min(ids[i])  (so a forward scan of rmin_rev is a suffix scan)
__global__ void ranges_kernel(const uint64_t* __restrict__ ids, size_t m, uint64_t* __restrict__ rmax, uint64_t* __restrict__ rmin_rev) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const uint64_t id = ids[i];
    rmax[i] = cellid_range_max(id);
    rmin_rev[m - 1 - i] = cellid_range_min(id);
  }
}

// ids sorted and distinct. Cell i is redundant iff another cell contains it: cell ranges are nested or disjoint, so that
// is the case iff an earlier cell reaches at least as far (prefix max of range_max) or a later cell starts at least as
// early (suffix min of range_min).
__global__ void contained_kernel(const uint64_t* __restrict__ ids, size_t m, const uint64_t* __restrict__ pmax, const uint64_t* __restrict__ smin_rev,
                                 uint8_t* __restrict__ keep) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const uint64_t id = ids[i];
    bool contained = false;
    if (i > 0 && pmax[i - 1] >= cellid_range_max(id)) contained = true;
    if (i + 1 < m && smin_rev[m - 2 - i] <= cellid_range_min(id)) contained = true;
    keep[i] = contained ? 0 : 1;
  }
}

__device__ __forceinline__ bool are_siblings(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {   // s2cell_union.cc:127-143
  if ((a ^ b ^ c) != d) return false;
  uint64_t mask = cellid_lsb(d) << 1;
  mask = ~(mask + (mask << 1));
  const uint64_t id_masked = d & mask;
  return (a & mask) == id_masked && (b & mask) == id_masked && (c & mask) == id_masked && cellid_lsb(d) != (1ull << 60);
}

// One merge round over sorted, disjoint ids: a run of four siblings becomes its parent (written at the first of the
// four), the other three are dropped. Runs cannot overlap (the first of a run is child 0 of its parent).
__global__ void merge_round_kernel(const uint64_t* __restrict__ ids, size_t m, uint64_t* __restrict__ vals, uint8_t* __restrict__ keep) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    auto head = [&](size_t h) { return h + 3 < m && are_siblings(ids[h], ids[h + 1], ids[h + 2], ids[h + 3]); };
    uint64_t v = ids[i];
    uint8_t k = 1;
    if (head(i)) {
      const uint64_t lsb = cellid_lsb(v) << 2;
      v = (v & (~lsb + 1)) | lsb;                                    // S2CellId::parent()
    } else if ((i >= 1 && head(i - 1)) || (i >= 2 && head(i - 2)) || (i >= 3 && head(i - 3))) {
      k = 0;
    }
    vals[i] = v;
    keep[i] = k;
  }
}

// Normalizes m ids (device) into out_dev (capacity m); *count_out (host) = number of cells. Synchronises `st`.
int normalize_dev(const uint64_t* ids_dev, size_t m, uint64_t* out_dev, size_t* count_out, cudaStream_t st) {
  *count_out = 0;
  if (m == 0) return S2B_OK;
  if (m > 0x7FFFFFFFull) return set_error(S2B_EINVAL, "cell union too large (max 2^31-1 ids)");
  const int im = (int)m;
  size_t t_sort = 0, t_uniq = 0, t_scan = 0, t_sel = 0;
  uint64_t* nul64 = nullptr; uint8_t* nul8 = nullptr; int* nuli = nullptr;
  cub::DeviceRadixSort::SortKeys(nullptr, t_sort, nul64, nul64, im, 0, 64, st);
  cub::DeviceSelect::Unique(nullptr, t_uniq, nul64, nul64, nuli, im, st);
  cub::DeviceScan::InclusiveScan(nullptr, t_scan, nul64, nul64, MaxOp(), im, st);
  cub::DeviceSelect::Flagged(nullptr, t_sel, nul64, nul8, nul64, nuli, im, st);
  size_t t_bytes = t_sort;
  if (t_uniq > t_bytes) t_bytes = t_uniq;
  if (t_scan > t_bytes) t_bytes = t_scan;
  if (t_sel > t_bytes) t_bytes = t_sel;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t b_ids = up(8 * m), b_flag = up(m);
  char* blob = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&blob, 4 * b_ids + b_flag + 256 + up(t_bytes), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(normalize)");
  uint64_t* a = (uint64_t*)blob;
  uint64_t* b = (uint64_t*)(blob + b_ids);
  uint64_t* c = (uint64_t*)(blob + 2 * b_ids);
  uint64_t* d = (uint64_t*)(blob + 3 * b_ids);
  uint8_t* keep = (uint8_t*)(blob + 4 * b_ids);
  int* d_count = (int*)(blob + 4 * b_ids + b_flag);
  void* tmp = blob + 4 * b_ids + b_flag + 256;
  int rc = S2B_OK;
  int cnt = 0;
  auto fetch_count = [&]() {
    cudaError_t ee = cudaMemcpyAsync(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ee == cudaSuccess) ee = cudaStreamSynchronize(st);
    return ee;
  };
  do {
    size_t tb = t_bytes;
    if ((e = cub::DeviceRadixSort::SortKeys(tmp, tb, ids_dev, a, im, 0, 64, st)) != cudaSuccess) break;
    count_launch();
    tb = t_bytes;
    if ((e = cub::DeviceSelect::Unique(tmp, tb, a, b, d_count, im, st)) != cudaSuccess) break;
    count_launch();
    if ((e = fetch_count()) != cudaSuccess) break;
    // b[0..cnt): sorted distinct ids. Drop the ones contained in another cell.
    const int g = grid_for_ops((size_t)cnt, (const void*)ranges_kernel);
    ranges_kernel<<<g, kThreads, 0, st>>>(b, (size_t)cnt, c, d);
    count_launch();
    tb = t_bytes;
    if ((e = cub::DeviceScan::InclusiveScan(tmp, tb, c, c, MaxOp(), cnt, st)) != cudaSuccess) break;
    tb = t_bytes;
    if ((e = cub::DeviceScan::InclusiveScan(tmp, tb, d, d, MinOp(), cnt, st)) != cudaSuccess) break;
    count_launch(); count_launch();
    contained_kernel<<<g, kThreads, 0, st>>>(b, (size_t)cnt, c, d, keep);
    count_launch();
    tb = t_bytes;
    if ((e = cub::DeviceSelect::Flagged(tmp, tb, b, keep, a, d_count, cnt, st)) != cudaSuccess) break;
    count_launch();
    if ((e = fetch_count()) != cudaSuccess) break;
    // a[0..cnt): sorted disjoint cells. Merge sibling quads until nothing changes (at most 30 rounds).
    uint64_t* cur = a;
    uint64_t* nxt = b;
    for (int round = 0; round < 31 && cnt >= 4; ++round) {
      const int before = cnt;
      merge_round_kernel<<<grid_for_ops((size_t)cnt, (const void*)merge_round_kernel), kThreads, 0, st>>>(cur, (size_t)cnt, c, keep);
      count_launch();
      tb = t_bytes;
      if ((e = cub::DeviceSelect::Flagged(tmp, tb, c, keep, nxt, d_count, cnt, st)) != cudaSuccess) break;
      count_launch();
      if ((e = fetch_count()) != cudaSuccess) break;
      uint64_t* sw = cur; cur = nxt; nxt = sw;
      if (cnt == before) break;
    }
    if (e != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(out_dev, cur, 8 * (size_t)cnt, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (false);
  if (e != cudaSuccess) rc = cuda_fail(e, "cell union normalize");
  cudaFreeAsync(blob, st);
  if (rc == S2B_OK) *count_out = (size_t)cnt;
  return rc;
}

struct DevBuf {   // small RAII helper for the host-pointer entry points
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

}  // namespace
}  // namespace s2b

using namespace s2b;

extern "C" {

int s2b_cellid_vertex_neighbors_dev(const uint64_t* ids_dev, size_t n, int level, uint64_t* neighbors4_out_dev, uint8_t* count_out_dev, void* stream) {
  if (level < 0 || level > 29) return set_error(S2B_EINVAL, "level %d out of range [0,29]", level);
  if (n == 0) return S2B_OK;
  if (!ids_dev || !neighbors4_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if ((uintptr_t)neighbors4_out_dev % 16) return set_error(S2B_EINVAL, "neighbors4_out_dev must be 16-byte aligned");
  vertex_neighbors_kernel<<<grid_for_ops(n, (const void*)vertex_neighbors_kernel), kThreads, 0, (cudaStream_t)stream>>>(ids_dev, n, level, neighbors4_out_dev, count_out_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "vertex_neighbors launch");
}

int s2b_cellid_info_dev(const uint64_t* ids_dev, size_t n, uint8_t* valid_out_dev, int8_t* level_out_dev, uint64_t* range_min_out_dev,
                        uint64_t* range_max_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev) return set_error(S2B_EINVAL, "null pointer");
  cellid_info_kernel<<<grid_for_ops(n, (const void*)cellid_info_kernel), kThreads, 0, (cudaStream_t)stream>>>(ids_dev, n, valid_out_dev, level_out_dev,
                                                                                                            range_min_out_dev, range_max_out_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cellid_info launch");
}

int s2b_cellid_all_neighbors_dev(const uint64_t* ids_dev, size_t n, int nbr_level, uint64_t* out_dev, size_t out_stride, uint32_t* count_out_dev, void* stream) {
  if (nbr_level < 0 || nbr_level > 30) return set_error(S2B_EINVAL, "nbr_level %d out of range [0,30]", nbr_level);
  if (n == 0) return S2B_OK;
  if (!ids_dev || !out_dev || out_stride == 0) return set_error(S2B_EINVAL, "null pointer or zero stride");
  all_neighbors_kernel<<<grid_for_ops(n, (const void*)all_neighbors_kernel), kThreads, 0, (cudaStream_t)stream>>>(ids_dev, n, nbr_level, out_dev, out_stride, count_out_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "all_neighbors launch");
}

int s2b_cellunion_contains_cells_dev(const uint64_t* sorted_ids_dev, size_t m, const uint64_t* ids_dev, size_t n, uint8_t* out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if ((m && !sorted_ids_dev) || !ids_dev || !out_dev) return set_error(S2B_EINVAL, "null pointer");
  union_cells_kernel<<<grid_for_ops(n, (const void*)union_cells_kernel), kThreads, 0, (cudaStream_t)stream>>>(sorted_ids_dev, m, ids_dev, n, out_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cellunion_contains_cells launch");
}

int s2b_cellunion_normalize_dev(const uint64_t* ids_dev, size_t m, uint64_t* out_dev, size_t* count_out, void* stream) {
  if (!count_out) return set_error(S2B_EINVAL, "null count_out");
  if (m && (!ids_dev || !out_dev)) return set_error(S2B_EINVAL, "null pointer");
  return normalize_dev(ids_dev, m, out_dev, count_out, (cudaStream_t)stream);
}

// Host-pointer forms: whole-array copies (these are utility-sized inputs, not the streaming hot path).
int s2b_cellid_vertex_neighbors(const uint64_t* ids, size_t n, int level, uint64_t* neighbors4_out, uint8_t* count_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !neighbors4_out) return set_error(S2B_EINVAL, "null pointer");
  DevBuf in, out, cnt;
  cudaError_t e;
  if ((e = in.alloc(8 * n)) != cudaSuccess || (e = out.alloc(32 * n)) != cudaSuccess || (e = cnt.alloc(n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(in.p, ids, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellid_vertex_neighbors_dev((const uint64_t*)in.p, n, level, (uint64_t*)out.p, (uint8_t*)cnt.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(neighbors4_out, out.p, 32 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (count_out && (e = cudaMemcpy(count_out, cnt.p, n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cellid_info(const uint64_t* ids, size_t n, uint8_t* valid_out, int8_t* level_out, uint64_t* range_min_out, uint64_t* range_max_out) {
  if (n == 0) return S2B_OK;
  if (!ids) return set_error(S2B_EINVAL, "null pointer");
  DevBuf in, v, l, a, b;
  cudaError_t e;
  if ((e = in.alloc(8 * n)) != cudaSuccess || (e = v.alloc(n)) != cudaSuccess || (e = l.alloc(n)) != cudaSuccess || (e = a.alloc(8 * n)) != cudaSuccess ||
      (e = b.alloc(8 * n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(in.p, ids, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellid_info_dev((const uint64_t*)in.p, n, (uint8_t*)v.p, (int8_t*)l.p, (uint64_t*)a.p, (uint64_t*)b.p, nullptr);
  if (rc != S2B_OK) return rc;
  if (valid_out && (e = cudaMemcpy(valid_out, v.p, n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (level_out && (e = cudaMemcpy(level_out, l.p, n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (range_min_out && (e = cudaMemcpy(range_min_out, a.p, 8 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (range_max_out && (e = cudaMemcpy(range_max_out, b.p, 8 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cellid_all_neighbors(const uint64_t* ids, size_t n, int nbr_level, uint64_t* out, size_t out_stride, uint32_t* count_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !out || out_stride == 0) return set_error(S2B_EINVAL, "null pointer or zero stride");
  DevBuf in, o, cnt;
  cudaError_t e;
  if ((e = in.alloc(8 * n)) != cudaSuccess || (e = o.alloc(8 * n * out_stride)) != cudaSuccess || (e = cnt.alloc(4 * n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(in.p, ids, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellid_all_neighbors_dev((const uint64_t*)in.p, n, nbr_level, (uint64_t*)o.p, out_stride, (uint32_t*)cnt.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(out, o.p, 8 * n * out_stride, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if (count_out && (e = cudaMemcpy(count_out, cnt.p, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cellunion_contains_cells(const uint64_t* sorted_ids, size_t m, const uint64_t* ids, size_t n, uint8_t* out) {
  if (n == 0) return S2B_OK;
  if ((m && !sorted_ids) || !ids || !out) return set_error(S2B_EINVAL, "null pointer");
  DevBuf u, in, o;
  cudaError_t e;
  if ((e = u.alloc(8 * m)) != cudaSuccess || (e = in.alloc(8 * n)) != cudaSuccess || (e = o.alloc(n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if (m && (e = cudaMemcpy(u.p, sorted_ids, 8 * m, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  if ((e = cudaMemcpy(in.p, ids, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellunion_contains_cells_dev((const uint64_t*)u.p, m, (const uint64_t*)in.p, n, (uint8_t*)o.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(out, o.p, n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cellunion_normalize(const uint64_t* ids, size_t m, uint64_t* out, size_t* count_out) {
  if (!count_out) return set_error(S2B_EINVAL, "null count_out");
  *count_out = 0;
  if (m == 0) return S2B_OK;
  if (!ids || !out) return set_error(S2B_EINVAL, "null pointer");
  DevBuf in, o;
  cudaError_t e;
  if ((e = in.alloc(8 * m)) != cudaSuccess || (e = o.alloc(8 * m)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(in.p, ids, 8 * m, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int rc = normalize_dev((const uint64_t*)in.p, m, (uint64_t*)o.p, count_out, nullptr);
  if (rc != S2B_OK) return rc;
  if (*count_out && (e = cudaMemcpy(out, o.p, 8 * *count_out, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}


int s2b_cellid_children_dev(const uint64_t* ids_dev, size_t n, uint64_t* children4_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev || !children4_out_dev) return set_error(S2B_EINVAL, "null pointer");
  cellid_traverse_kernel<<<grid_for_ops(n, (const void*)cellid_traverse_kernel), kThreads, 0, (cudaStream_t)stream>>>(ids_dev, nullptr, nullptr, n, 0, children4_out_dev, nullptr, nullptr);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cellid_children launch");
}

int s2b_cellid_advance_dev(const uint64_t* ids_dev, const int64_t* steps_dev, size_t n, int wrap, uint64_t* out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev || !steps_dev || !out_dev) return set_error(S2B_EINVAL, "null pointer");
  cellid_traverse_kernel<<<grid_for_ops(n, (const void*)cellid_traverse_kernel), kThreads, 0, (cudaStream_t)stream>>>(ids_dev, nullptr, (const long long*)steps_dev, n, wrap, nullptr, out_dev, nullptr);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cellid_advance launch");
}

int s2b_cellid_common_ancestor_level_dev(const uint64_t* a_dev, const uint64_t* b_dev, size_t n, int8_t* level_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!a_dev || !b_dev || !level_out_dev) return set_error(S2B_EINVAL, "null pointer");
  cellid_traverse_kernel<<<grid_for_ops(n, (const void*)cellid_traverse_kernel), kThreads, 0, (cudaStream_t)stream>>>(a_dev, b_dev, nullptr, n, 0, nullptr, nullptr, level_out_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cellid_common_ancestor_level launch");
}

// Host-pointer forms of the three above (whole-array copies).
int s2b_cellid_children(const uint64_t* ids, size_t n, uint64_t* children4_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !children4_out) return set_error(S2B_EINVAL, "null pointer");
  DevBuf in, out;
  cudaError_t e;
  if ((e = in.alloc(8 * n)) != cudaSuccess || (e = out.alloc(32 * n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(in.p, ids, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellid_children_dev((const uint64_t*)in.p, n, (uint64_t*)out.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(children4_out, out.p, 32 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cellid_advance(const uint64_t* ids, const int64_t* steps, size_t n, int wrap, uint64_t* out) {
  if (n == 0) return S2B_OK;
  if (!ids || !steps || !out) return set_error(S2B_EINVAL, "null pointer");
  DevBuf in, st, o;
  cudaError_t e;
  if ((e = in.alloc(8 * n)) != cudaSuccess || (e = st.alloc(8 * n)) != cudaSuccess || (e = o.alloc(8 * n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(in.p, ids, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess || (e = cudaMemcpy(st.p, steps, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess)
    return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellid_advance_dev((const uint64_t*)in.p, (const int64_t*)st.p, n, wrap, (uint64_t*)o.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(out, o.p, 8 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

int s2b_cellid_common_ancestor_level(const uint64_t* a, const uint64_t* b, size_t n, int8_t* level_out) {
  if (n == 0) return S2B_OK;
  if (!a || !b || !level_out) return set_error(S2B_EINVAL, "null pointer");
  DevBuf da, db, o;
  cudaError_t e;
  if ((e = da.alloc(8 * n)) != cudaSuccess || (e = db.alloc(8 * n)) != cudaSuccess || (e = o.alloc(n)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(da.p, a, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess || (e = cudaMemcpy(db.p, b, 8 * n, cudaMemcpyHostToDevice)) != cudaSuccess)
    return cuda_fail(e, "cudaMemcpy");
  int rc = s2b_cellid_common_ancestor_level_dev((const uint64_t*)da.p, (const uint64_t*)db.p, n, (int8_t*)o.p, nullptr);
  if (rc != S2B_OK) return rc;
  if ((e = cudaMemcpy(level_out, o.p, n, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
  return S2B_OK;
}

}  // extern "C"
