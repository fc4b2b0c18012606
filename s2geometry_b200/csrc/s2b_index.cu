// s2b_index.cu — K2/K3/K4: the device-resident flattened shape index and the batched point queries over it.
//
// One thread per point, persistent grid. Per point: K1 (leaf cell id, Hilbert table in shared memory) -> Locate
// (lookup-table bucket + 0-3 probes of the sorted cell ids, all L2/L1 resident) -> for the (few) points that land in
// an index cell, the crossing test of s2b_contains.cuh from the precomputed cell centre over that cell's clipped
// edges, with the Sign chain (triage -> stable -> exact superaccumulator -> symbolic) entirely on the device.
// Index arrays are read through the read-only path; points are streamed (24 B in, 1 B / a few atomics out).
#include <new>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_contains.cuh"
#include "s2b_flat_host.h"
#include "s2b_hilbert.h"
#include "s2b_internal.h"

namespace s2b {

__device__ const HilbertTables g_hilbert_ix = kHilbert;

constexpr int kThreads = 256;

enum QueryMode { kModeAny = 0, kModeShape = 1, kModeHistogram = 2, kModeCount = 3, kModeFill = 4 };

struct QueryOut {
  uint8_t* flags;                 // any / shape
  int shape_id;                   // shape
  unsigned long long* counts;     // histogram
  uint32_t* per_point;            // count
  const uint32_t* offsets;        // fill
  int32_t* shape_ids;             // fill
};

template <int kMode>
__global__ void __launch_bounds__(kThreads, 4)
query_kernel(const FlatIndexView ix, const double* __restrict__ xyz, size_t n, int model, const QueryOut out) {
  __shared__ uint16_t s_pos[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_ix.pos[k];
  __syncthreads();
  // Warp-uniform trip count + __syncwarp() per iteration: the few lanes that land in an index cell (and the rarer
  // ones in the exact stage) rejoin their warp before the next batch of 32 points instead of drifting apart for good.
  const size_t stride = (size_t)gridDim.x * kThreads;
  const unsigned lane = threadIdx.x & 31u;
  for (size_t base = (size_t)blockIdx.x * kThreads + (threadIdx.x & ~31u); base < n; base += stride, __syncwarp()) {
    const size_t i = base + lane;
    if (i >= n) continue;
    const double* q = xyz + 3 * i;
    P3 p{__ldcs(q), __ldcs(q + 1), __ldcs(q + 2)};
    if (kMode == kModeAny) {
      bool any = false;
      visit_containing_shapes(ix, p, s_pos, model, [&](int) { any = true; return false; });
      out.flags[i] = any ? 1 : 0;
    } else if (kMode == kModeShape) {
      // ShapeContains(shape_id, p): only the clip of that shape is evaluated (find_clipped).
      bool hit = false;
      uint64_t leaf = cellid_from_point(p, s_pos);
      int cell = locate_cell(ix, leaf);
      if (cell >= 0) {
        FlatCell fc = ix.cells[cell];
        for (uint32_t k = 0; k < fc.clip_count; ++k) {
          FlatClip clip = ix.clips[fc.clip_begin + k];
          if (clip.shape_id == out.shape_id) {
            P3 a{fc.cx, fc.cy, fc.cz};
            hit = clip_contains(ix, clip, a, p, cross(a, p), model);
            break;
          }
        }
      }
      out.flags[i] = hit ? 1 : 0;
    } else if (kMode == kModeHistogram) {
      visit_containing_shapes(ix, p, s_pos, model, [&](int id) { atomicAdd(out.counts + id, 1ull); return true; });
    } else if (kMode == kModeCount) {
      uint32_t c = 0;
      visit_containing_shapes(ix, p, s_pos, model, [&](int) { ++c; return true; });
      out.per_point[i] = c;
    } else {
      uint32_t w = out.offsets[i];
      visit_containing_shapes(ix, p, s_pos, model, [&](int id) { out.shape_ids[w++] = id; return true; });
    }
  }
}

// S2CellUnion::Contains(S2Point): leaf id + binary search over the (<= a few thousand) sorted ids, staged in shared
// memory when they fit.
constexpr int kUnionSmemIds = 4096;
__global__ void __launch_bounds__(kThreads)
cellunion_kernel(const uint64_t* __restrict__ ids, uint32_t m, const double* __restrict__ xyz, size_t n, uint8_t* __restrict__ out) {
  __shared__ uint16_t s_pos[1024];
  __shared__ uint64_t s_ids[kUnionSmemIds];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_ix.pos[k];
  const bool in_smem = m <= kUnionSmemIds;
  if (in_smem) for (uint32_t k = threadIdx.x; k < m; k += kThreads) s_ids[k] = ids[k];
  __syncthreads();
  const uint64_t* tab = in_smem ? s_ids : ids;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const unsigned lane = threadIdx.x & 31u;
  for (size_t base = (size_t)blockIdx.x * kThreads + (threadIdx.x & ~31u); base < n; base += stride, __syncwarp()) {
    const size_t i = base + lane;
    if (i >= n) continue;
    const double* q = xyz + 3 * i;
    P3 p{__ldcs(q), __ldcs(q + 1), __ldcs(q + 2)};
    uint64_t leaf = cellid_from_point(p, s_pos);
    // lower_bound with EntirelyPrecedes: first cell whose range_max >= leaf (s2cell_union.cc:285-300)
    uint32_t lo = 0, hi = m;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (cellid_range_max(tab[mid]) < leaf) lo = mid + 1; else hi = mid;
    }
    out[i] = (lo < m && cellid_contains(tab[lo], leaf)) ? 1 : 0;
  }
}

static int grid_for_ix(size_t n, const void* kernel) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  size_t want = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

template <int kMode>
static cudaError_t launch_query_t(const FlatIndexView& ix, const double* xyz, size_t n, int model, const QueryOut& out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  auto kern = query_kernel<kMode>;
  kern<<<grid_for_ix(n, (const void*)kern), kThreads, 0, st>>>(ix, xyz, n, model, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace s2b

using namespace s2b;

// ---------------------------------------------------------------------------------------------------
struct S2bIndex {
  int device = 0;
  FlatIndexView view{};
  void* blob = nullptr;   // one allocation holding every array
  size_t blob_bytes = 0;
  uint64_t num_clips = 0, num_edges = 0;
};

namespace s2b {
int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
}

static int check_model(int m) {
  if (m < 0 || m > 2) return set_error(S2B_EINVAL, "vertex_model must be 0 (OPEN), 1 (SEMI_OPEN) or 2 (CLOSED)");
  return S2B_OK;
}

extern "C" {

int s2b_index_create(const S2bFlatIndex* f, S2bIndex** out) {
  if (!f || !out) return set_error(S2B_EINVAL, "null argument");
  *out = nullptr;
  HostFlat hf;
  std::string err;
  int prc = pack_flat_index(*f, &hf, &err);
  if (prc != S2B_OK) return set_error(prc, "%s", err.c_str());
  const uint64_t C = f->num_cells, K = f->num_clips, E = f->num_edges, nb = hf.num_buckets;
  const int lut_shift = hf.lut_shift;
  std::vector<FlatCell>& cells = hf.cells;
  std::vector<FlatClip>& clips = hf.clips;
  std::vector<FlatEdge>& edges = hf.edges;
  std::vector<uint32_t>& lut = hf.lut;

  S2bIndex* ix = new (std::nothrow) S2bIndex();
  if (!ix) return set_error(S2B_ENOMEM, "host allocation failed");
  cudaError_t e = cudaGetDevice(&ix->device);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaGetDevice"); }
  auto align = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o_ids = 0, o_cells = o_ids + align(8 * (C + 1)), o_clips = o_cells + align(sizeof(FlatCell) * (C + 1)),
         o_edges = o_clips + align(sizeof(FlatClip) * (K + 1)), o_lut = o_edges + align(sizeof(FlatEdge) * (E + 1)),
         total = o_lut + align(4 * (nb + 2));
  e = cudaMalloc(&ix->blob, total);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc(index)"); }
  ix->blob_bytes = total;
  char* base = (char*)ix->blob;
  bool ok = true;
  ok &= C == 0 || cudaMemcpy(base + o_ids, f->cell_ids, 8 * C, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= C == 0 || cudaMemcpy(base + o_cells, cells.data(), sizeof(FlatCell) * C, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= K == 0 || cudaMemcpy(base + o_clips, clips.data(), sizeof(FlatClip) * K, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= E == 0 || cudaMemcpy(base + o_edges, edges.data(), sizeof(FlatEdge) * E, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(base + o_lut, lut.data(), 4 * (nb + 2), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { e = cudaGetLastError(); cudaFree(ix->blob); delete ix; return cuda_fail(e, "cudaMemcpy(index)"); }
  ix->view.cell_ids = (const uint64_t*)(base + o_ids);
  ix->view.cells = (const FlatCell*)(base + o_cells);
  ix->view.clips = (const FlatClip*)(base + o_clips);
  ix->view.edges = (const FlatEdge*)(base + o_edges);
  ix->view.lut = (const uint32_t*)(base + o_lut);
  ix->view.num_cells = (uint32_t)C;
  ix->view.lut_shift = lut_shift;
  ix->view.num_shape_ids = f->num_shape_ids;
  ix->num_clips = K; ix->num_edges = E;
  *out = ix;
  return S2B_OK;
}

void s2b_index_destroy(S2bIndex* ix) {
  if (!ix) return;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != ix->device) cudaSetDevice(ix->device);
  if (ix->blob) cudaFree(ix->blob);
  if (cur != ix->device) cudaSetDevice(cur);
  delete ix;
}

int s2b_index_info(const S2bIndex* ix, uint64_t* cells, uint64_t* clips, uint64_t* edges, int32_t* shapes, uint64_t* bytes) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (cells) *cells = ix->view.num_cells;
  if (clips) *clips = ix->num_clips;
  if (edges) *edges = ix->num_edges;
  if (shapes) *shapes = ix->view.num_shape_ids;
  if (bytes) *bytes = ix->blob_bytes;
  return S2B_OK;
}

#define S2B_QUERY_PROLOGUE                                                            \
  if (!ix) return set_error(S2B_EINVAL, "null index");                                \
  { int rc__ = check_model(model); if (rc__) return rc__; }                           \
  if (n == 0) return S2B_OK;

int s2b_contains_dev(const S2bIndex* ix, const double* xyz, size_t n, int model, uint8_t* out, void* stream) {
  S2B_QUERY_PROLOGUE
  if (!xyz || !out) return set_error(S2B_EINVAL, "null pointer");
  QueryOut o{}; o.flags = out;
  cudaError_t e = launch_query_t<kModeAny>(ix->view, xyz, n, model, o, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "contains launch");
}

int s2b_shape_contains_dev(const S2bIndex* ix, int shape_id, const double* xyz, size_t n, int model, uint8_t* out, void* stream) {
  S2B_QUERY_PROLOGUE
  if (!xyz || !out) return set_error(S2B_EINVAL, "null pointer");
  QueryOut o{}; o.flags = out; o.shape_id = shape_id;
  cudaError_t e = launch_query_t<kModeShape>(ix->view, xyz, n, model, o, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "shape_contains launch");
}

int s2b_shape_histogram_dev(const S2bIndex* ix, const double* xyz, size_t n, int model, uint64_t* counts, void* stream) {
  S2B_QUERY_PROLOGUE
  if (!xyz || !counts) return set_error(S2B_EINVAL, "null pointer");
  QueryOut o{}; o.counts = (unsigned long long*)counts;
  cudaError_t e = launch_query_t<kModeHistogram>(ix->view, xyz, n, model, o, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "histogram launch");
}

int s2b_containing_shapes_count_dev(const S2bIndex* ix, const double* xyz, size_t n, int model, uint32_t* counts, void* stream) {
  S2B_QUERY_PROLOGUE
  if (!xyz || !counts) return set_error(S2B_EINVAL, "null pointer");
  QueryOut o{}; o.per_point = counts;
  cudaError_t e = launch_query_t<kModeCount>(ix->view, xyz, n, model, o, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "count launch");
}

int s2b_containing_shapes_fill_dev(const S2bIndex* ix, const double* xyz, size_t n, int model, const uint32_t* offsets,
                                   int32_t* shape_ids, void* stream) {
  S2B_QUERY_PROLOGUE
  if (!xyz || !offsets || !shape_ids) return set_error(S2B_EINVAL, "null pointer");
  QueryOut o{}; o.offsets = offsets; o.shape_ids = shape_ids;
  cudaError_t e = launch_query_t<kModeFill>(ix->view, xyz, n, model, o, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "fill launch");
}

int s2b_exact_stage_count(uint64_t* count_out, int reset) {
  unsigned long long v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_exact_count, sizeof v);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyFromSymbol");
  if (count_out) *count_out = v;
  if (reset) {
    v = 0;
    e = cudaMemcpyToSymbol(g_exact_count, &v, sizeof v);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyToSymbol");
  }
  return S2B_OK;
}

int s2b_cellunion_contains_dev(const uint64_t* ids, size_t m, const double* xyz, size_t n, uint8_t* out, void* stream) {
  if (n == 0) return S2B_OK;
  if ((m && !ids) || !xyz || !out) return set_error(S2B_EINVAL, "null pointer");
  if (m > 0xFFFFFFF0ull) return set_error(S2B_EINVAL, "cell union too large");
  cellunion_kernel<<<grid_for_ix(n, (const void*)cellunion_kernel), kThreads, 0, (cudaStream_t)stream>>>(ids, (uint32_t)m, xyz, n, out);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cellunion launch");
}

}  // extern "C"
