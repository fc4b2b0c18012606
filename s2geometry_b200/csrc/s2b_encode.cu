// s2b_encode.cu — K1: batched point -> S2CellId encoding for sm_100a.
//
// One thread per point, persistent grid (SM count x resident CTAs), grid-stride loop. The 2 KB Hilbert
// lookup_pos table (s2cell_id.cc:75-115) and, for lat/lng input, the 3.3 KB sin/cos node table live in shared
// memory. Input is streamed with cache-streaming loads (read once), ids are written coalesced with
// streaming stores. All FP64 arithmetic is IEEE (div.rn/sqrt.rn, -fmad=false) so ids equal the reference's bit for bit.
//
// (A cp.async shared-memory staged variant of the xyz kernel was measured and rejected: 96 vs 112 Gpts/s — this kernel
// is bound by instruction issue, not by DRAM latency; the staged design pays off only in the ~150-instruction Locate
// pre-filter of s2b_index.cu.)
//
// Algorithmic traffic per point: xyz 24 B in + 8 B per level out; lat/lng 16 B (E7: 8 B) in + 8 B per level out.
#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_internal.h"
#include "s2b_sincos.cuh"

namespace s2b {

__device__ const HilbertTables g_hilbert = kHilbert;
__device__ uint16_t g_hilbert_wide[16384];   // uploaded on first use per device (ensure_wide_table)

static cudaError_t ensure_wide_table() {
  static thread_local int done_for_device = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev == done_for_device) return e;
  e = cudaMemcpyToSymbol(g_hilbert_wide, wide_hilbert_host(), 16384 * sizeof(uint16_t));
  if (e == cudaSuccess) done_for_device = dev;
  return e;
}
__device__ const double g_sincos_tab[sc_tab::kTableSize][4] = {
#define S2B_ROW(i) {sc_tab::kSinCosTable[i][0], sc_tab::kSinCosTable[i][1], sc_tab::kSinCosTable[i][2], sc_tab::kSinCosTable[i][3]}
#define S2B_ROW5(i) S2B_ROW(i), S2B_ROW(i + 1), S2B_ROW(i + 2), S2B_ROW(i + 3), S2B_ROW(i + 4)
#define S2B_ROW25(i) S2B_ROW5(i), S2B_ROW5(i + 5), S2B_ROW5(i + 10), S2B_ROW5(i + 15), S2B_ROW5(i + 20)
    S2B_ROW25(0), S2B_ROW25(25), S2B_ROW25(50), S2B_ROW25(75), S2B_ROW5(100)};
static_assert(sc_tab::kTableSize == 105, "table rows are spelled out above");

constexpr int kThreads = 256;

// levels live in kernel parameter space; a switch keeps them out of local memory (dynamic indexing would copy them).
__device__ __forceinline__ int level_at(const LevelSpec& l, int k) {
  return k == 0 ? l.level[0] : (k == 1 ? l.level[1] : (k == 2 ? l.level[2] : l.level[3]));
}

// Epilogue shared by the encode kernels: parent() at each requested level and, optionally, one token — in the same
// pass, so ids are never re-read.
__device__ __forceinline__ void emit_ids_and_token(uint64_t id, size_t i, size_t n, const EncodeOut& o) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < o.levels.n) {
      // parent(level): (id & -lsb) | lsb; for level 30 lsb = 1 and the id (odd) is unchanged (s2cell_id.h:662-668)
      const uint64_t lsb = lsb_for_level(level_at(o.levels, k));
      __stcs(o.ids + (size_t)k * n + i, (id & (~lsb + 1)) | lsb);
    }
  }
  if (o.token_level_index >= 0) {
    int L = level_at(o.levels, o.token_level_index);
    uint32_t w[4];
    int len = cellid_to_token(L < 30 ? cellid_parent(id, L) : id, w);
    __stcs((uint4*)o.tokens16 + i, make_uint4(w[0], w[1], w[2], w[3]));
    if (o.token_len) o.token_len[i] = (uint8_t)len;
  }
}

template <int kKind>
__global__ void __launch_bounds__(kThreads, 6)
encode_kernel(const void* __restrict__ in, size_t n, const EncodeOut o) {
  __shared__ uint16_t s_pos[16384];   // 6-bit Hilbert table: 5 dependent lookups per point
  __shared__ double s_sc[kKind == kInXYZ ? 1 : sc_tab::kTableSize][4];
  for (int k = threadIdx.x; k < 16384 / 8; k += kThreads) ((uint4*)s_pos)[k] = ((const uint4*)g_hilbert_wide)[k];
  if (kKind != kInXYZ) {
    for (int k = threadIdx.x; k < sc_tab::kTableSize * 4; k += kThreads) (&s_sc[0][0])[k] = (&g_sincos_tab[0][0])[k];
  }
  __syncthreads();

  // Warp-uniform trip count + __syncwarp() at the end of every iteration: lanes that took a rare path (double-double
  // sin/cos, division/sqrt slow paths) rejoin their warp instead of running the rest of the loop on their own.
  const size_t stride = (size_t)gridDim.x * kThreads;
  const unsigned lane = threadIdx.x & 31u;
  for (size_t base = (size_t)blockIdx.x * kThreads + (threadIdx.x & ~31u); base < n; base += stride, __syncwarp()) {
    const size_t i = base + lane;
    if (i >= n) continue;
    P3 p;
    double fi, fj;
    int face;
    bool sensitive = false;
    if (kKind == kInXYZ) {
      const double* q = (const double*)in + 3 * i;
      p.x = __ldcs(q); p.y = __ldcs(q + 1); p.z = __ldcs(q + 2);
    } else {
      double lat, lng;
      if (kKind == kInLatLngRad) {
        double2 ll = __ldcs((const double2*)in + i);
        lat = ll.x; lng = ll.y;
      } else {
        int2 e7 = __ldcs((const int2*)in + i);
        // S1Angle::E7 = Degrees(1e-7 * e7) = (M_PI / 180) * (1e-7 * e7)   (s1angle.h:363-377)
        lat = (M_PI / 180) * (1e-7 * (double)e7.x);
        lng = (M_PI / 180) * (1e-7 * (double)e7.y);
      }
      face = latlng_to_face_fij(lat, lng, s_sc, o.force_cr != 0, &p, &fi, &fj, &sensitive);
    }
    if (kKind == kInXYZ) face = point_to_face_fij(p, &fi, &fj);
    const uint64_t id = from_face_ij_wide(face, fij_to_ij(fi), fij_to_ij(fj), s_pos);
    emit_ids_and_token(id, i, n, o);
    if (o.flags) o.flags[i] = (kKind == kInXYZ || sensitive) && near_cell_boundary(fi, fj, kFlagTolIj) ? 1 : 0;
  }
}

static int grid_for(size_t n, const void* kernel) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  size_t want = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

template <int kKind>
static cudaError_t launch_encode_k(const void* in, size_t n, const EncodeOut& o, cudaStream_t st) {
  auto kern = encode_kernel<kKind>;
  kern<<<grid_for(n, (const void*)kern), kThreads, 0, st>>>(in, n, o);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_encode(InputKind kind, const void* in, size_t n, const EncodeOut& o, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  cudaError_t te = ensure_wide_table();
  if (te != cudaSuccess) return te;
  switch (kind) {
    case kInXYZ: return launch_encode_k<kInXYZ>(in, n, o, st);
    case kInLatLngRad: return launch_encode_k<kInLatLngRad>(in, n, o, st);
    default: return launch_encode_k<kInLatLngE7>(in, n, o, st);
  }
}

// ---------------------------------------------------------------------------------------------------
// S2CellId::parent (s2cell_id.h:662-668) and ToToken (s2cell_id.cc:217-233): pure integer, HBM-bound.
__global__ void __launch_bounds__(kThreads) parent_kernel(const uint64_t* __restrict__ ids, size_t n, int level, uint64_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) __stcs(out + i, cellid_parent(__ldcs(ids + i), level));
}

__global__ void __launch_bounds__(kThreads) token_kernel(const uint64_t* __restrict__ ids, size_t n, uint4* __restrict__ tok, uint8_t* __restrict__ len) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    uint32_t w[4];
    int l = cellid_to_token(__ldcs(ids + i), w);
    __stcs(tok + i, make_uint4(w[0], w[1], w[2], w[3]));
    if (len) len[i] = (uint8_t)l;
  }
}

// S2CellId::ToPoint (s2cell_id.h:184, :555-581; s2coords.cc:68-72): inverse Hilbert table (2 KB) in shared memory.
__global__ void __launch_bounds__(kThreads) to_point_kernel(const uint64_t* __restrict__ ids, size_t n, double* __restrict__ xyz) {
  __shared__ uint16_t s_ij[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_ij[k] = g_hilbert.ij[k];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const P3 p = cellid_to_point(__ldcs(ids + i), s_ij);
    double* q = xyz + 3 * i;
    __stcs(q, p.x); __stcs(q + 1, p.y); __stcs(q + 2, p.z);
  }
}

// S2CellId::GetEdgeNeighbors (s2cell_id.cc:499-512): out[4*i + k], k = down, right, up, left.
__global__ void __launch_bounds__(kThreads) edge_neighbors_kernel(const uint64_t* __restrict__ ids, size_t n, uint64_t* __restrict__ out) {
  __shared__ uint16_t s_pos[1024], s_ij[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) { s_pos[k] = g_hilbert.pos[k]; s_ij[k] = g_hilbert.ij[k]; }
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    uint64_t nb[4];
    cellid_edge_neighbors(__ldcs(ids + i), s_pos, s_ij, nb);
    ulonglong2* o = (ulonglong2*)(out + 4 * i);
    __stcs(o, make_ulonglong2(nb[0], nb[1]));
    __stcs(o + 1, make_ulonglong2(nb[2], nb[3]));
  }
}

cudaError_t launch_edge_neighbors(const uint64_t* ids, size_t n, uint64_t* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  edge_neighbors_kernel<<<grid_for(n, (const void*)edge_neighbors_kernel), kThreads, 0, st>>>(ids, n, out);
  count_launch();
  return cudaGetLastError();
}

// S2CellId::FromToken (s2cell_id.cc:235-254).
__global__ void __launch_bounds__(kThreads) from_token_kernel(const uint4* __restrict__ tok, const uint8_t* __restrict__ len, size_t n, uint64_t* __restrict__ ids) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const uint4 t = __ldcs(tok + i);
    unsigned char b[16];
    *(uint4*)b = t;
    __stcs(ids + i, cellid_from_token(b, len[i]));
  }
}

cudaError_t launch_to_point(const uint64_t* ids, size_t n, double* xyz, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  to_point_kernel<<<grid_for(n, (const void*)to_point_kernel), kThreads, 0, st>>>(ids, n, xyz);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_from_token(const char* tok, const uint8_t* len, size_t n, uint64_t* ids, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  from_token_kernel<<<grid_for(n, (const void*)from_token_kernel), kThreads, 0, st>>>((const uint4*)tok, len, n, ids);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_parent(const uint64_t* ids, size_t n, int level, uint64_t* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  parent_kernel<<<grid_for(n, (const void*)parent_kernel), kThreads, 0, st>>>(ids, n, level, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_to_token(const uint64_t* ids, size_t n, char* tok, uint8_t* len, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  token_kernel<<<grid_for(n, (const void*)token_kernel), kThreads, 0, st>>>(ids, n, (uint4*)tok, len);
  count_launch();
  return cudaGetLastError();
}

}  // namespace s2b
