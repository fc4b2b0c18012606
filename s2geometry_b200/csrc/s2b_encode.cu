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
#include <atomic>

#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_internal.h"
#include "s2b_sincos.cuh"

namespace s2b {

__device__ const HilbertTables g_hilbert = kHilbert;
__device__ uint16_t g_hilbert_wide[16384];   // uploaded on first use per device (ensure_wide_table)
__device__ double g_sincos512[512][2];       // sin, cos of k*pi/256 (uploaded with it)

static cudaError_t ensure_wide_table() {
  static std::atomic<uint64_t> done_mask{0};   // bit d: device d holds the table (one host thread may drive several GPUs)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || (dev < 64 && ((done_mask.load(std::memory_order_acquire) >> dev) & 1))) return e;
  e = cudaMemcpyToSymbol(g_hilbert_wide, wide_hilbert_host(), 16384 * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_sincos512, sc_tab::kSinCos512, sizeof(sc_tab::kSinCos512));
  if (e == cudaSuccess && dev < 64) done_mask.fetch_or(1ull << dev, std::memory_order_release);
  return e;
}
__device__ const double g_sincos_tab[sc_tab::kTableSize][4] = {
#define S2B_ROW(i) {sc_tab::kSinCosTable[i][0], sc_tab::kSinCosTable[i][1], sc_tab::kSinCosTable[i][2], sc_tab::kSinCosTable[i][3]}
#define S2B_ROW5(i) S2B_ROW(i), S2B_ROW(i + 1), S2B_ROW(i + 2), S2B_ROW(i + 3), S2B_ROW(i + 4)
#define S2B_ROW25(i) S2B_ROW5(i), S2B_ROW5(i + 5), S2B_ROW5(i + 10), S2B_ROW5(i + 15), S2B_ROW5(i + 20)
    S2B_ROW25(0), S2B_ROW25(25), S2B_ROW25(50), S2B_ROW25(75), S2B_ROW5(100)};
static_assert(sc_tab::kTableSize == 105, "table rows are spelled out above");

constexpr int kThreads = 256;
// Resident CTAs per SM of the encode kernels. With the input prefetch 5 x 8 warps (48 registers, no spills) measured
// 81 Gpts/s on the fused lat/lng step, 6 (40 registers, two spilled values) 79, 4 (64 registers) 81.
#ifndef S2B_ENCODE_CTAS
#define S2B_ENCODE_CTAS 5
#endif
#ifndef S2B_PREFETCH_DIST
#define S2B_PREFETCH_DIST 1   // iterations beyond the register prefetch that the L2 prefetch runs ahead
#endif
constexpr int kEncodeCtas = S2B_ENCODE_CTAS;

// What one launch emits, with everything that is the same for all points worked out on the host: per-level output
// bases and parent() masks (parent(level) = (id & -lsb) | lsb, s2cell_id.h:662-668; for level 30 the masks are ~0 and 1),
// and for the token its length and the byte mask of the 16-byte slot — a level-L id has exactly (30 - L) / 2 trailing
// zero hex digits (ToToken strips them, s2cell_id.cc:217-233), and encoded ids are never 0.
struct EncodeArgs {
  uint64_t* ids[4];
  uint64_t and_mask[4], or_mask[4];
  uint8_t* flags;          // nullable
  uint4* tokens16;         // token slots (kToken kernels only)
  uint8_t* token_len;      // nullable
  uint64_t tok_and, tok_or;
  uint32_t tok_keep[4];
  uint32_t tok_len;
  int force_cr;
  FixRecord* fix_list;     // nullable: strict host-libm mode (s2b_internal.h)
  uint32_t* fix_count;
  uint32_t fix_cap, fix_base;
  uint32_t* fix_overflow;
  uint64_t index_base;
};

template <int kKind> struct RawPoint;
template <> struct RawPoint<kInXYZ> { double x, y, z; };
template <> struct RawPoint<kInLatLngRad> { double2 ll; };
template <> struct RawPoint<kInLatLngE7> { int2 e7; };

__device__ __forceinline__ RawPoint<kInXYZ> load_raw(const void* in, uint32_t i, RawPoint<kInXYZ>*) {
  const double* q = (const double*)in + 3 * (size_t)i;
  return {__ldcs(q), __ldcs(q + 1), __ldcs(q + 2)};
}
__device__ __forceinline__ RawPoint<kInLatLngRad> load_raw(const void* in, uint32_t i, RawPoint<kInLatLngRad>*) { return {__ldcs((const double2*)in + i)}; }
__device__ __forceinline__ RawPoint<kInLatLngE7> load_raw(const void* in, uint32_t i, RawPoint<kInLatLngE7>*) { return {__ldcs((const int2*)in + i)}; }

// Strict host-libm mode, rare path (out of line so that it costs the main loop no registers).
__device__ __noinline__ void append_fix(FixRecord* list, uint32_t* count, uint32_t cap, uint32_t base, uint32_t* overflow, uint64_t index, double lat, double lng, uint64_t id) {
  const uint32_t slot = atomicAdd(count, 1u) - base;   // unsigned arithmetic: correct across a wrap of the counter
  if (slot < cap) {
    // The list may live in mapped host memory that the host reads WHILE this kernel runs (s2b_capi.cu, LiveList): the id
    // (never 0) is the "record complete" marker and is written last, behind a system-scope fence.
    volatile unsigned long long* r = (volatile unsigned long long*)(list + slot);
    r[0] = index;
    r[1] = (unsigned long long)__double_as_longlong(lat);
    r[2] = (unsigned long long)__double_as_longlong(lng);
    __threadfence_system();
    r[3] = id;
  } else if (overflow) {
    *(volatile uint32_t*)overflow = 1u;   // the host falls back to a piecewise re-scan (s2b_capi.cu)
  }
}

template <class Raw>
__device__ __forceinline__ void prefetch_l2(const void* in, uint32_t i, Raw*) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)in + (size_t)i * sizeof(Raw)));
}

// One thread per point, persistent grid. The input of the NEXT iteration is loaded before the current point is
// processed, so a warp does not sit on the long scoreboard at the top of every iteration.
// (Handing out work dynamically — 256-point chunks per warp from a global cursor — was measured and made no
// difference: 80.6 vs 81.2 Gpts/s. The kernel is bound by instruction issue: 316 instructions per point, 0.71 IPC.)
// n < 2^32 per launch (the launcher splits larger batches), so indices are 32-bit and every address is one IMAD.WIDE.
template <int kKind, int kLevels, bool kToken>
__global__ void __launch_bounds__(kThreads, kEncodeCtas)
encode_kernel(const void* __restrict__ in, uint32_t n, const EncodeArgs a) {
  // 6-bit Hilbert table (5 dependent lookups per point), repacked for the lookup chain (from_face_ij_chain)
  __shared__ __align__(16) uint16_t s_pos[16384];
  __shared__ double s_sc[kKind == kInXYZ ? 1 : sc_tab::kTableSize][4];       // correctly rounded path (rare)
  __shared__ __align__(16) double s_sc512[kKind == kInXYZ ? 1 : 512][2];     // the filter's sin/cos table
  for (int k = threadIdx.x; k < 16384 / 2; k += kThreads) {
    const uint32_t w = ((const uint32_t*)g_hilbert_wide)[k];
    ((uint32_t*)s_pos)[k] = (uint32_t)hilbert_chain_repack((uint16_t)w) | ((uint32_t)hilbert_chain_repack((uint16_t)(w >> 16)) << 16);
  }
  if (kKind != kInXYZ) {
    for (int k = threadIdx.x; k < sc_tab::kTableSize * 4; k += kThreads) (&s_sc[0][0])[k] = (&g_sincos_tab[0][0])[k];
    for (int k = threadIdx.x; k < 512; k += kThreads) ((double2*)s_sc512)[k] = ((const double2*)g_sincos512)[k];
  }
  __syncthreads();

  // Warp-uniform trip count + __syncwarp() at the end of every iteration: lanes that took a rare path (double-double
  // sin/cos, division/sqrt slow paths) rejoin their warp instead of running the rest of the loop on their own.
  using Raw = RawPoint<kKind>;
  uint32_t i = blockIdx.x * kThreads + threadIdx.x;
  Raw cur{};
  if (i < n) cur = load_raw(in, i, (Raw*)nullptr);
  // (i & ~31) is the warp's first index (the stride is a multiple of 256), so the trip count is warp-uniform without
  // keeping the lane id around; the grid size is re-read from its constant-bank slot every iteration instead of living
  // in (and being spilled from) a register.
  while ((i & ~31u) < n) {
    uint32_t nctas;
    asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(nctas));
    const uint32_t stride = nctas * kThreads;
    const uint32_t i_next = i + stride;    // n <= 2^31 and stride < 2^20: no wrap-around
    Raw nxt{};
    if (i_next < n) nxt = load_raw(in, i_next, (Raw*)nullptr);
    // the iteration after next: pull its line into the L2 now, so that the register prefetch above finds it there
    if (i_next + S2B_PREFETCH_DIST * stride < n) prefetch_l2(in, i_next + S2B_PREFETCH_DIST * stride, (Raw*)nullptr);
    if (i < n) {
      P3 p;
      double fi, fj;
      int face;
      bool sensitive = false;
      if constexpr (kKind == kInXYZ) {
        p = P3{cur.x, cur.y, cur.z};
        face = a.force_cr ? point_to_face_fij(p, &fi, &fj) : point_to_face_fij_checked(p, &fi, &fj);
      } else {
        double lat, lng;
        if constexpr (kKind == kInLatLngRad) {
          lat = cur.ll.x; lng = cur.ll.y;
        } else {
          // S1Angle::E7 = Degrees(1e-7 * e7) = (M_PI / 180) * (1e-7 * e7)   (s1angle.h:363-377)
          lat = (M_PI / 180) * (1e-7 * (double)cur.e7.x);
          lng = (M_PI / 180) * (1e-7 * (double)cur.e7.y);
        }
        face = latlng_to_face_fij(lat, lng, s_sc, s_sc512, a.force_cr != 0, &p, &fi, &fj, &sensitive);
      }
      const uint64_t id = from_face_ij_chain(face, fij_to_ij(fi), fij_to_ij(fj), (const unsigned char*)s_pos);
#pragma unroll
      for (int k = 0; k < kLevels; ++k) __stcs(a.ids[k] + i, (id & a.and_mask[k]) | a.or_mask[k]);
      if constexpr (kToken) {
        const uint64_t tid = (id & a.tok_and) | a.tok_or;
        uint32_t w[4];
        hex8((uint32_t)(tid >> 32), &w[0], &w[1]);
        hex8((uint32_t)tid, &w[2], &w[3]);
        __stcs(a.tokens16 + i, make_uint4(w[0] & a.tok_keep[0], w[1] & a.tok_keep[1], w[2] & a.tok_keep[2], w[3] & a.tok_keep[3]));
        if (a.token_len) a.token_len[i] = (uint8_t)a.tok_len;
      }
      if (a.flags) a.flags[i] = (kKind == kInXYZ || sensitive) && near_cell_boundary(fi, fj, kFlagTolIj) ? 1 : 0;
      if constexpr (kKind != kInXYZ) {
        // Strict mode: hand every point whose leaf cell a <= 1 ulp change of sin/cos could move (and every angle the
        // device routines do not cover) to the host, which re-evaluates it with its own libm. ~1.6e-5 of uniform input.
        if (sensitive && a.fix_list != nullptr) {
          // the input is read again here (rare) rather than kept live in registers across the whole iteration
          const Raw again = load_raw(in, i, (Raw*)nullptr);
          double lat, lng;
          if constexpr (kKind == kInLatLngRad) {
            lat = again.ll.x; lng = again.ll.y;
          } else {
            lat = (M_PI / 180) * (1e-7 * (double)again.e7.x);
            lng = (M_PI / 180) * (1e-7 * (double)again.e7.y);
          }
          if (near_cell_boundary(fi, fj, kFlagTolIj) || !(fabs(lat) < 1048576.0 && fabs(lng) < 1048576.0))
            append_fix(a.fix_list, a.fix_count, a.fix_cap, a.fix_base, a.fix_overflow, a.index_base + i, lat, lng, id);
        }
      }
    }
    cur = nxt;
    i = i_next;
    __syncwarp();
  }
}

// Resident CTAs per SM of a kernel on the current device, looked up once per (thread, kernel, device): the occupancy query
// costs several microseconds, which a synchronous 1 ms call should not pay every time.
static int resident_ctas(const void* kernel) {
  struct Entry { const void* kernel; int device; int per_sm; };
  static thread_local Entry cache[64];
  static thread_local int used = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int k = 0; k < used; ++k)
    if (cache[k].kernel == kernel && cache[k].device == dev) return cache[k].per_sm;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (used < 64) cache[used++] = Entry{kernel, dev, per_sm};
  return per_sm;
}

static int grid_for(size_t n, const void* kernel) {
  const int per_sm = resident_ctas(kernel);
  size_t want = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

template <int kKind, int kLevels, bool kToken>
static cudaError_t launch_encode_k(const void* in, uint32_t n, const EncodeArgs& a, cudaStream_t st) {
  auto kern = encode_kernel<kKind, kLevels, kToken>;
  kern<<<grid_for(n, (const void*)kern), kThreads, 0, st>>>(in, n, a);
  count_launch();
  return cudaGetLastError();
}

template <int kKind>
static cudaError_t launch_encode_kind(const void* in, uint32_t n, int levels, bool token, const EncodeArgs& a, cudaStream_t st) {
  switch (levels * 2 + (token ? 1 : 0)) {
    case 2: return launch_encode_k<kKind, 1, false>(in, n, a, st);
    case 3: return launch_encode_k<kKind, 1, true>(in, n, a, st);
    case 4: return launch_encode_k<kKind, 2, false>(in, n, a, st);
    case 5: return launch_encode_k<kKind, 2, true>(in, n, a, st);
    case 6: return launch_encode_k<kKind, 3, false>(in, n, a, st);
    case 7: return launch_encode_k<kKind, 3, true>(in, n, a, st);
    case 8: return launch_encode_k<kKind, 4, false>(in, n, a, st);
    case 9: return launch_encode_k<kKind, 4, true>(in, n, a, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_encode(InputKind kind, const void* in, size_t n, const EncodeOut& o, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (o.levels.n < 1 || o.levels.n > 4) return cudaErrorInvalidValue;
  cudaError_t te = ensure_wide_table();
  if (te != cudaSuccess) return te;
  const bool token = o.token_level_index >= 0;
  const size_t in_stride = kind == kInXYZ ? 24 : (kind == kInLatLngRad ? 16 : 8);
  const size_t ids_stride = o.ids_stride ? o.ids_stride : n;
  constexpr size_t kMaxLaunchPoints = (size_t)1 << 31;
  for (size_t off = 0; off < n; off += kMaxLaunchPoints) {
    const uint32_t cnt = (uint32_t)(n - off < kMaxLaunchPoints ? n - off : kMaxLaunchPoints);
    EncodeArgs a{};
    for (int k = 0; k < o.levels.n; ++k) {
      const uint64_t lsb = lsb_for_level(o.levels.level[k]);
      a.ids[k] = o.ids + (size_t)k * ids_stride + off;
      a.and_mask[k] = ~lsb + 1;
      a.or_mask[k] = lsb;
    }
    a.flags = o.flags ? o.flags + off : nullptr;
    a.force_cr = o.force_cr;
    a.fix_list = kind == kInXYZ ? nullptr : o.fix_list;
    a.fix_count = o.fix_count;
    a.fix_cap = o.fix_cap;
    a.fix_base = o.fix_base;
    a.fix_overflow = o.fix_overflow;
    a.index_base = o.index_base + off;
    if (token) {
      const int L = o.levels.level[o.token_level_index];
      const uint64_t lsb = lsb_for_level(L);
      a.tok_and = ~lsb + 1;
      a.tok_or = lsb;
      a.tok_len = 16u - (uint32_t)((30 - L) >> 1);
      for (int k = 0; k < 4; ++k) {
        const int keep = (int)a.tok_len - 4 * k;
        a.tok_keep[k] = keep >= 4 ? 0xFFFFFFFFu : (keep <= 0 ? 0u : (1u << (8 * keep)) - 1u);
      }
      a.tokens16 = (uint4*)o.tokens16 + off;
      a.token_len = o.token_len ? o.token_len + off : nullptr;
    }
    const void* src = (const char*)in + off * in_stride;
    cudaError_t e;
    switch (kind) {
      case kInXYZ: e = launch_encode_kind<kInXYZ>(src, cnt, o.levels.n, token, a, st); break;
      case kInLatLngRad: e = launch_encode_kind<kInLatLngRad>(src, cnt, o.levels.n, token, a, st); break;
      default: e = launch_encode_kind<kInLatLngE7>(src, cnt, o.levels.n, token, a, st); break;
    }
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Strict host-libm mode: the host re-evaluated the boundary-sensitive points and found `m` whose leaf id differs under
// its libm; rewrite their outputs (every level + token) in place. m is tiny (usually 0: then this is not launched).
__global__ void __launch_bounds__(kThreads) encode_patch_kernel(const FixPatch* __restrict__ patches, uint32_t m, const EncodeArgs a, int levels, bool token) {
  const uint32_t k = blockIdx.x * kThreads + threadIdx.x;
  if (k >= m) return;
  const uint64_t i = patches[k].index - a.index_base, id = patches[k].id;
  for (int l = 0; l < levels; ++l) a.ids[l][i] = (id & a.and_mask[l]) | a.or_mask[l];
  if (token) {
    uint32_t w[4];
    const int len = cellid_to_token((id & a.tok_and) | a.tok_or, w);
    a.tokens16[i] = make_uint4(w[0], w[1], w[2], w[3]);
    if (a.token_len) a.token_len[i] = (uint8_t)len;
  }
}

cudaError_t launch_encode_patch(const FixPatch* patches_dev, uint32_t m, const EncodeOut& o, size_t n, cudaStream_t st) {
  if (m == 0) return cudaSuccess;
  const size_t ids_stride = o.ids_stride ? o.ids_stride : n;
  EncodeArgs a{};
  for (int k = 0; k < o.levels.n; ++k) {
    const uint64_t lsb = lsb_for_level(o.levels.level[k]);
    a.ids[k] = o.ids + (size_t)k * ids_stride;
    a.and_mask[k] = ~lsb + 1;
    a.or_mask[k] = lsb;
  }
  a.index_base = o.index_base;
  const bool token = o.token_level_index >= 0;
  if (token) {
    const uint64_t lsb = lsb_for_level(o.levels.level[o.token_level_index]);
    a.tok_and = ~lsb + 1;
    a.tok_or = lsb;
    a.tokens16 = (uint4*)o.tokens16;
    a.token_len = o.token_len;
  }
  encode_patch_kernel<<<(m + kThreads - 1) / kThreads, kThreads, 0, st>>>(patches_dev, m, a, o.levels.n, token);
  count_launch();
  return cudaGetLastError();
}

// Diagnostics (s2b_diag_filter_deviation_dev): how far the filter path's (fi, fj) are from the correctly rounded path's,
// in leaf cells — the quantity kFilterTolIj must dominate. Points whose two paths disagree on the face are counted
// separately; they must all be ones the filter sends to the correctly rounded path anyway (count in out[1] otherwise).
template <int kKind>
__global__ void __launch_bounds__(kThreads) filter_deviation_kernel(const double* __restrict__ in, size_t n, unsigned long long* __restrict__ out) {
  double worst = 0.0;
  unsigned long long unflagged_face_flips = 0;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    double fi0, fj0, fi1, fj1;
    int f0, f1;
    if (kKind == kInXYZ) {
      const P3 p{in[3 * i], in[3 * i + 1], in[3 * i + 2]};
      const double m = fmax(fmax(fabs(p.x), fabs(p.y)), fabs(p.z));
      if (!(m > 3.0549363634996047e-151 && m < 3.2733906078961419e+150)) continue;
      f0 = point_to_face_fij_filter(p, &fi0, &fj0);
      f1 = point_to_face_fij(p, &fi1, &fj1);
    } else {
      const double lat = in[2 * i], lng = in[2 * i + 1];
      if (!(fabs(lat) < 1048576.0 && fabs(lng) < 1048576.0)) continue;
      f0 = point_to_face_fij_filter(latlng_to_point_fast(lat, lng, g_sincos512), &fi0, &fj0);
      f1 = point_to_face_fij(latlng_to_point(lat, lng, g_sincos_tab), &fi1, &fj1);
    }
    if (f0 != f1) {
      if (!near_cell_boundary(fi0, fj0, kFilterTolIj)) ++unflagged_face_flips;
      continue;
    }
    worst = fmax(worst, fmax(fabs(fi0 - fi1), fabs(fj0 - fj1)));
  }
  atomicMax(out, (unsigned long long)__double_as_longlong(worst));   // non-negative doubles order like their bit patterns
  if (unflagged_face_flips) atomicAdd(out + 1, unflagged_face_flips);
}

cudaError_t launch_filter_deviation(InputKind kind, const double* in, size_t n, unsigned long long* out2_dev, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (kind == kInXYZ) filter_deviation_kernel<kInXYZ><<<grid_for(n, (const void*)filter_deviation_kernel<kInXYZ>), kThreads, 0, st>>>(in, n, out2_dev);
  else filter_deviation_kernel<kInLatLngRad><<<grid_for(n, (const void*)filter_deviation_kernel<kInLatLngRad>), kThreads, 0, st>>>(in, n, out2_dev);
  count_launch();
  return cudaGetLastError();
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

// S2CellId::ToLatLng = S2LatLng(ToPointRaw()) (s2cell_id.cc:381, s2latlng.h:235-250): (lat, lng) radians. atan2 is the CUDA
// math library's (documented max error 2 ulp); the reference uses the host libm, so this output is equal within a few ulp,
// not bit for bit (it is a floating-point result, not an id).
__global__ void __launch_bounds__(kThreads) to_latlng_kernel(const uint64_t* __restrict__ ids, size_t n, double* __restrict__ latlng) {
  __shared__ uint16_t s_ij[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_ij[k] = g_hilbert.ij[k];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const P3 p = cellid_to_point_raw(__ldcs(ids + i), s_ij);
    const double lat = atan2(p.z + 0.0, sqrt(p.x * p.x + p.y * p.y));
    const double lng = atan2(p.y + 0.0, p.x + 0.0);
    __stcs((double2*)(latlng + 2 * i), make_double2(lat, lng));
  }
}

cudaError_t launch_to_latlng(const uint64_t* ids, size_t n, double* latlng, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  to_latlng_kernel<<<grid_for(n, (const void*)to_latlng_kernel), kThreads, 0, st>>>(ids, n, latlng);
  count_launch();
  return cudaGetLastError();
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
