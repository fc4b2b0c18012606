// s2b_index.cu — K2/K3/K4: the device-resident flattened shape index and the batched point queries over it.
//
// One thread per point, persistent grid. Per point: K1 (leaf cell id, Hilbert table in shared memory) -> Locate
// (lookup-table bucket + 0-3 probes of the sorted cell ids, all L2/L1 resident) -> for the (few) points that land in
// an index cell, the crossing test of s2b_contains.cuh from the precomputed cell centre over that cell's clipped
// edges, with the Sign chain (triage -> stable -> exact superaccumulator -> symbolic) entirely on the device.
// Index arrays are read through the read-only path; points are streamed (24 B in, 1 B / a few atomics out).
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <cstdlib>
#include <new>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_contains.cuh"
#include "s2b_flat_host.h"
#include "s2b_hilbert.h"
#include "s2b_index_handle.h"
#include "s2b_region.cuh"
#include "s2b_internal.h"
#include "s2b_stream.cuh"

namespace s2b {

__device__ const HilbertTables g_hilbert_ix = kHilbert;
__device__ uint16_t g_hilbert_wide_ix[16384];   // uploaded on first use per device

static cudaError_t ensure_wide_table_ix() {
  static std::atomic<uint64_t> done_mask{0};   // bit d: device d holds the table (one host thread may drive several GPUs)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || (dev < 64 && ((done_mask.load(std::memory_order_acquire) >> dev) & 1))) return e;
  e = cudaMemcpyToSymbol(g_hilbert_wide_ix, wide_hilbert_host(), 16384 * sizeof(uint16_t));
  if (e == cudaSuccess && dev < 64) done_mask.fetch_or(1ull << dev, std::memory_order_release);
  return e;
}

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

// Two phases per batch, both on the caller's stream:
//  A  locate_kernel   every point: K1 + Locate. Lean (no predicates inlined, ~40 registers, 6 CTAs/SM). Points that
//                     land in an index cell are appended to a work list (point, cell) with one warp-aggregated
//                     atomicAdd per warp; every point's output slot gets its "miss" value.
//  B  process_kernel  one thread per work-list entry: crossing parity over the cell's clipped edges (s2b_contains.cuh),
//                     so the divergent, register-hungry part runs on fully populated warps instead of on the 1-10 %
//                     of lanes that hit, and does not cap the occupancy of phase A.
struct HitWork { uint32_t point; int32_t cell; };

// Work-list slots are reserved per warp in chunks of kSlotChunk with ONE global atomicAdd per chunk (a same-address
// atomic per warp iteration serialises at the L2: ~600k of them per 20M points when most warps see a hit). Unused slots
// of a chunk are filled with a sentinel (cell = -1) that phase B skips.
struct WarpSlots {
  uint32_t next = 0, end = 0;   // warp-uniform
  // Appends one entry for every lane with `want` set. dir = +1: list grows forward from work[0]; dir = -1: backward from
  // work[capacity-1]. All 32 lanes must call.
  __device__ __forceinline__ void append(bool want, HitWork value, HitWork* work, uint32_t capacity, unsigned int* counter, int dir, unsigned lane) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return;
    const uint32_t k = (uint32_t)__popc(m);
    if (next + k > end) {
      // retire the rest of the current chunk, then take a new one
      for (uint32_t s = next + lane; s < end; s += 32) work[dir > 0 ? s : capacity - 1 - s] = HitWork{0u, -1};
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(counter, kSlotChunk);
      next = __shfl_sync(0xffffffffu, base, 0);
      end = next + kSlotChunk;
    }
    if (want) {
      const uint32_t s = next + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (s < capacity) work[dir > 0 ? s : capacity - 1 - s] = value;   // never false: see work_list_capacity()
    }
    next += k;
  }
  __device__ __forceinline__ void finish(HitWork* work, uint32_t capacity, int dir, unsigned lane) {
    for (uint32_t s = next + lane; s < end; s += 32) work[dir > 0 ? s : capacity - 1 - s] = HitWork{0u, -1};
  }
};

// Slots a list can consume for `entries` appended items, written by `warps` warps: an append that does not fit retires
// the rest of its chunk (at most 31 slots, since an append holds at most 32 entries) and takes a new one, so a chunk
// holds at least kSlotChunk - 31 entries, and every warp leaves at most one partly used chunk at the end:
//   slots <= kSlotChunk * (entries / (kSlotChunk - 31) + 1 + warps)  <=  1.32 * entries + kSlotChunk * (warps + 1).
// The edge list and the interior list share one buffer from its two ends and a point is in at most one of them, so
// `entries` is bounded by the points of the round for the pair.
static size_t work_list_capacity(size_t entries, size_t warps) {
  return kSlotChunk * (entries / (kSlotChunk - 31) + 2 + warps);
}

// What phase A does with a located point. interior_mode says what may be answered on the spot for points in pure
// interior cells (no edges, every clip contains the centre):
//   0 nothing (fill)      1 "any": flag = 1      2 "count": per-point count = number of clips
//   3 histogram: cells with a single clip add 1 to that shape's counter right here
//   4 one shape: cells with a single clip compare its shape id with the target
// Everything else goes to the work lists: edge-bearing hits at the front of `work` (and, if cell_counts != nullptr, into the
// per-cell histogram of the counting sort), interior hits that still need per-shape output at the back.
struct PhaseAOut {
  HitWork* work; uint32_t capacity; unsigned int* counters;
  uint8_t* flags; uint32_t* counts;
  unsigned long long* hist; int target_shape; unsigned int* cell_counts;
};

// What is known about a located point when it is routed:
//   kHitNone     not in any index cell
//   kHitEdges    in cell c, which has clipped edges (the lookup table said so)           -> edge list
//   kHitSingle   in a pure interior cell whose only clipped shape is `shape` (from the table; c unknown)
//   kHitCell     in cell c, nothing else known: its info is read here (or was read by the caller: have_info)
enum HitKind { kHitNone = 0, kHitEdges = 1, kHitSingle = 2, kHitCell = 3 };

// Interior-list entries of a single-shape interior cell carry the shape instead of the cell: cell = -(shape + 2).
__device__ __forceinline__ int32_t encode_single_shape(int shape) { return -(shape + 2); }

template <int kInteriorMode>
__device__ __forceinline__ void route_located(const FlatIndexView& ix, const PhaseAOut& o, bool write_outputs, uint32_t i, int hit_kind, int c, int shape,
                                              WarpSlots& edge_slots, WarpSlots& interior_slots, unsigned lane, bool have_info = false, int loaded_info = 0) {
  uint8_t flag = 0;
  uint32_t cnt = 0;
  bool to_edges = hit_kind == kHitEdges, to_interior = false;
  int32_t interior_entry = c;
  if (hit_kind == kHitCell) {
    const int info = have_info ? loaded_info : ix.interior_info[c];
    if (info == 0) to_edges = true;
    else if (info < 0) { hit_kind = kHitSingle; shape = -info - 1; }
    else {                                               // pure interior cell with `info` clipped shapes
      if (kInteriorMode == 1) flag = 1;
      else if (kInteriorMode == 2) cnt = (uint32_t)info;
      else to_interior = true;
    }
  }
  if (hit_kind == kHitSingle) {
    if (kInteriorMode == 1) flag = 1;
    else if (kInteriorMode == 2) cnt = 1u;
    else if (kInteriorMode == 3) atomicAdd(o.hist + shape, 1ull);
    else if (kInteriorMode == 4) flag = shape == o.target_shape;
    else { to_interior = true; interior_entry = encode_single_shape(shape); }
  }
  // which per-point output exists is a property of the mode (flags: any / one shape; counts: count), known at compile time
  if (write_outputs) {
    if (kInteriorMode == 1 || kInteriorMode == 4) o.flags[i] = flag;
    if (kInteriorMode == 2) o.counts[i] = cnt;
  }
  if (to_edges && o.cell_counts) atomicAdd(o.cell_counts + c, 1u);
  edge_slots.append(to_edges, HitWork{i, c}, o.work, o.capacity, o.counters, +1, lane);
  if (kInteriorMode != 1 && kInteriorMode != 2) interior_slots.append(to_interior, HitWork{i, interior_entry}, o.work, o.capacity, o.counters + 1, -1, lane);
}

// Phase A, exact: FP64 K1 + Locate, then route_located. The kernel is a chain of dependent gathers per point
// (list -> point -> table -> candidate range -> ~6 probes of cell_ids -> cell info), so every stage is issued for all
// kPointsPerThread points of a thread before any of them is consumed: the loads of a stage are in flight together
// (the binary searches of a thread advance in lock step, one probe each per round).
template <int kInteriorMode, int kCtasPerSm, int kPointsPerThread>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
locate_kernel(const FlatIndexView ix, const double* __restrict__ xyz, uint32_t n_points, const uint32_t* __restrict__ list,
              const PhaseAOut o) {
  unsigned int* const counters = o.counters;
  // list == nullptr: points 0..n_points-1. Otherwise: the points whose indices are list[0..counters[2]) (entries equal
  // to 0xFFFFFFFF are unused reservation slots) — the ones the FP32 pre-filter could not decide.
  const uint32_t n = list ? counters[2] : n_points;
  __shared__ uint16_t s_pos[16384];   // 6-bit Hilbert table
  for (int k = threadIdx.x; k < 16384 / 8; k += kThreads) ((uint4*)s_pos)[k] = ((const uint4*)g_hilbert_wide_ix)[k];
  __syncthreads();
  constexpr int kP = kPointsPerThread;
  constexpr uint32_t kSpan = kThreads * kP;
  const uint32_t stride = gridDim.x * kSpan;
  const unsigned lane = threadIdx.x & 31u;
  const bool have_cells = ix.num_cells != 0;
  WarpSlots edge_slots, interior_slots;
  // whole warps iterate together (the ballots need all lanes)
  for (uint32_t base = blockIdx.x * kSpan + (threadIdx.x & ~31u); base < n; base += stride) {
    uint32_t idx[kP];
    P3 p[kP];
    // stage 0: the indices, then the points (independent gathers)
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      const uint32_t k = base + t * kThreads + lane;
      idx[t] = 0xFFFFFFFFu;
      if (k < n) idx[t] = list ? list[k] : k;
    }
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      if (idx[t] != 0xFFFFFFFFu) {
        const double* q = xyz + 3 * (size_t)idx[t];
        p[t] = P3{__ldcs(q), __ldcs(q + 1), __ldcs(q + 2)};
      } else {
        p[t] = P3{1.0, 0.0, 0.0};
      }
    }
    // stage 1: projection; stage 2: the lookup-table entries
    int face[kP];
    uint32_t ci[kP], cj[kP], e0[kP];
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      double fi, fj;
      face[t] = point_to_face_fij(p[t], &fi, &fj);
      ci[t] = fij_to_ij(fi); cj[t] = fij_to_ij(fj);
    }
#pragma unroll
    for (int t = 0; t < kP; ++t)
      e0[t] = (have_cells && idx[t] != 0xFFFFFFFFu) ? ix.lut[lut_bucket(ix.lut_level, face[t], ci[t], cj[t])] : kLutEmpty;
    // stage 3: candidate ranges of the mixed buckets; stage 4: their leaf ids
    uint32_t lo[kP], hi[kP];
    uint64_t leaf[kP];
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      lo[t] = hi[t] = 0;
      if ((e0[t] & ~kLutIndexMask) == kLutMixed) { const LutRange r = ix.mixed[e0[t] & kLutIndexMask]; lo[t] = r.lo; hi[t] = r.hi; }
    }
#pragma unroll
    for (int t = 0; t < kP; ++t)
      leaf[t] = (e0[t] & ~kLutIndexMask) == kLutMixed ? from_face_ij_wide(face[t], ci[t], cj[t], s_pos) : 0;
    // stage 5: the searches (locate_cell's loop: last cell whose range_min <= leaf), one probe per point per round
    bool any = true;
    while (any) {
      uint32_t mid[kP];
      uint64_t probe[kP];
#pragma unroll
      for (int t = 0; t < kP; ++t) {
        mid[t] = (lo[t] + hi[t] + 1) >> 1;
        probe[t] = lo[t] < hi[t] ? ix.cell_ids[mid[t]] : 0;
      }
      any = false;
#pragma unroll
      for (int t = 0; t < kP; ++t) {
        if (lo[t] < hi[t]) {
          if (cellid_range_min(probe[t]) <= leaf[t]) lo[t] = mid[t]; else hi[t] = mid[t] - 1;
          any |= lo[t] < hi[t];
        }
      }
    }
    // stage 6: the candidate's id and, speculatively, its interior info (both depend only on the search result)
    int cell[kP], info[kP];
    uint64_t cand[kP];
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      const uint32_t kind = e0[t] & ~kLutIndexMask;
      cell[t] = kind == kLutMixed ? (int)lo[t] : (kind == kLutInside ? (int)(e0[t] & kLutCellMask) : -1);
      cand[t] = kind == kLutMixed ? ix.cell_ids[cell[t]] : 0;
      info[t] = kind == kLutMixed ? ix.interior_info[cell[t]] : 0;
    }
    // stage 7: outputs, work lists
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      const uint32_t i = idx[t];
      const bool valid = i != 0xFFFFFFFFu;
      const uint32_t kind = e0[t] & ~kLutIndexMask;
      int hit = kHitNone;
      if (kind == kLutMixed) hit = (cellid_range_min(cand[t]) <= leaf[t] && leaf[t] <= cellid_range_max(cand[t])) ? kHitCell : kHitNone;
      else if (kind == kLutInside) hit = (e0[t] & kLutHasEdges) ? kHitEdges : kHitCell;
      else if (kind == kLutInterior) hit = kHitSingle;
      if (!valid) hit = kHitNone;
      route_located<kInteriorMode>(ix, o, valid, i, hit, cell[t], (int)(e0[t] & kLutIndexMask), edge_slots, interior_slots, lane, kind == kLutMixed, info[t]);
    }
  }
  edge_slots.finish(o.work, o.capacity, +1, lane);
  interior_slots.finish(o.work, o.capacity, -1, lane);
}

// Phase A0: FP32 pre-filter for Locate. Most points are nowhere near a boundary that matters, so the face, (i,j) and
// the lookup-table bucket are first computed in single precision (approximate reciprocal and square root: ~110
// instructions instead of ~300), with a conservative error radius kFastRadius (in leaf cells) around the result:
//   bucket kind EMPTY  and the point is > radius inside the bucket  -> certain miss
//   bucket kind INSIDE (whole bucket inside index cell c), same margin -> certain hit of cell c
//   anything else (mixed bucket, near a bucket/face boundary, NaN, ...) -> index appended to the "uncertain" list, which
//   the exact FP64 kernel (locate_kernel) then processes. Results are therefore exactly those of the FP64 path.
// Error budget: double->float inputs 2^-24 each, hardware approximate reciprocal and square root <= 2 ulp each, a few
// roundings: |u| error < 5e-7, s error < 7.5e-7, i.e. < 800 leaf cells at 2^30; kFastRadius = 2048 leaves 2.5x margin. A wrong face choice needs
// two |components| within 6e-8 of each other, which puts (i,j) within ~60 leaf cells of a face edge: inside the radius.
// The point stream is staged through shared memory by the TMA unit, double buffered: one thread issues one bulk copy per
// 12 KB tile (kThreads * kPointsPerThread points) and the tile after the current one is in flight while the current one
// is processed, so the bytes in flight occupy neither registers nor issue slots. (History: with register staging this
// kernel was DRAM-latency-bound at 2.0 TB/s; with per-thread cp.async the copy loop itself cost ~35 instructions per
// point; per-WARP tiles of 1.5 KB — no block barrier — were slower still, 112 vs 149 Gpts/s: eight times as many bulk
// copies.) Requires xyz to be 16-byte aligned; an odd point count leaves 8 bytes that the issuing thread copies itself.
// (Measured and rejected here: folding the located cell's info into 8-byte table entries — one load instead of two —
// 113 vs 149 Gpts/s on uniform points: the 50 MB table no longer stays in the L2 next to the point stream; issuing the
// table and info loads of a thread's two points together before routing either: join +2 %, uniform points -5 %.)
template <int kInteriorMode, int kPointsPerThread>
__global__ void __launch_bounds__(kThreads, 6)
locate_fast_kernel(const FlatIndexView ix, const double* __restrict__ xyz, uint32_t n, uint32_t* __restrict__ uncertain, const PhaseAOut o) {
  unsigned int* const counters = o.counters;
  constexpr int kP = kPointsPerThread;
  constexpr uint32_t kTilePoints = kThreads * kP;
  constexpr uint32_t kTileBytes = kTilePoints * 24;
  __shared__ __align__(128) double s_tile[2][kTilePoints * 3];
  __shared__ __align__(8) uint64_t s_bar[2];
  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u;
  const int lut_level = ix.lut_level;                           // 2..10
  const int bucket_bits = 30 - lut_level;                       // log2 of the bucket size in leaf cells (>= 20)
  const uint32_t bucket_mask = (1u << bucket_bits) - 1u;
  const uint32_t margin_span = bucket_mask - 2u * (uint32_t)kFastRadius;
  const uint64_t total_bytes = (uint64_t)n * 24;
  const uint32_t num_tiles = (n + kTilePoints - 1) / kTilePoints;
  WarpSlots edge_slots, interior_slots;
  U32Slots unc_slots;

  auto issue_tile = [&](uint32_t tile, int stage) {             // thread 0 only
    const uint64_t tile_byte0 = (uint64_t)tile * kTileBytes;
    const uint64_t left = total_bytes - tile_byte0;
    const uint32_t bytes = left >= kTileBytes ? kTileBytes : (uint32_t)left;
    const uint32_t bulk = bytes & ~15u;
    if (bytes != bulk) s_tile[stage][bulk / 8] = xyz[(tile_byte0 + bulk) / 8];   // made visible to the waiters by the arrive below
    mbar_expect_tx(&s_bar[stage], bulk);
    if (bulk) bulk_copy_g2s(s_tile[stage], (const char*)xyz + tile_byte0, bulk, &s_bar[stage]);
  };

  uint32_t tile = blockIdx.x;
  int stage = 0;
  uint32_t parity_bits = 0;                                     // bit s = phase parity of stage s's barrier
  if (tile < num_tiles && threadIdx.x == 0) issue_tile(tile, 0);
  for (; tile < num_tiles; tile += gridDim.x, stage ^= 1) {
    const uint32_t next = tile + gridDim.x;
    if (next < num_tiles && threadIdx.x == 0) issue_tile(next, stage ^ 1);   // that buffer was released by the barrier below
    mbar_wait(&s_bar[stage], (parity_bits >> stage) & 1u);
    parity_bits ^= 1u << stage;
    const double* pts = s_tile[stage];
    const uint32_t tile_point0 = tile * kTilePoints;
#pragma unroll
    for (int t = 0; t < kP; ++t) {
      const uint32_t li = t * kThreads + threadIdx.x;
      const uint32_t i = tile_point0 + li;
      const bool valid = i < n;
      // Slots past the end of the batch hold stale or uninitialised bytes: they are read like any point (garbage is just an
      // "undecided" candidate) and `valid` gates everything that is written.
      const float x = (float)pts[3 * li], y = (float)pts[3 * li + 1], z = (float)pts[3 * li + 2];
      const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
      // major axis as XYZtoFaceUV picks it (largest |component|, first wins ties), spelled as selects: no branch
      const bool xy = ax > ay, xz = ax > az, yz = ay > az;
      const bool axis0 = xy & xz, axis1 = !xy & yz;
      const float m = axis0 ? x : (axis1 ? y : z);
      const bool neg = m < 0.0f;
      const float a = axis0 ? y : -x;
      const float b = (axis0 | axis1) ? z : -y;
      const float inv = rcp_approx(m);
      const int fi = approx_leaf_coord((neg ? b : a) * inv);
      const int fj = approx_leaf_coord((neg ? a : b) * inv);
      const int face = (axis0 ? 0 : (axis1 ? 1 : 2)) + (neg ? 3 : 0);
      // certain only if both coordinates are valid and well inside their lookup-table bucket (NaN/garbage gives fi = 0 or
      // saturated / negative values, which fail these tests). Unsigned arithmetic folds each pair of bounds into one compare.
      const uint32_t ufi = (uint32_t)fi, ufj = (uint32_t)fj;
      const uint32_t ri = ufi & bucket_mask, rj = ufj & bucket_mask;
      bool certain = valid & ((ufi | ufj) < (1u << 30)) & (max(ri - (uint32_t)kFastRadius, rj - (uint32_t)kFastRadius) <= margin_span);
      // the table is addressed by (face, i, j) of the bucket: no Hilbert encoding on this path. The masked coordinates
      // always give a valid bucket, so the read is unconditional (no branch around it).
      const uint32_t e0 = ix.lut[lut_bucket(lut_level, face, ufi & 0x3FFFFFFFu, ufj & 0x3FFFFFFFu)];
      // Everything a decided point needs is in the table entry (no second load unless its cell is a pure interior cell
      // with several clipped shapes). The top three bits of the entry (kind, has-edges) index a packed 2-bit table of
      // hit kinds: mixed -> none (undecided), empty -> none, inside/no edges -> kHitCell, inside/edges -> kHitEdges,
      // single-shape interior -> kHitSingle. Undecided points are routed (and their outputs written) by the exact kernel.
      certain &= (e0 >> 30) != 0u;
      const int hit = certain ? (int)((0xA700u >> (2u * (e0 >> 29))) & 3u) : kHitNone;
      route_located<kInteriorMode>(ix, o, certain, i, hit, (int)(e0 & kLutCellMask), (int)(e0 & kLutIndexMask), edge_slots, interior_slots, lane);
      unc_slots.append(valid & !certain, i, uncertain, counters + 2, lane);
    }
    __syncthreads();                                            // everyone is done with this stage before it is refilled
  }
  edge_slots.finish(o.work, o.capacity, +1, lane);
  interior_slots.finish(o.work, o.capacity, -1, lane);
  unc_slots.finish(uncertain, lane);
}

// Cell-major ordering of the edge-bearing hits (large indexes only). Phase B is bound by its memory access pattern: with
// hits in point order every thread reads a different cell's record, clips and edges (L2 hit rate 59 % on the 136 MB join
// index, 12 of 32 lanes busy because neighbouring lanes have different edge counts). A counting sort by cell index —
// per-cell hit counts (accumulated by phase A as it appends), exclusive scan, scatter — makes neighbouring threads work on the same cell: coalesced index
// reads and identical loop trip counts.
__global__ void __launch_bounds__(kThreads)
cell_scatter_kernel(const HitWork* __restrict__ work, const unsigned int* __restrict__ counters, unsigned int* __restrict__ cell_cursor,
                    HitWork* __restrict__ sorted) {
  const unsigned int count = counters[0];
  for (unsigned int w = blockIdx.x * kThreads + threadIdx.x; w < count; w += gridDim.x * kThreads) {
    const HitWork hw = work[w];
    if (hw.cell >= 0) sorted[atomicAdd(cell_cursor + hw.cell, 1u)] = hw;
  }
}

// One work item: crossing parity of point `hw.point` against every clipped shape of cell `hw.cell`.
template <int kMode>
__device__ __forceinline__ void process_hit(const FlatIndexView& ix, const double* __restrict__ xyz, const HitWork hw, int model, const QueryOut& out) {
  const double* q = xyz + 3 * (size_t)hw.point;
  const P3 p{q[0], q[1], q[2]};
  const FlatCell fc = ix.cells[hw.cell];
  const P3 a{fc.cx, fc.cy, fc.cz};
  const P3 a_cross_b = cross(a, p);
  uint32_t result = 0;
  uint32_t fill_at = kMode == kModeFill ? out.offsets[hw.point] : 0u;
  for (uint32_t k = 0; k < fc.clip_count; ++k) {
    const FlatClip clip = ix.clips[fc.clip_begin + k];
    if (kMode == kModeShape && clip.shape_id != out.shape_id) continue;
    if (clip_contains(ix, clip, a, p, a_cross_b, model)) {
      if (kMode == kModeAny || kMode == kModeShape) { result = 1; break; }
      if (kMode == kModeHistogram) atomicAdd(out.counts + clip.shape_id, 1ull);
      if (kMode == kModeCount) ++result;
      if (kMode == kModeFill) out.shape_ids[fill_at++] = clip.shape_id;
    } else if (kMode == kModeShape) {
      break;
    }
  }
  if (kMode == kModeAny || kMode == kModeShape) { if (result) out.flags[hw.point] = 1; }
  if (kMode == kModeCount) { if (result) out.per_point[hw.point] = result; }
}

// Phase B, both work lists in one launch: first the edge-bearing hits — `edge_list`[0 .. *edge_count), either the forward
// list of phase A or its cell-sorted copy — then, when `with_interior`, the interior hits that still need per-shape
// output, stored backward from work_base[capacity - 1] (their number is counters[1]). (As two launches the second one,
// a few thousand items, cost as much as half of the first: launch latency and an almost empty GPU.)
template <int kMode, int kCtasPerSm>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
process_kernel(const FlatIndexView ix, const double* __restrict__ xyz, const HitWork* __restrict__ edge_list, const unsigned int* __restrict__ edge_count,
               const HitWork* __restrict__ work_base, uint32_t capacity, const unsigned int* __restrict__ counters, bool with_interior, int model,
               const QueryOut out) {
  const unsigned int stride = gridDim.x * kThreads;
  const unsigned lane = threadIdx.x & 31u;
  for (int pass = 0; pass < (with_interior ? 2 : 1); ++pass) {
    const unsigned int count = pass == 0 ? *edge_count : counters[1];
    const HitWork* work = pass == 0 ? edge_list : work_base + (capacity - count);
    // warp-uniform trip count + __syncwarp(): lanes with long edge lists or exact-stage work rejoin every iteration
    for (unsigned int wbase = blockIdx.x * kThreads + (threadIdx.x & ~31u); wbase < count; wbase += stride, __syncwarp()) {
      const unsigned int w = wbase + lane;
      if (w >= count) continue;
      const HitWork hw = work[w];
      if (hw.cell == -1) continue;   // unused slot of a reservation chunk
      if (hw.cell < -1) {            // a pure interior cell with the single shape -(cell + 2) (only the fill mode lists these)
        if (kMode == kModeFill) out.shape_ids[out.offsets[hw.point]] = -(hw.cell + 2);
        continue;
      }
      process_hit<kMode>(ix, xyz, hw, model, out);
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

// S2ShapeIndex::Iterator::Locate(S2CellId target) = S2CellIterator::LocateImpl (s2cell_iterator.h:159-189): relation of a
// target CELL to the index cells (0 INDEXED: contained in index cell `cell`; 1 SUBDIVIDED: it contains index cells, the
// first of them is `cell`; 2 DISJOINT, cell = -1) — the classification step of coverers and region joins.
__global__ void __launch_bounds__(kThreads)
locate_cells_kernel(const FlatIndexView ix, const uint64_t* __restrict__ targets, size_t n, uint8_t* __restrict__ relation, int32_t* __restrict__ cell) {
  const uint32_t C = ix.num_cells;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const uint64_t target = __ldcs(targets + t);
    const uint64_t tmin = cellid_range_min(target), tmax = cellid_range_max(target);
    uint32_t lo = 0, hi = C;                                   // Seek(target.range_min()): first index cell with id >= tmin
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ix.cell_ids[mid] < tmin) lo = mid + 1; else hi = mid; }
    uint8_t rel = 2;
    int32_t at = -1;
    bool decided = false;
    if (lo < C) {
      const uint64_t id = ix.cell_ids[lo];
      if (id >= target && cellid_range_min(id) <= target) { rel = 0; at = (int32_t)lo; decided = true; }
      else if (id <= tmax) { rel = 1; at = (int32_t)lo; decided = true; }
    }
    if (!decided && lo > 0 && cellid_range_max(ix.cell_ids[lo - 1]) >= target) { rel = 0; at = (int32_t)lo - 1; }
    relation[t] = rel;
    if (cell) cell[t] = at;
  }
}

// S2ShapeIndex::Iterator::Locate(const S2Point&) (S2CellIterator::LocateImpl, s2cell_iterator.h:137-157) as a query of its
// own: cell[i] = position of the index cell containing point i, or -1. Exact FP64 projection + the lookup table.
__global__ void __launch_bounds__(kThreads)
locate_points_kernel(const FlatIndexView ix, const double* __restrict__ xyz, size_t n, int32_t* __restrict__ cell) {
  __shared__ uint16_t s_pos[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_ix.pos[k];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const double* q = xyz + 3 * t;
    cell[t] = locate_point(ix, P3{__ldcs(q), __ldcs(q + 1), __ldcs(q + 2)}, s_pos);
  }
}

// S2ContainsPointQuery::VisitIncidentEdges (s2contains_point_query.h:305-328): the edges of the located cell that have
// the query point as an endpoint (operator== on S2Point, so -0 == +0), as (shape id, edge id) pairs in the reference's
// visiting order (clipped shapes in cell order, edges in clipped order). kFill = false: per-point counts; true: pairs
// written at offsets[i]. One thread per point, exact FP64 Locate (this is a low-volume query next to Contains).
template <bool kFill>
__global__ void __launch_bounds__(kThreads)
incident_edges_kernel(const FlatIndexView ix, const int32_t* __restrict__ edge_ids, const double* __restrict__ xyz, size_t n,
                      uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, int32_t* __restrict__ shape_out, int32_t* __restrict__ edge_out) {
  __shared__ uint16_t s_pos[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_ix.pos[k];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t t = (size_t)blockIdx.x * kThreads + threadIdx.x; t < n; t += stride) {
    const double* q = xyz + 3 * t;
    const P3 p{q[0], q[1], q[2]};
    uint32_t cnt = 0;
    uint32_t at = kFill ? offsets[t] : 0u;
    const int c = locate_point(ix, p, s_pos);
    if (c >= 0) {
      const FlatCell fc = ix.cells[c];
      for (uint32_t k = 0; k < fc.clip_count; ++k) {
        const FlatClip clip = ix.clips[fc.clip_begin + k];
        for (uint32_t e = 0; e < clip.num_edges; ++e) {
          const FlatEdge ed = ix.edges[clip.edge_begin + e];
          if (eq(edge_v0(ed), p) || eq(edge_v1(ed), p)) {
            if (kFill) { shape_out[at] = clip.shape_id; edge_out[at] = edge_ids[clip.edge_begin + e]; ++at; }
            ++cnt;
          }
        }
      }
    }
    if (!kFill) counts[t] = cnt;
  }
}

// Number of non-zero flags (the hit count of a Contains batch): 16 flags per load, one atomic per CTA.
__global__ void __launch_bounds__(kThreads)
count_flags_kernel(const uint8_t* __restrict__ flags, size_t n, unsigned long long* __restrict__ count) {
  unsigned int local = 0;
  const size_t n16 = n / 16;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t k = (size_t)blockIdx.x * kThreads + threadIdx.x; k < n16; k += stride) {
    const uint4 v = __ldcs((const uint4*)flags + k);
    // flags are 0 or 1 per byte for this library's outputs, but any non-zero byte counts: fold each byte to its low bit
    auto nz = [](uint32_t w) { w |= w >> 4; w |= w >> 2; w |= w >> 1; return __popc(w & 0x01010101u); };
    local += nz(v.x) + nz(v.y) + nz(v.z) + nz(v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 15)) local += flags[n16 * 16 + threadIdx.x] != 0;
  for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  __shared__ unsigned int s_part[kThreads / 32];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += s_part[w];
    if (t) atomicAdd(count, t);
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

// Points per phase-A/phase-B round. A round is ~8 launches of persistent kernels, each with a tail in which SMs idle, so few
// large rounds beat many small ones: measured on the 1B-point join, 2^25 / 2^26 / 2^27 / 2^28 / 2^29 points per round give
// 68 / 81 / 90 / 96 / 98 Gpts/s. 2^28 needs 24 B of scratch per point of the round in the worst case (work list 10.7 B,
// cell-sorted copy 8 B, undecided list 5.3 B: 6.4 GB), so the round shrinks when the allocation fails.
#ifndef S2B_ROUND_BITS
#define S2B_ROUND_BITS 28
#endif
constexpr size_t kRoundPoints = size_t(1) << S2B_ROUND_BITS;
constexpr size_t kMinRoundPoints = size_t(1) << 22;

// The work list comes from the stream-ordered allocator; keep freed blocks cached in the pool (the default release
// threshold of 0 hands them back to the driver at every synchronisation, which costs milliseconds per call).
static void keep_pool_memory_once() {
  static std::atomic<uint64_t> done_mask{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64 || ((done_mask.load() >> dev) & 1)) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_mask.fetch_or(1ull << dev);
}

// Diagnostics (s2b_diag_force_exact_locate): skip the FP32 pre-filter, every point goes through the FP64 Locate.
static bool g_force_exact_locate = false;

// Launch shapes, fixed after measurement on B200 (DESIGN.md "Measured and rejected" lists the alternatives): exact Locate
// 4 CTAs/SM x 2 points per thread, FP32 pre-filter 2 points per thread, phase B 2 CTAs/SM.
#ifndef S2B_LOCATE_POINTS
#define S2B_LOCATE_POINTS 4
#endif
#ifndef S2B_LOCATE_CTAS
#define S2B_LOCATE_CTAS 3
#endif
#ifndef S2B_FAST_POINTS
#define S2B_FAST_POINTS 2
#endif
constexpr int kLocateCtas = S2B_LOCATE_CTAS, kLocatePoints = S2B_LOCATE_POINTS, kFastPoints = S2B_FAST_POINTS, kProcessCtas = 2;

template <int kMode>
static cudaError_t launch_query_t(const FlatIndexView& ix, const double* xyz, size_t n, int model, const QueryOut& out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  keep_pool_memory_once();
  {
    cudaError_t te = ensure_wide_table_ix();
    if (te != cudaSuccess) return te;
  }
  HitWork* work = nullptr;
  unsigned int* counter = nullptr;
  // worst-case slot consumption (work_list_capacity): two lists x the warps of the two phase-A kernels (<= 6 + 4 CTAs
  // per SM); the undecided list is written by the pre-filter's warps only
  const size_t max_warps = (size_t)sm_count() * 8 * (kThreads / 32);
  const bool use_fast = !g_force_exact_locate && ((uintptr_t)xyz % 16 == 0);   // cp.async staging needs 16-byte alignment
  // cell-major ordering of the edge hits pays off when the index is too large to stay cache resident and has enough
  // cells for the per-cell atomics not to collide (S2B_SORT_BY_CELL=0/1 overrides)
  const size_t C = ix.num_cells;
  bool sort_by_cell = C >= 65536;
  if (const char* ev = getenv("S2B_SORT_BY_CELL")) sort_by_cell = atoi(ev) != 0 && C > 0;
  size_t scan_tmp_bytes = 0;
  if (sort_by_cell) cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp_bytes, (unsigned int*)nullptr, (unsigned int*)nullptr, (int)(C + 1), st);
  auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t round = n < kRoundPoints ? n : kRoundPoints;
  size_t capacity = 0, b_work = 0, b_unc = 0, b_sorted = 0;
  const size_t b_counter = 256, b_cells = sort_by_cell ? up256(sizeof(unsigned int) * (C + 1)) : 0, b_scan = sort_by_cell ? up256(scan_tmp_bytes) : 0;
  cudaError_t e;
  for (;;) {
    capacity = work_list_capacity(round, 4 * max_warps);
    const size_t unc_capacity = work_list_capacity(round, max_warps);
    b_work = up256(sizeof(HitWork) * capacity);
    b_unc = use_fast ? up256(sizeof(uint32_t) * unc_capacity) : 0;
    b_sorted = sort_by_cell ? up256(sizeof(HitWork) * round) : 0;
    e = cudaMallocAsync((void**)&work, b_work + b_counter + b_unc + b_sorted + b_cells + b_scan, st);
    if (e != cudaErrorMemoryAllocation || round <= kMinRoundPoints) break;
    cudaGetLastError();                                         // not enough memory for this round size: halve it
    round = (round + 1) / 2;
  }
  if (e != cudaSuccess) return e;
  counter = (unsigned int*)((char*)work + b_work);   // [0] edge slots, [1] interior slots, [2] undecided points
  HitWork* sorted = (HitWork*)((char*)counter + b_counter + b_unc);
  unsigned int* cell_counts = (unsigned int*)((char*)sorted + b_sorted);
  void* scan_tmp = (char*)cell_counts + b_cells;
  uint32_t* uncertain = (uint32_t*)((char*)counter + 256);
  constexpr int kInteriorMode = kMode == kModeAny ? 1 : (kMode == kModeCount ? 2 : (kMode == kModeHistogram ? 3 : (kMode == kModeShape ? 4 : 0)));
  const auto proc = process_kernel<kMode, kProcessCtas>;
  const auto fast = locate_fast_kernel<kInteriorMode, kFastPoints>;
  const auto exact = locate_kernel<kInteriorMode, kLocateCtas, kLocatePoints>;
  const int proc_grid = grid_for_ix(round, (const void*)proc);
  for (size_t off = 0; off < n && e == cudaSuccess; off += round) {
    const size_t cnt = n - off < round ? n - off : round;
    e = cudaMemsetAsync(counter, 0, 4 * sizeof(unsigned int), st);
    if (e != cudaSuccess) break;
    QueryOut o = out;   // rebase per-point arrays to this round
    if (o.flags) o.flags += off;
    if (o.per_point) o.per_point += off;
    if (o.offsets) o.offsets += off;
    uint8_t* fl = (kMode == kModeAny || kMode == kModeShape) ? o.flags : nullptr;
    uint32_t* pc = kMode == kModeCount ? o.per_point : nullptr;
    if (sort_by_cell) {
      // counting sort of list 0 by cell; the per-cell hit counts are accumulated by phase A itself
      e = cudaMemsetAsync(cell_counts, 0, sizeof(unsigned int) * (C + 1), st);
      if (e != cudaSuccess) break;
    }
    const PhaseAOut pa{work, (uint32_t)capacity, counter, fl, pc, o.counts, o.shape_id, sort_by_cell ? cell_counts : nullptr};
    const uint32_t* list = nullptr;
    if (use_fast) {
      fast<<<grid_for_ix((cnt + kFastPoints - 1) / kFastPoints, (const void*)fast), kThreads, 0, st>>>(ix, xyz + 3 * off, (uint32_t)cnt, uncertain, pa);
      count_launch();
      list = uncertain;
    }
    exact<<<grid_for_ix((cnt + kLocatePoints - 1) / kLocatePoints, (const void*)exact), kThreads, 0, st>>>(ix, xyz + 3 * off, (uint32_t)cnt, list, pa);
    count_launch();
    if (sort_by_cell) {
      // cell_counts[C+1] -> exclusive scan (cell_counts[C] = number of valid hits) -> scatter
      const int sort_grid = grid_for_ix(round, (const void*)cell_scatter_kernel);
      e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_tmp_bytes, cell_counts, cell_counts, (int)(C + 1), st);
      if (e != cudaSuccess) break;
      count_launch();
      // the scatter advances cell_counts[c] from the start to the end of cell c's segment; cell_counts[C] stays the total
      cell_scatter_kernel<<<sort_grid, kThreads, 0, st>>>(work, counter, cell_counts, sorted);
      count_launch();
    }
    constexpr bool kWithInterior = kInteriorMode != 1 && kInteriorMode != 2;
    // cell_counts[C] (after the scan) = number of valid edge hits = length of the sorted list
    proc<<<proc_grid, kThreads, 0, st>>>(ix, xyz + 3 * off, sort_by_cell ? sorted : work, sort_by_cell ? cell_counts + C : counter, work,
                                         (uint32_t)capacity, counter, kWithInterior, model, o);
    count_launch();
    e = cudaGetLastError();
  }
  cudaError_t e2 = cudaFreeAsync(work, st);
  return e != cudaSuccess ? e : e2;
}

}  // namespace s2b

using namespace s2b;

// ---------------------------------------------------------------------------------------------------

namespace s2b {
int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
cudaError_t region_exact_stage_count(unsigned long long* count, bool reset);   // s2b_region.cu
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
  const uint64_t M = hf.mixed.size();
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
         o_mixed = o_lut + align(4 * nb), o_kind = o_mixed + align(sizeof(LutRange) * M), o_idtab = o_kind + align(4 * (C + 1));
  // Hilbert-order bucket table over the cell ids (Locate(S2CellId), s2b_region.cuh): about two buckets per cell
  int id_level = 0;
  while (id_level < 11 && (size_t)(6u << (2 * id_level)) < 2 * C) ++id_level;
  const uint32_t id_buckets = 6u << (2 * id_level);
  const int id_shift = 61 - 2 * id_level;
  std::vector<uint32_t> id_first((size_t)id_buckets + 1, 0);
  {
    size_t k = 0;
    for (uint32_t b = 0; b <= id_buckets; ++b) {
      while (k < C && (f->cell_ids[k] >> id_shift) < b) ++k;
      id_first[b] = (uint32_t)k;
    }
  }
  size_t o_view = o_idtab + align(4 * ((size_t)id_buckets + 1)), total = o_view + align(sizeof(FlatIndexView));
  e = cudaMalloc(&ix->blob, total);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc(index)"); }
  ix->blob_bytes = total;
  char* base = (char*)ix->blob;
  bool ok = true;
  ok &= C == 0 || cudaMemcpy(base + o_ids, f->cell_ids, 8 * C, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= C == 0 || cudaMemcpy(base + o_cells, cells.data(), sizeof(FlatCell) * C, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= K == 0 || cudaMemcpy(base + o_clips, clips.data(), sizeof(FlatClip) * K, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= E == 0 || cudaMemcpy(base + o_edges, edges.data(), sizeof(FlatEdge) * E, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(base + o_lut, lut.data(), 4 * nb, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(base + o_mixed, hf.mixed.data(), sizeof(LutRange) * M, cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(base + o_kind, hf.interior_info.data(), 4 * (C + 1), cudaMemcpyHostToDevice) == cudaSuccess;
  ok &= cudaMemcpy(base + o_idtab, id_first.data(), 4 * id_first.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { e = cudaGetLastError(); cudaFree(ix->blob); delete ix; return cuda_fail(e, "cudaMemcpy(index)"); }
  ix->view.cell_ids = (const uint64_t*)(base + o_ids);
  ix->view.cells = (const FlatCell*)(base + o_cells);
  ix->view.clips = (const FlatClip*)(base + o_clips);
  ix->view.edges = (const FlatEdge*)(base + o_edges);
  ix->view.lut = (const uint32_t*)(base + o_lut);
  ix->view.mixed = (const LutRange*)(base + o_mixed);
  ix->view.interior_info = (const int32_t*)(base + o_kind);
  ix->view.id_first = (const uint32_t*)(base + o_idtab);
  ix->view.id_buckets = id_buckets;
  ix->view.id_shift = id_shift;
  ix->view.num_cells = (uint32_t)C;
  ix->view.lut_level = hf.lut_level;
  ix->view.num_shape_ids = f->num_shape_ids;
  ix->num_clips = K; ix->num_edges = E;
  for (const FlatEdge& ed : edges)
    if (!region_edge_is_plain(edge_v0(ed), edge_v1(ed))) ++ix->extended_region_edges;
  if (cudaMemcpy(base + o_view, &ix->view, sizeof(FlatIndexView), cudaMemcpyHostToDevice) != cudaSuccess) {
    e = cudaGetLastError(); cudaFree(ix->blob); delete ix; return cuda_fail(e, "cudaMemcpy(view)");
  }
  ix->view_dev = (const FlatIndexView*)(base + o_view);
  *out = ix;
  return S2B_OK;
}

// A cell union as an index: every cell becomes a pure interior cell of shape 0 (no edges, contains_center set), so
// S2CellUnion::Contains(S2Point) (s2cell_union.cc:564-566) is s2b_contains over it and runs on the K3 Locate path
// (lookup table + FP32 pre-filter) instead of a per-point binary search.
int s2b_index_create_from_cellunion(const uint64_t* sorted_ids, size_t m, S2bIndex** out) {
  if (!out || (m && !sorted_ids)) return set_error(S2B_EINVAL, "null argument");
  std::vector<uint32_t> clip_begin(m + 1), edge_begin(m + 1, 0);
  std::vector<int32_t> shape(m, 0);
  std::vector<uint8_t> flags(m, (uint8_t)(1 | (2 << 1)));
  for (size_t i = 0; i <= m; ++i) clip_begin[i] = (uint32_t)i;
  S2bFlatIndex f{};
  f.num_cells = m; f.num_clips = m; f.num_edges = 0; f.num_shape_ids = 1;
  f.cell_ids = sorted_ids; f.cell_center = nullptr; f.cell_clip_begin = clip_begin.data();
  f.clip_shape_id = shape.data(); f.clip_flags = flags.data(); f.clip_edge_begin = edge_begin.data();
  f.edge_v0 = nullptr; f.edge_v1 = nullptr;
  return s2b_index_create(&f, out);
}

void s2b_index_destroy(S2bIndex* ix) {
  if (!ix) return;
  if (ix->replicas) { s2b_replica_set_destroy(ix->replicas); ix->replicas = nullptr; }
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != ix->device) cudaSetDevice(ix->device);
  if (ix->blob) cudaFree(ix->blob);
  if (ix->edge_ids) cudaFree(ix->edge_ids);
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

int s2b_index_locate_points_dev(const S2bIndex* ix, const double* xyz_dev, size_t n, int32_t* cell_out_dev, void* stream) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (n == 0) return S2B_OK;
  if (!xyz_dev || !cell_out_dev) return set_error(S2B_EINVAL, "null pointer");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the index lives on device %d but the current device is %d", ix->device, dev);
  locate_points_kernel<<<grid_for_ix(n, (const void*)locate_points_kernel), kThreads, 0, (cudaStream_t)stream>>>(ix->view, xyz_dev, n, cell_out_dev);
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "locate_points launch");
}

int s2b_index_set_edge_ids(S2bIndex* ix, const int32_t* edge_id, size_t num_edges) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (num_edges != ix->num_edges) return set_error(S2B_EINVAL, "the index has %llu clipped edges, %zu edge ids were given", (unsigned long long)ix->num_edges, num_edges);
  if (num_edges && !edge_id) return set_error(S2B_EINVAL, "null edge ids");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the index lives on device %d but the current device is %d", ix->device, dev);
  if (ix->edge_ids) { cudaFree(ix->edge_ids); ix->edge_ids = nullptr; }
  e = cudaMalloc((void**)&ix->edge_ids, sizeof(int32_t) * (num_edges ? num_edges : 1));
  if (e != cudaSuccess) { ix->edge_ids = nullptr; return cuda_fail(e, "cudaMalloc(edge ids)"); }
  if (num_edges && (e = cudaMemcpy(ix->edge_ids, edge_id, sizeof(int32_t) * num_edges, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy(edge ids)");
  return S2B_OK;
}

static int incident_prologue(const S2bIndex* ix) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (!ix->edge_ids) return set_error(S2B_EINVAL, "incident-edge queries need the edge ids: call s2b_index_set_edge_ids first");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the index lives on device %d but the current device is %d", ix->device, dev);
  return S2B_OK;
}

int s2b_incident_edges_count_dev(const S2bIndex* ix, const double* xyz_dev, size_t n, uint32_t* counts_out_dev, void* stream) {
  int rc = incident_prologue(ix);
  if (rc != S2B_OK || n == 0) return rc;
  if (!xyz_dev || !counts_out_dev) return set_error(S2B_EINVAL, "null pointer");
  const auto k = incident_edges_kernel<false>;
  k<<<grid_for_ix(n, (const void*)k), kThreads, 0, (cudaStream_t)stream>>>(ix->view, ix->edge_ids, xyz_dev, n, counts_out_dev, nullptr, nullptr, nullptr);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "incident_edges launch");
}

int s2b_incident_edges_fill_dev(const S2bIndex* ix, const double* xyz_dev, size_t n, const uint32_t* offsets_dev, int32_t* shape_ids_out_dev,
                                int32_t* edge_ids_out_dev, void* stream) {
  int rc = incident_prologue(ix);
  if (rc != S2B_OK || n == 0) return rc;
  if (!xyz_dev || !offsets_dev || !shape_ids_out_dev || !edge_ids_out_dev) return set_error(S2B_EINVAL, "null pointer");
  const auto k = incident_edges_kernel<true>;
  k<<<grid_for_ix(n, (const void*)k), kThreads, 0, (cudaStream_t)stream>>>(ix->view, ix->edge_ids, xyz_dev, n, nullptr, offsets_dev, shape_ids_out_dev, edge_ids_out_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "incident_edges launch");
}

// Host-pointer form: counts, host scan, fill (whole-array copies; this is a low-volume query).
int s2b_incident_edges(const S2bIndex* ix, const double* xyz, size_t n, uint32_t* offsets_out, int32_t* shape_ids_out, int32_t* edge_ids_out,
                       size_t capacity, size_t* total_out) {
  int rc = incident_prologue(ix);
  if (rc != S2B_OK) return rc;
  if (!offsets_out) return set_error(S2B_EINVAL, "null offsets_out");
  if (total_out) *total_out = 0;
  offsets_out[0] = 0;
  if (n == 0) return S2B_OK;
  if (!xyz) return set_error(S2B_EINVAL, "null pointer");
  void *d_xyz = nullptr, *d_off = nullptr, *d_shape = nullptr, *d_edge = nullptr;
  cudaError_t e = cudaMalloc(&d_xyz, 24 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_off, 4 * (n + 1));
  if (e == cudaSuccess) e = cudaMemcpy(d_xyz, xyz, 24 * n, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK) rc = s2b_incident_edges_count_dev(ix, (const double*)d_xyz, n, (uint32_t*)d_off, nullptr);
  if (rc == S2B_OK && (e = cudaMemcpy(offsets_out + 1, d_off, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  uint64_t total = 0;
  if (rc == S2B_OK) {
    for (size_t i = 1; i <= n; ++i) { total += offsets_out[i]; offsets_out[i] = (uint32_t)total; }
    if (total > 0xFFFFFFFFull) rc = set_error(S2B_ECAPACITY, "more than 2^32 incident edges: split the batch");
  }
  if (rc == S2B_OK && total_out) *total_out = (size_t)total;
  if (rc == S2B_OK && total > capacity) rc = set_error(S2B_ECAPACITY, "capacity %zu < %llu incident edges", capacity, (unsigned long long)total);
  if (rc == S2B_OK && total > 0) {
    if (!shape_ids_out || !edge_ids_out) rc = set_error(S2B_EINVAL, "null output arrays");
    if (rc == S2B_OK && (e = cudaMalloc(&d_shape, 4 * total)) != cudaSuccess) rc = cuda_fail(e, "cudaMalloc");
    if (rc == S2B_OK && (e = cudaMalloc(&d_edge, 4 * total)) != cudaSuccess) rc = cuda_fail(e, "cudaMalloc");
    if (rc == S2B_OK && (e = cudaMemcpy(d_off, offsets_out, 4 * n, cudaMemcpyHostToDevice)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
    if (rc == S2B_OK) rc = s2b_incident_edges_fill_dev(ix, (const double*)d_xyz, n, (const uint32_t*)d_off, (int32_t*)d_shape, (int32_t*)d_edge, nullptr);
    if (rc == S2B_OK && (e = cudaMemcpy(shape_ids_out, d_shape, 4 * total, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
    if (rc == S2B_OK && (e = cudaMemcpy(edge_ids_out, d_edge, 4 * total, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  }
  cudaFree(d_xyz); cudaFree(d_off); cudaFree(d_shape); cudaFree(d_edge);
  return rc;
}

int s2b_index_locate_cells_dev(const S2bIndex* ix, const uint64_t* cell_ids_dev, size_t n, uint8_t* relation_out_dev, int32_t* cell_out_dev, void* stream) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (n == 0) return S2B_OK;
  if (!cell_ids_dev || !relation_out_dev) return set_error(S2B_EINVAL, "null pointer");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the index lives on device %d but the current device is %d", ix->device, dev);
  locate_cells_kernel<<<grid_for_ix(n, (const void*)locate_cells_kernel), kThreads, 0, (cudaStream_t)stream>>>(ix->view, cell_ids_dev, n, relation_out_dev, cell_out_dev);
  count_launch();
  e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "locate_cells launch");
}

int s2b_index_locate_cells(const S2bIndex* ix, const uint64_t* cell_ids, size_t n, uint8_t* relation_out, int32_t* cell_out) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  if (n == 0) return S2B_OK;
  if (!cell_ids || !relation_out) return set_error(S2B_EINVAL, "null pointer");
  void *d_ids = nullptr, *d_rel = nullptr, *d_cell = nullptr;
  cudaError_t e = cudaMalloc(&d_ids, 8 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_rel, n);
  if (e == cudaSuccess) e = cudaMalloc(&d_cell, 4 * n);
  if (e == cudaSuccess) e = cudaMemcpy(d_ids, cell_ids, 8 * n, cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? s2b_index_locate_cells_dev(ix, (const uint64_t*)d_ids, n, (uint8_t*)d_rel, (int32_t*)d_cell, nullptr) : cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK && (e = cudaMemcpy(relation_out, d_rel, n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_OK && cell_out && (e = cudaMemcpy(cell_out, d_cell, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  cudaFree(d_ids); cudaFree(d_rel); cudaFree(d_cell);
  return rc;
}

#define S2B_QUERY_PROLOGUE                                                            \
  if (!ix) return set_error(S2B_EINVAL, "null index");                                \
  { int rc__ = check_model(model); if (rc__) return rc__; }                           \
  if (n == 0) return S2B_OK;                                                          \
  { int dev__ = -1;                                                                   \
    cudaError_t de__ = cudaGetDevice(&dev__);                                         \
    if (de__ != cudaSuccess) return cuda_fail(de__, "cudaGetDevice");                 \
    if (dev__ != ix->device) return set_error(S2B_EINVAL, "the index lives on device %d but the current device is %d", ix->device, dev__); }

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

int s2b_count_flags_dev(const uint8_t* flags_dev, size_t n, uint64_t* count_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!flags_dev || !count_dev) return set_error(S2B_EINVAL, "null pointer");
  if ((uintptr_t)flags_dev % 16) return set_error(S2B_EINVAL, "flags_dev must be 16-byte aligned");
  count_flags_kernel<<<grid_for_ix(n / 16 + 1, (const void*)count_flags_kernel), kThreads, 0, (cudaStream_t)stream>>>(flags_dev, n, (unsigned long long*)count_dev);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "count_flags launch");
}

int s2b_diag_force_exact_locate(int on) {
  int old = g_force_exact_locate ? 1 : 0;
  g_force_exact_locate = on != 0;
  return old;
}

int s2b_exact_stage_count(uint64_t* count_out, int reset) {
  unsigned long long v = 0, vr = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, g_exact_count, sizeof v);
  if (e == cudaSuccess) e = region_exact_stage_count(&vr, reset != 0);   // s2b_region.cu's copy of the counter
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyFromSymbol");
  if (count_out) *count_out = v + vr;
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
