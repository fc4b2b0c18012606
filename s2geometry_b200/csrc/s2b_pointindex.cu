// s2b_pointindex.cu — batched S2ClosestPointQuery over an S2PointIndex (the nearest-point /
// radius joins callers pair with point-in-polygon; doc/examples/point_index.cc:25-52 is exactly the radius join).
//   S2PointIndex<Data>::Add(point, data)                       s2point_index.h       (index = points keyed by leaf cell id)
//   S2ClosestPointQuery<Data>::FindClosestPoints(PointTarget)   s2closest_point_query.h, s2closest_point_query_base.h
//     with options.max_distance (exclusive) / max_results = 1 (FindClosestPoint)
//   distance = S1ChordAngle(p, target) = min(4, |p - target|^2)  s1chord_angle.h:352-359, compared as length2
//   candidate cells = S2Cap::GetCellUnionBound                  s2cap.cc:202-223 (4 vertex neighbours at the level whose
//                                                               cells are at least twice as wide as the cap radius)
// The result of a query is a pure function of the pairwise chord distances, so any search order gives the reference's
// answer; here the index is a cell-id-sorted point array on the device and each target scans the <= 4 id ranges of its
// covering. Exactness: the candidate test is the reference's own expression, evaluated in the same order, no FMA.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <new>

#include "../../include/s2b_capi.h"
#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_internal.h"
#include "s2b_region.cuh"   // RobustCrossProd (edge targets), S2Cell(S2CellId) bounds (cell targets)
#include "s2b_index_handle.h"   // S2bIndex: the target of a ShapeIndexTarget query

struct S2bPointIndex {
  int device = 0;
  size_t m = 0;
  void* blob = nullptr;
  const uint64_t* ids = nullptr;     // sorted leaf cell ids
  const double* pts = nullptr;       // the points in that order (3 doubles each)
  const int32_t* data = nullptr;     // original position of each sorted point (the index's "data")
  const uint32_t* first = nullptr;   // first[b] = number of sorted ids below bucket b (a cell of table_level in Hilbert order)
  int table_level = 0;
};

namespace s2b {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

namespace {

constexpr int kThreads = 256;
__device__ const HilbertTables g_hilbert_pi = kHilbert;

int grid_for_pi(size_t n, const void* kernel) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  size_t want = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)sm_count() * per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

__global__ void iota_kernel(int32_t* v, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = (int32_t)i;
}

__global__ void gather_points_kernel(const double* __restrict__ xyz, const int32_t* __restrict__ order, size_t n, double* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double* q = xyz + 3 * (size_t)order[i];
    out[3 * i] = q[0]; out[3 * i + 1] = q[1]; out[3 * i + 2] = q[2];
  }
}

struct IndexView { const uint64_t* ids; const double* pts; const int32_t* data; const uint32_t* first; uint32_t m; int table_shift; };

// first[b] for every bucket b in (bucket of sorted id p-1, bucket of sorted id p]; p == m closes the table.
__global__ void bucket_table_kernel(const uint64_t* __restrict__ ids, size_t m, int shift, uint32_t num_buckets, uint32_t* __restrict__ first) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p <= m; p += stride) {
    const uint64_t prev = p == 0 ? 0 : (ids[p - 1] >> shift) + 1;
    const uint64_t cur = p == m ? num_buckets : (ids[p] >> shift);
    for (uint64_t b = prev; b <= cur; ++b) first[b] = (uint32_t)p;
  }
}

// [first, last) of the sorted ids inside cell `c`. The bucket table (cells of table_level, ~8 points each, in Hilbert
// order) answers cells at or above its level with two loads; finer cells search inside their one bucket.
__device__ __forceinline__ void cell_range(const IndexView& ix, uint64_t c, uint32_t* first, uint32_t* last) {
  const uint64_t lo_id = cellid_range_min(c), hi_id = cellid_range_max(c);
  const uint64_t b0 = lo_id >> ix.table_shift, b1 = hi_id >> ix.table_shift;
  uint32_t lo = ix.first[b0], end = ix.first[b1 + 1];
  if (b0 != b1 || lo == end) { *first = lo; *last = end; return; }      // the cell is a union of whole buckets (or empty)
  uint32_t hi = end;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ix.ids[mid] < lo_id) lo = mid + 1; else hi = mid; }
  *first = lo;
  hi = end;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ix.ids[mid] <= hi_id) lo = mid + 1; else hi = mid; }
  *last = lo;
}

// The covering of S2Cap::GetCellUnionBound for a cap centred on the leaf cell (face, i, j): AppendVertexNeighbors(level)
// (s2cell_id.cc:514-556). level < 0: the six face cells. Returns the number of cells (3, 4 or 6).
__device__ __forceinline__ int covering_cells(int face, int i, int j, int level, const uint16_t* pos, uint64_t cells[6]) {
  if (level < 0) {
    for (int f = 0; f < 6; ++f) cells[f] = ((uint64_t)f << 61) | (1ull << 60);
    return 6;
  }
  const int kMaxSize = 1 << 30;
  const int halfsize = 1 << (30 - (level + 1));
  const int size = halfsize << 1;
  bool isame, jsame;
  int ioffset, joffset;
  if (i & halfsize) { ioffset = size; isame = (i + size) < kMaxSize; } else { ioffset = -size; isame = (i - size) >= 0; }
  if (j & halfsize) { joffset = size; jsame = (j + size) < kMaxSize; } else { joffset = -size; jsame = (j - size) >= 0; }
  auto at = [&](int ii, int jj, bool same) {
    return cellid_parent(same ? from_face_ij(face, (uint32_t)ii, (uint32_t)jj, pos) : from_face_ij_wrap(face, ii, jj, pos), level);
  };
  cells[0] = at(i, j, true);
  cells[1] = at(i + ioffset, j, isame);
  cells[2] = at(i, j + joffset, jsame);
  int n = 3;
  if (isame || jsame) cells[n++] = at(i + ioffset, j + joffset, isame && jsame);
  return n;
}

__device__ __forceinline__ double chord_length2(const P3& p, const P3& q) {   // S1ChordAngle(p, q).length2()
  const double d = norm2(sub(p, q));
  return d < 4.0 ? d : 4.0;
}

// counts[t] = |{ index points with S1ChordAngle(point, target_t) < limit }|. kGroup lanes share one target (1, 8 or 32,
// chosen by the expected number of index points per covering): the lanes stride over the candidates of each covering
// cell and add up their counts. Smaller groups keep more targets in flight per warp — a target is a serial chain
// target -> covering -> table -> candidate points of dependent loads, so the kernel is latency-bound per group.
template <int kGroup>
__global__ void __launch_bounds__(kThreads)
radius_count_kernel(const IndexView ix, const double* __restrict__ targets, const uint32_t* __restrict__ order, size_t n, int level, double limit2,
                    uint32_t* __restrict__ counts) {
  __shared__ uint16_t s_pos[1024];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_pi.pos[k];
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u;
  const unsigned sub = lane & (kGroup - 1);
  const unsigned group_mask = kGroup == 32 ? 0xffffffffu : (((1u << kGroup) - 1u) << (lane & ~(unsigned)(kGroup - 1)));
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * kThreads / kGroup;
  for (size_t slot = tid / kGroup; slot < n; slot += stride) {
    const size_t t = order ? order[slot] : slot;     // spatially sorted processing order: neighbouring lanes scan the same cells
    const double* tq = targets + 3 * t;
    const P3 q{tq[0], tq[1], tq[2]};
    double fi, fj;
    const int face = point_to_face_fij(q, &fi, &fj);
    uint64_t cells[6];
    const int nc = covering_cells(face, (int)fij_to_ij(fi), (int)fij_to_ij(fj), level, s_pos, cells);
    uint32_t first[6], last[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {                       // all table reads are issued before any candidate is touched
      first[k] = last[k] = 0;
      if (k < nc) cell_range(ix, cells[k], &first[k], &last[k]);
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      for (uint32_t c = first[k] + sub; c < last[k]; c += kGroup) {
        const double* pp = ix.pts + 3 * (size_t)c;
        if (chord_length2(P3{pp[0], pp[1], pp[2]}, q) < limit2) ++cnt;
      }
    }
    if (kGroup > 1) cnt = __reduce_add_sync(group_mask, cnt);
    if (sub == 0) counts[t] = cnt;
  }
}

// FindClosestPoint: the nearest index point within `limit2` (exclusive), or data = -1. The search starts at
// `start_level` and coarsens: a best distance found inside the covering of level L is final as soon as it is below
// the cap radius that covering is guaranteed to contain (the minimum width of a level L+1 cell); at level -1 the
// covering is the whole sphere. Ties in distance go to the smaller data value.
__global__ void __launch_bounds__(kThreads)
closest_point_kernel(const IndexView ix, const double* __restrict__ targets, const uint32_t* __restrict__ order, size_t n, int start_level, double limit2,
                     const double* __restrict__ covered_length2 /* [31]: entry L+1 = chord^2 of the cap radius covered at level L */,
                     int32_t* __restrict__ data_out, double* __restrict__ length2_out) {
  __shared__ uint16_t s_pos[1024];
  __shared__ double s_cov[31];
  for (int k = threadIdx.x; k < 1024; k += kThreads) s_pos[k] = g_hilbert_pi.pos[k];
  if (threadIdx.x < 31) s_cov[threadIdx.x] = covered_length2[threadIdx.x];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t slot = (size_t)blockIdx.x * kThreads + threadIdx.x; slot < n; slot += stride) {
    const size_t t = order ? order[slot] : slot;
    const double* tq = targets + 3 * t;
    const P3 q{tq[0], tq[1], tq[2]};
    double fi, fj;
    const int face = point_to_face_fij(q, &fi, &fj);
    const int i = (int)fij_to_ij(fi), j = (int)fij_to_ij(fj);
    double best = limit2;
    int32_t best_data = -1;
    for (int level = start_level;; --level) {
      uint64_t cells[6];
      const int nc = covering_cells(face, i, j, level, s_pos, cells);
      best = limit2;
      best_data = -1;
      for (int k = 0; k < nc; ++k) {
        uint32_t first, last;
        cell_range(ix, cells[k], &first, &last);
        for (uint32_t c = first; c < last; ++c) {
          const double* pp = ix.pts + 3 * (size_t)c;
          const double d = chord_length2(P3{pp[0], pp[1], pp[2]}, q);
          const int32_t dv = ix.data[c];
          if (d < best || (d == best && best_data >= 0 && dv < best_data)) { best = d; best_data = dv; }
        }
      }
      if (level < 0) break;                                   // whole sphere scanned
      const double covered = s_cov[level + 1];
      if (best_data >= 0 && best < covered) break;            // nothing outside the covering can be closer
      if (covered >= limit2) break;                           // everything within the distance limit was covered
    }
    data_out[t] = best_data;
    if (length2_out) length2_out[t] = best_data >= 0 ? best : HUGE_VAL;
  }
}

// FindClosestPoints with options.max_results = k (1..kMaxResults) and max_distance: the k nearest index points of every
// target within the limit, nearest first (ties in distance: smaller data first). Same widening search as above, with the
// k-th best distance in place of the best. Row t of the outputs holds count[t] results, padded with -1 / +inf.
constexpr int kMaxResults = 32;
__global__ void __launch_bounds__(kThreads)
closest_points_kernel(const IndexView ix, const double* __restrict__ targets, const uint32_t* __restrict__ order, size_t n, int k, int start_level, double limit2,
                      const double* __restrict__ covered_length2, int32_t* __restrict__ data_out, double* __restrict__ length2_out,
                      uint32_t* __restrict__ count_out) {
  __shared__ uint16_t s_pos[1024];
  __shared__ double s_cov[31];
  for (int q = threadIdx.x; q < 1024; q += kThreads) s_pos[q] = g_hilbert_pi.pos[q];
  if (threadIdx.x < 31) s_cov[threadIdx.x] = covered_length2[threadIdx.x];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t slot = (size_t)blockIdx.x * kThreads + threadIdx.x; slot < n; slot += stride) {
    const size_t t = order ? order[slot] : slot;
    const double* tq = targets + 3 * t;
    const P3 q{tq[0], tq[1], tq[2]};
    double fi, fj;
    const int face = point_to_face_fij(q, &fi, &fj);
    const int i = (int)fij_to_ij(fi), j = (int)fij_to_ij(fj);
    double dist[kMaxResults];
    int32_t data[kMaxResults];
    int cnt = 0;
    for (int level = start_level;; --level) {
      uint64_t cells[6];
      const int nc = covering_cells(face, i, j, level, s_pos, cells);
      cnt = 0;
      for (int c6 = 0; c6 < nc; ++c6) {
        uint32_t first, last;
        cell_range(ix, cells[c6], &first, &last);
        for (uint32_t c = first; c < last; ++c) {
          const double* pp = ix.pts + 3 * (size_t)c;
          const double d = chord_length2(P3{pp[0], pp[1], pp[2]}, q);
          if (!(d < limit2)) continue;
          const int32_t dv = ix.data[c];
          if (cnt == k && !(d < dist[k - 1] || (d == dist[k - 1] && dv < data[k - 1]))) continue;
          int pos = cnt < k ? cnt : k - 1;                    // insertion into the sorted prefix
          while (pos > 0 && (d < dist[pos - 1] || (d == dist[pos - 1] && dv < data[pos - 1]))) {
            dist[pos] = dist[pos - 1]; data[pos] = data[pos - 1];
            --pos;
          }
          dist[pos] = d; data[pos] = dv;
          if (cnt < k) ++cnt;
        }
      }
      if (level < 0) break;
      const double covered = s_cov[level + 1];
      if (cnt == k && dist[k - 1] < covered) break;           // the k-th nearest lies inside the guaranteed radius
      if (covered >= limit2) break;                           // everything within the limit was covered
    }
    for (int r = 0; r < k; ++r) {
      data_out[t * (size_t)k + r] = r < cnt ? data[r] : -1;
      if (length2_out) length2_out[t * (size_t)k + r] = r < cnt ? dist[r] : HUGE_VAL;
    }
    if (count_out) count_out[t] = (uint32_t)cnt;
  }
}

// ---------------------------------------------------------------------------------------------------
// Targets other than points (S2ClosestPointQuery::EdgeTarget / CellTarget, s2closest_point_query.h:110-131).
// The distance from an index point to the target is the reference's own expression:
//   edge   S2::UpdateMinDistance(p, a, b, &min_dist)   s2edge_distances.cc:93-223 (S2MinDistanceEdgeTarget, s2min_distance_targets.cc:79-82)
//   cell   S2Cell(id).GetDistance(p)                    s2cell.cc:335-436          (S2MinDistanceCellTarget, :114-117)
// A query's result is a function of those distances only, so the search order is free: every target scans the
// S2Cap::GetCellUnionBound-style covering (covering_cells) of a cap around the target's bounding cap.
enum TargetKind { kTargetPoint = 0, kTargetEdge = 1, kTargetCell = 2 };

// AlwaysUpdateMinDistance<false>: true (and *dist2 = the S1ChordAngle length2 of the distance) iff it is below limit2.
// The interior case's early exits compare against the limit exactly as the reference does.
__device__ __forceinline__ bool edge_update_min_distance(const P3& x, const P3& a, const P3& b, double limit2, double* dist2) {
  const double xa2 = norm2(sub(x, a)), xb2 = norm2(sub(x, b));
  const double ab2 = norm2(sub(a, b));
  const double max_error = (4.75 * DBL_EPSILON) * ((xa2 + xb2) + ab2) + (8 * DBL_EPSILON * DBL_EPSILON);
  if (!(fabs(xa2 - xb2) >= ab2 + max_error)) {
    const P3 c = robust_cross_prod(a, b);
    const double c2 = norm2(c);
    const double x_dot_c = dot(x, c);
    const double x_dot_c2 = x_dot_c * x_dot_c;
    if (!(x_dot_c2 > c2 * limit2)) {
      const P3 cx = cross(c, x);
      if (!(dot(sub(a, x), cx) >= 0 || dot(sub(b, x), cx) <= 0)) {
        const double qr = 1 - sqrt(norm2(cx) / c2);
        const double d2 = (x_dot_c2 / c2) + (qr * qr);
        if (!(d2 >= limit2)) { *dist2 = d2 < 4.0 ? d2 : 4.0; return true; }
      }
    }
  }
  const double v = xa2 < xb2 ? xa2 : xb2;      // std::min(xa2, xb2)
  if (v >= limit2) return false;
  *dist2 = v < 4.0 ? v : 4.0;
  return true;
}

// S2Cell::GetDistance(const S2Point&) = GetDistanceInternal(target, to_interior = true) as a length2.
__device__ __forceinline__ double cell_edge_distance2(double dir, double uv) {     // EdgeDistance (s2cell.cc:366-380)
  const double pq2 = (dir * dir) / (1 + uv * uv);
  const double qr = 1 - sqrt(1 - pq2);
  const double d = pq2 + qr * qr;
  return d < 4.0 ? d : 4.0;
}
__device__ __forceinline__ double cell_vertex_distance2(const P3& p, double u, double v) {   // VertexChordDist (:335-339)
  const double d = norm2(sub(p, normalize(P3{u, v, 1.0})));
  return d < 4.0 ? d : 4.0;
}
__device__ __forceinline__ double cell_point_distance2(const TargetCell& cell, const P3& xyz) {
  const P3 t = face_xyz_to_uvw(cell.face, xyz);
  const double u0 = cell.bound.ulo, u1 = cell.bound.uhi, v0 = cell.bound.vlo, v1 = cell.bound.vhi;
  const double dir00 = t.x - t.z * u0, dir01 = t.x - t.z * u1, dir10 = t.y - t.z * v0, dir11 = t.y - t.z * v1;
  // VEdgeIsClosest(p, u_end) / UEdgeIsClosest(p, v_end) (:344-362)
  auto v_edge_closest = [&](double u) { return dot(t, P3{-u * v0, u * u + 1, -v0}) > 0 && dot(t, P3{-u * v1, u * u + 1, -v1}) < 0; };
  auto u_edge_closest = [&](double v) { return dot(t, P3{v * v + 1, -u0 * v, -u0}) > 0 && dot(t, P3{v * v + 1, -u1 * v, -u1}) < 0; };
  bool inside = true;
  if (dir00 < 0) { inside = false; if (v_edge_closest(u0)) return cell_edge_distance2(-dir00, u0); }
  if (dir01 > 0) { inside = false; if (v_edge_closest(u1)) return cell_edge_distance2(dir01, u1); }
  if (dir10 < 0) { inside = false; if (u_edge_closest(v0)) return cell_edge_distance2(-dir10, v0); }
  if (dir11 > 0) { inside = false; if (u_edge_closest(v1)) return cell_edge_distance2(dir11, v1); }
  if (inside) return 0.0;
  const double a = cell_vertex_distance2(t, u0, v0), b = cell_vertex_distance2(t, u1, v0);
  const double c = cell_vertex_distance2(t, u0, v1), d = cell_vertex_distance2(t, u1, v1);
  const double ab = b < a ? b : a, cd = d < c ? d : c;          // std::min(x, y) = (y < x) ? y : x
  return cd < ab ? cd : ab;
}

struct TargetGeom {
  P3 a, b;            // point: a; edge: a, b
  TargetCell cell;    // cell
  P3 center;          // of a cap that bounds the target
  double rho;         // its radius in radians, rounded up (only steers the search: any valid bound gives the same results)
};

template <int kKind>
__device__ __forceinline__ TargetGeom load_target(const void* targets, size_t t, const uint16_t* ij_table) {
  TargetGeom g{};
  if (kKind == kTargetPoint) {
    const double* q = (const double*)targets + 3 * t;
    g.a = P3{q[0], q[1], q[2]};
    g.center = g.a; g.rho = 0.0;
  } else if (kKind == kTargetEdge) {
    const double* q = (const double*)targets + 6 * t;
    g.a = P3{q[0], q[1], q[2]}; g.b = P3{q[3], q[4], q[5]};
    const P3 m = add(g.a, g.b);
    const double n2 = norm2(m);
    if (n2 > 1e-20) {
      g.center = normalize(m);
      const double d2 = fmax(norm2(sub(g.center, g.a)), norm2(sub(g.center, g.b)));
      g.rho = 2.0 * asin(fmin(1.0, 0.5 * sqrt(d2))) * (1 + 1e-9) + 1e-15;
    } else {          // (nearly) antipodal endpoints: no useful bound, search the whole sphere
      g.center = g.a; g.rho = 4.0;
    }
  } else {
    const uint64_t id = ((const uint64_t*)targets)[t];
    g.cell = target_cell(id, ij_table);
    g.center = cellid_to_point(id, ij_table);
    double d2 = 0.0;
    const P3 c = face_xyz_to_uvw(g.cell.face, g.center);
    d2 = fmax(fmax(cell_vertex_distance2(c, g.cell.bound.ulo, g.cell.bound.vlo), cell_vertex_distance2(c, g.cell.bound.uhi, g.cell.bound.vlo)),
              fmax(cell_vertex_distance2(c, g.cell.bound.ulo, g.cell.bound.vhi), cell_vertex_distance2(c, g.cell.bound.uhi, g.cell.bound.vhi)));
    g.rho = 2.0 * asin(fmin(1.0, 0.5 * sqrt(d2))) * (1 + 1e-9) + 1e-15;
  }
  return g;
}

template <int kKind>
__device__ __forceinline__ bool target_distance(const TargetGeom& g, const P3& p, double limit2, double* dist2) {
  if (kKind == kTargetPoint) { const double d = chord_length2(p, g.a); if (!(d < limit2)) return false; *dist2 = d; return true; }
  if (kKind == kTargetEdge) return edge_update_min_distance(p, g.a, g.b, limit2, dist2);
  const double d = cell_point_distance2(g.cell, p);
  if (!(d < limit2)) return false;
  *dist2 = d;
  return true;
}

// FindClosestPoints for any target kind. k == 0: count every index point within the limit (options.max_results unlimited,
// counts only); k >= 1: the k nearest within the limit, nearest first (ties: smaller data first), rows padded with -1 / +inf.
// One thread per target. The covering widens level by level (as closest_points_kernel) from the finest level whose
// covering is guaranteed to contain the target's bounding cap; a level's result is final once the k-th distance lies
// within what that covering guarantees for the target (covered angle minus the bounding cap's radius), or once the
// covering contains everything within the distance limit.
template <int kKind>
__global__ void __launch_bounds__(kThreads)
closest_targets_kernel(const IndexView ix, const void* __restrict__ targets, size_t n, int k, int start_level, double limit2, double limit_angle,
                       const double* __restrict__ covered_angle /* [31]: entry L+1 = angle of the cap covered at level L */,
                       int32_t* __restrict__ data_out, double* __restrict__ length2_out, uint32_t* __restrict__ count_out,
                       unsigned long long* __restrict__ min_bits /* nullable, k == 0: min_bits[c] = min(min_bits[c], bits of the distance) */,
                       int lane_shift /* with min_bits: 2^lane_shift neighbouring lanes share one target and split its candidates */) {
  __shared__ uint16_t s_pos[1024], s_ij[1024];
  __shared__ double s_cov[31];
  for (int q = threadIdx.x; q < 1024; q += kThreads) { s_pos[q] = g_hilbert_pi.pos[q]; s_ij[q] = g_hilbert_pi.ij[q]; }
  if (threadIdx.x < 31) s_cov[threadIdx.x] = covered_angle[threadIdx.x];
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  const uint32_t sub_step = 1u << lane_shift;
  for (size_t gt = (size_t)blockIdx.x * kThreads + threadIdx.x; (gt >> lane_shift) < n; gt += stride) {
    const size_t t = gt >> lane_shift;
    const uint32_t lane_off = (uint32_t)gt & (sub_step - 1u);
    const TargetGeom g = load_target<kKind>(targets, t, s_ij);
    double fi, fj;
    const int face = point_to_face_fij(g.center, &fi, &fj);
    const int i = (int)fij_to_ij(fi), j = (int)fij_to_ij(fj);
    // finest level whose covering contains the bounding cap itself
    int level = start_level;
    while (level >= 0 && !(s_cov[level + 1] > g.rho)) --level;
    const double need = limit_angle + g.rho;     // the covering that contains everything within the limit
    const double rc = g.rho >= 3.14 ? 4.0 : 2.0 * sin(0.5 * g.rho) * (1 + 1e-9);   // chord length of the bounding cap's radius
    if (k == 0) { while (level >= 0 && !(s_cov[level + 1] >= need)) --level; }
    double dist[kMaxResults];
    int32_t data[kMaxResults];
    int cnt = 0;
    uint32_t total = 0;
    for (;; --level) {
      uint64_t cells[6];
      const int nc = covering_cells(face, i, j, level, s_pos, cells);
      cnt = 0; total = 0;
      for (int c6 = 0; c6 < nc; ++c6) {
        uint32_t first, last;
        cell_range(ix, cells[c6], &first, &last);
        for (uint32_t c = first + lane_off; c < last; c += sub_step) {
          const double* pp = ix.pts + 3 * (size_t)c;
          const P3 p{pp[0], pp[1], pp[2]};
          double d;
          if (min_bits) {
            // Per-point minimum over many targets: the point's current best (possibly stale: that only weakens the pruning)
            // bounds what this target still has to beat, as S2ClosestEdgeQuery's shrinking max_distance does. A target
            // whose bounding ball — centre g.center, chord radius rc — is further away than that is skipped after ~12
            // operations: |p - x| >= |p - centre| - rc for every point x of the target.
            const double cur = __longlong_as_double((long long)min_bits[c]);
            const double beat = cur < limit2 ? cur : limit2;
            const double dc2 = norm2(sub(p, g.center));
            const double reach = rc + (double)(__fsqrt_ru((float)beat) * 1.000001f) + 1e-12;
            if (beat < 4.0 && dc2 * (1 - 1e-6) >= reach * reach) continue;
            if (!target_distance<kKind>(g, p, beat, &d)) continue;
            ++total;
            atomicMin(min_bits + c, (unsigned long long)__double_as_longlong(d));   // d >= 0: bit order = numeric order
            continue;
          }
          // with a full result set the reference's limit is the k-th distance; ties are then resolved on the data value below
          if (!target_distance<kKind>(g, p, limit2, &d)) continue;
          ++total;
          if (k == 0) continue;
          const int32_t dv = ix.data[c];
          if (cnt == k && !(d < dist[k - 1] || (d == dist[k - 1] && dv < data[k - 1]))) continue;
          int pos = cnt < k ? cnt : k - 1;
          while (pos > 0 && (d < dist[pos - 1] || (d == dist[pos - 1] && dv < data[pos - 1]))) {
            dist[pos] = dist[pos - 1]; data[pos] = data[pos - 1];
            --pos;
          }
          dist[pos] = d; data[pos] = dv;
          if (cnt < k) ++cnt;
        }
      }
      if (level < 0) break;                                      // whole sphere scanned
      const double covered = s_cov[level + 1];
      if (covered >= need) break;                                // everything within the limit was covered
      if (k > 0 && cnt == k) {
        const double guaranteed = covered - g.rho;               // every point within this angle of the target was scanned
        if (guaranteed > 0) {
          const double h = sin(0.5 * guaranteed);
          if (dist[k - 1] < 4.0 * h * h * (1 - 1e-9)) break;
        }
      }
    }
    if (count_out) count_out[t] = k == 0 ? total : (uint32_t)cnt;
    for (int r = 0; r < k; ++r) {
      if (data_out) data_out[t * (size_t)k + r] = r < cnt ? data[r] : -1;
      if (length2_out) length2_out[t * (size_t)k + r] = r < cnt ? dist[r] : HUGE_VAL;
    }
  }
}

// ShapeIndexTarget helpers: per-index-point distance slots.
constexpr unsigned long long kInfBits = 0x7FF0000000000000ull;
__global__ void fill_u64_kernel(unsigned long long* v, size_t n, unsigned long long value) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = value;
}
__global__ void zero_where_flag_kernel(unsigned long long* v, const uint8_t* __restrict__ flags, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) v[i] = 0ull;
}
// keys sorted ascending: *total = number of keys below `limit`
__global__ void count_below_sorted_kernel(const unsigned long long* __restrict__ keys, size_t n, unsigned long long limit, unsigned long long* total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keys[i] < limit && (i + 1 == n || keys[i + 1] >= limit)) *total = i + 1;
}

// out[data[c]] = the distance recorded for sorted point c (+inf where nothing within the limit was found)
__global__ void scatter_distances_kernel(const unsigned long long* __restrict__ bits, const int32_t* __restrict__ data, size_t n,
                                         unsigned long long limit_bits, double* __restrict__ out) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) out[data[c]] = bits[c] < limit_bits ? __longlong_as_double((long long)bits[c]) : HUGE_VAL;
}

// kMinWidth (quadratic projection, s2metrics.cc:54-58) and Metric::GetLevelForMinValue (s2metrics.h:185-200).
const double kMinWidthDeriv = 2 * std::sqrt(2.0) / 3;
int level_for_min_width(double value) {
  if (!(value > 0)) return 30;
  int level = std::ilogb(kMinWidthDeriv / value);
  return level < 0 ? 0 : (level > 30 ? 30 : level);
}
// angle of a chord^2, slightly inflated so that the covering level is chosen conservatively
double angle_of_length2(double l2) {
  if (l2 >= 4.0) return M_PI;
  return 2 * std::asin(0.5 * std::sqrt(l2)) * (1 + 1e-9) + 1e-300;
}

// Processing order for a large batch of targets: their indices sorted by leaf cell id (K1 + radix sort on the top bits),
// so that the threads of a warp work on neighbouring targets — the same covering cells, the same candidate points (L1 hits
// and broadcasts instead of 32 unrelated streams). The results do not depend on the order. *blob_out must be released
// with cudaFreeAsync on the same stream after the consumer kernel has been enqueued. Small batches keep their order.
constexpr size_t kSortTargetsFrom = size_t(1) << 18;
cudaError_t sorted_target_order(const double* targets_dev, size_t n, cudaStream_t st, const uint32_t** order_out, void** blob_out) {
  *order_out = nullptr;
  *blob_out = nullptr;
  if (n < kSortTargetsFrom || n > 0x7FFFFFFFull) return cudaSuccess;
  size_t t_sort = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 32, 64, st);
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_ids2 = up(8 * n), o_iota = 2 * up(8 * n), o_order = o_iota + up(4 * n), o_tmp = o_order + up(4 * n);
  char* blob = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&blob, o_tmp + up(t_sort), st);
  if (e != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }      // no memory for the sort: keep the caller's order
  EncodeOut eo{};
  eo.levels.n = 1; eo.levels.level[0] = 30;
  eo.ids = (uint64_t*)blob;
  eo.token_level_index = -1;
  e = launch_encode(kInXYZ, targets_dev, n, eo, st);
  if (e == cudaSuccess) {
    iota_kernel<<<grid_for_pi(n, (const void*)iota_kernel), kThreads, 0, st>>>((int32_t*)(blob + o_iota), n);
    count_launch();
    // face + the top 14 levels (bits 63..32) order the targets finely enough; half the radix passes of a full sort
    e = cub::DeviceRadixSort::SortPairs(blob + o_tmp, t_sort, (const uint64_t*)blob, (uint64_t*)(blob + o_ids2), (const uint32_t*)(blob + o_iota),
                                        (uint32_t*)(blob + o_order), (int)n, 32, 64, st);
    count_launch();
  }
  if (e != cudaSuccess) { cudaFreeAsync(blob, st); return e; }
  *order_out = (const uint32_t*)(blob + o_order);
  *blob_out = blob;
  return cudaSuccess;
}

}  // namespace
}  // namespace s2b

using namespace s2b;

extern "C" {

int s2b_point_index_create(const double* xyz, size_t m, S2bPointIndex** out) {
  if (!out) return set_error(S2B_EINVAL, "null output");
  *out = nullptr;
  if (m && !xyz) return set_error(S2B_EINVAL, "null points");
  if (m > 0x7FFFFFF0ull) return set_error(S2B_EINVAL, "point index too large (max 2^31 points)");
  S2bPointIndex* ix = new (std::nothrow) S2bPointIndex();
  if (!ix) return set_error(S2B_ENOMEM, "host allocation failed");
  cudaError_t e = cudaGetDevice(&ix->device);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaGetDevice"); }
  ix->m = m;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  // bucket table level: ~8 points per bucket, between level 0 (6 buckets) and level 12 (100M buckets)
  int table_level = 0;
  while (table_level < 12 && (6ull << (2 * (table_level + 1))) * 8 <= m) ++table_level;
  ix->table_level = table_level;
  const uint32_t num_buckets = (uint32_t)(6ull << (2 * table_level));
  const size_t b_ids = up(8 * (m + 1)), b_pts = up(24 * (m + 1)), b_data = up(4 * (m + 1)), b_first = up(4 * ((size_t)num_buckets + 2));
  e = cudaMalloc(&ix->blob, b_ids + b_pts + b_data + b_first);
  if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc(point index)"); }
  char* base = (char*)ix->blob;
  ix->ids = (const uint64_t*)base;
  ix->pts = (const double*)(base + b_ids);
  ix->data = (const int32_t*)(base + b_ids + b_pts);
  ix->first = (const uint32_t*)(base + b_ids + b_pts + b_data);
  if (m == 0) {
    e = cudaMemset((void*)ix->first, 0, 4 * ((size_t)num_buckets + 2));
    if (e != cudaSuccess) { cudaFree(ix->blob); delete ix; return cuda_fail(e, "cudaMemset"); }
    *out = ix;
    return S2B_OK;
  }
  // temporaries: raw points, unsorted ids, iota, sort scratch
  size_t t_sort = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)m);
  char* tmp = nullptr;
  const size_t o_ids = up(24 * m), o_iota = o_ids + up(8 * m), o_sort = o_iota + up(4 * m);
  e = cudaMalloc((void**)&tmp, o_sort + up(t_sort));
  int rc = S2B_OK;
  if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(point index build)");
  if (rc == S2B_OK && (e = cudaMemcpy(tmp, xyz, 24 * m, cudaMemcpyDefault)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(points)");
  if (rc == S2B_OK) {
    EncodeOut eo{};
    eo.levels.n = 1; eo.levels.level[0] = 30;
    eo.ids = (uint64_t*)(tmp + o_ids);
    eo.flags = nullptr; eo.token_level_index = -1; eo.tokens16 = nullptr; eo.token_len = nullptr;
    e = launch_encode(kInXYZ, tmp, m, eo, nullptr);                      // S2CellId(point) for every index point (K1)
    if (e == cudaSuccess) {
      iota_kernel<<<grid_for_pi(m, (const void*)iota_kernel), kThreads>>>((int32_t*)(tmp + o_iota), m);
      count_launch();
      e = cub::DeviceRadixSort::SortPairs(tmp + o_sort, t_sort, (const uint64_t*)(tmp + o_ids), (uint64_t*)ix->ids,
                                          (const int32_t*)(tmp + o_iota), (int32_t*)ix->data, (int)m);
      count_launch();
    }
    if (e == cudaSuccess) {
      bucket_table_kernel<<<grid_for_pi(m + 1, (const void*)bucket_table_kernel), kThreads>>>(ix->ids, m, 61 - 2 * table_level, num_buckets, (uint32_t*)ix->first);
      count_launch();
      gather_points_kernel<<<grid_for_pi(m, (const void*)gather_points_kernel), kThreads>>>((const double*)tmp, ix->data, m, (double*)ix->pts);
      count_launch();
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "point index build");
  }
  if (tmp) cudaFree(tmp);
  if (rc != S2B_OK) { cudaFree(ix->blob); delete ix; return rc; }
  *out = ix;
  return S2B_OK;
}

void s2b_point_index_destroy(S2bPointIndex* ix) {
  if (!ix) return;
  if (ix->blob) cudaFree(ix->blob);
  delete ix;
}

size_t s2b_point_index_size(const S2bPointIndex* ix) { return ix ? ix->m : 0; }

static int check_index_device(const S2bPointIndex* ix) {
  if (!ix) return set_error(S2B_EINVAL, "null point index");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != ix->device) return set_error(S2B_EINVAL, "the point index lives on device %d but the current device is %d", ix->device, dev);
  return S2B_OK;
}

int s2b_closest_points_count_dev(const S2bPointIndex* ix, const double* targets_dev, size_t n, double max_length2, uint32_t* counts_out_dev, void* stream) {
  int rc = check_index_device(ix);
  if (rc != S2B_OK) return rc;
  if (n == 0) return S2B_OK;
  if (!targets_dev || !counts_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if (std::isnan(max_length2)) return set_error(S2B_EINVAL, "max_length2 is NaN");
  cudaStream_t st = (cudaStream_t)stream;
  if (ix->m == 0 || !(max_length2 > 0)) {            // nothing is at a distance < 0 (S1ChordAngle::Negative/Zero limits)
    cudaError_t e = cudaMemsetAsync(counts_out_dev, 0, 4 * n, st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "cudaMemsetAsync");
  }
  const int level = level_for_min_width(angle_of_length2(max_length2)) - 1;
  const IndexView v{ix->ids, ix->pts, ix->data, ix->first, (uint32_t)ix->m, 61 - 2 * ix->table_level};
  // expected index points per covering: 4 cells of `level` out of 6 * 4^level
  const double per_cover = level < 0 ? (double)ix->m : 4.0 * (double)ix->m / (6.0 * std::ldexp(1.0, 2 * level));
  // measured on B200 (10M index points, ~100 per covering, 50M targets): 957 / 848 / 365 M targets/s with 1 / 8 / 32
  // lanes per target, so lanes are only shared once a covering holds far more points than one thread should scan
  const int group = per_cover >= 8192.0 ? 32 : (per_cover >= 1024.0 ? 8 : 1);
  const uint32_t* order = nullptr;
  void* order_blob = nullptr;
  cudaError_t e = sorted_target_order(targets_dev, n, st, &order, &order_blob);
  if (e != cudaSuccess) return cuda_fail(e, "target sort");
  if (group == 32) {
    const auto k = radius_count_kernel<32>;
    k<<<grid_for_pi(n * 32, (const void*)k), kThreads, 0, st>>>(v, targets_dev, order, n, level, max_length2, counts_out_dev);
  } else if (group == 8) {
    const auto k = radius_count_kernel<8>;
    k<<<grid_for_pi(n * 8, (const void*)k), kThreads, 0, st>>>(v, targets_dev, order, n, level, max_length2, counts_out_dev);
  } else {
    const auto k = radius_count_kernel<1>;
    k<<<grid_for_pi(n, (const void*)k), kThreads, 0, st>>>(v, targets_dev, order, n, level, max_length2, counts_out_dev);
  }
  count_launch();
  e = cudaGetLastError();
  if (order_blob) cudaFreeAsync(order_blob, st);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "radius_count launch");
}

int s2b_closest_point_dev(const S2bPointIndex* ix, const double* targets_dev, size_t n, double max_length2, int32_t* data_out_dev,
                          double* length2_out_dev, void* stream) {
  int rc = check_index_device(ix);
  if (rc != S2B_OK) return rc;
  if (n == 0) return S2B_OK;
  if (!targets_dev || !data_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if (std::isnan(max_length2)) return set_error(S2B_EINVAL, "max_length2 is NaN");
  cudaStream_t st = (cudaStream_t)stream;
  // chord^2 of the cap radius the covering of level L is guaranteed to contain: min width of a level L+1 cell, shrunk a
  // little (entry 0 = level -1 = the whole sphere)
  double cov[31];
  cov[0] = HUGE_VAL;
  for (int L = 0; L < 30; ++L) {
    const double r = std::ldexp(kMinWidthDeriv, -(L + 1)) * (1 - 1e-9);
    const double len = 2 * std::sin(0.5 * r);
    cov[L + 1] = len * len * (1 - 1e-9);
  }
  double* cov_dev = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&cov_dev, sizeof(cov), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  e = cudaMemcpyAsync(cov_dev, cov, sizeof(cov), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);                    // cov[] is a stack array
  if (e != cudaSuccess) { cudaFreeAsync(cov_dev, st); return cuda_fail(e, "cudaMemcpyAsync"); }
  // start where a covering holds ~16 index points on average
  int start = -1;
  if (ix->m > 0) {
    start = (int)std::floor(0.5 * std::log2((double)ix->m * 4.0 / (6.0 * 16.0)));
    start = start < -1 ? -1 : (start > 29 ? 29 : start);
  }
  const IndexView v{ix->ids, ix->pts, ix->data, ix->first, (uint32_t)ix->m, 61 - 2 * ix->table_level};
  const uint32_t* order = nullptr;
  void* order_blob = nullptr;
  e = sorted_target_order(targets_dev, n, st, &order, &order_blob);
  if (e != cudaSuccess) { cudaFreeAsync(cov_dev, st); return cuda_fail(e, "target sort"); }
  closest_point_kernel<<<grid_for_pi(n, (const void*)closest_point_kernel), kThreads, 0, st>>>(v, targets_dev, order, n, start, max_length2, cov_dev, data_out_dev, length2_out_dev);
  count_launch();
  e = cudaGetLastError();
  if (order_blob) cudaFreeAsync(order_blob, st);
  cudaFreeAsync(cov_dev, st);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "closest_point launch");
}

int s2b_closest_points_dev(const S2bPointIndex* ix, const double* targets_dev, size_t n, int max_results, double max_length2,
                           int32_t* data_out_dev, double* length2_out_dev, uint32_t* count_out_dev, void* stream) {
  int rc = check_index_device(ix);
  if (rc != S2B_OK) return rc;
  if (max_results < 1 || max_results > kMaxResults) return set_error(S2B_EINVAL, "max_results must be in [1, %d]", kMaxResults);
  if (n == 0) return S2B_OK;
  if (!targets_dev || !data_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if (std::isnan(max_length2)) return set_error(S2B_EINVAL, "max_length2 is NaN");
  cudaStream_t st = (cudaStream_t)stream;
  double cov[31];
  cov[0] = HUGE_VAL;
  for (int L = 0; L < 30; ++L) {
    const double r = std::ldexp(kMinWidthDeriv, -(L + 1)) * (1 - 1e-9);
    const double len = 2 * std::sin(0.5 * r);
    cov[L + 1] = len * len * (1 - 1e-9);
  }
  double* cov_dev = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&cov_dev, sizeof(cov), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  e = cudaMemcpyAsync(cov_dev, cov, sizeof(cov), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cudaFreeAsync(cov_dev, st); return cuda_fail(e, "cudaMemcpyAsync"); }
  int start = -1;   // a covering that holds ~4k + 16 index points on average
  if (ix->m > 0) {
    start = (int)std::floor(0.5 * std::log2((double)ix->m * 4.0 / (6.0 * (16.0 + 4.0 * max_results))));
    start = start < -1 ? -1 : (start > 29 ? 29 : start);
  }
  const IndexView v{ix->ids, ix->pts, ix->data, ix->first, (uint32_t)ix->m, 61 - 2 * ix->table_level};
  const uint32_t* order = nullptr;
  void* order_blob = nullptr;
  e = sorted_target_order(targets_dev, n, st, &order, &order_blob);
  if (e != cudaSuccess) { cudaFreeAsync(cov_dev, st); return cuda_fail(e, "target sort"); }
  closest_points_kernel<<<grid_for_pi(n, (const void*)closest_points_kernel), kThreads, 0, st>>>(v, targets_dev, order, n, max_results, start, max_length2, cov_dev,
                                                                                               data_out_dev, length2_out_dev, count_out_dev);
  count_launch();
  e = cudaGetLastError();
  if (order_blob) cudaFreeAsync(order_blob, st);
  cudaFreeAsync(cov_dev, st);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "closest_points launch");
}

// Launches closest_targets_kernel; min_bits != nullptr (with max_results == 0) makes it record, per index point, the
// smallest distance to any of the targets instead of per-target results.
static int launch_closest_targets(const S2bPointIndex* ix, int target_kind, const void* targets_dev, size_t n, int max_results, double max_length2,
                                  int32_t* data_out_dev, double* length2_out_dev, uint32_t* count_out_dev, unsigned long long* min_bits,
                                  cudaStream_t st);

int s2b_closest_points_to_targets_dev(const S2bPointIndex* ix, int target_kind, const void* targets_dev, size_t n, int max_results,
                                      double max_length2, double max_error_radians, int32_t* data_out_dev, double* length2_out_dev,
                                      uint32_t* count_out_dev, void* stream) {
  int rc = check_index_device(ix);
  if (rc != S2B_OK) return rc;
  if (target_kind < 0 || target_kind > 2) return set_error(S2B_EINVAL, "target_kind must be S2B_TARGET_POINT, _EDGE or _CELL");
  if (max_results < 0 || max_results > kMaxResults) return set_error(S2B_EINVAL, "max_results must be in [0, %d] (0 = count only)", kMaxResults);
  if (std::isnan(max_length2) || !(max_error_radians >= 0)) return set_error(S2B_EINVAL, "bad max_length2 / max_error");
  if (n == 0) return S2B_OK;
  if (!targets_dev || (max_results > 0 && !data_out_dev) || (max_results == 0 && !count_out_dev)) return set_error(S2B_EINVAL, "null pointer");
  return launch_closest_targets(ix, target_kind, targets_dev, n, max_results, max_length2, data_out_dev, length2_out_dev, count_out_dev, nullptr,
                                (cudaStream_t)stream);
}

static int launch_closest_targets(const S2bPointIndex* ix, int target_kind, const void* targets_dev, size_t n, int max_results, double max_length2,
                                  int32_t* data_out_dev, double* length2_out_dev, uint32_t* count_out_dev, unsigned long long* min_bits,
                                  cudaStream_t st) {
  // max_error: results up to that much further away than the true ones MAY be substituted (s2closest_point_query_base.h:
  // 82-101); the results computed here are exact, which satisfies every max_error >= 0.
  double cov[31];
  cov[0] = HUGE_VAL;
  for (int L = 0; L < 30; ++L) cov[L + 1] = std::ldexp(kMinWidthDeriv, -(L + 1)) * (1 - 1e-9);
  double* cov_dev = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&cov_dev, sizeof(cov), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  e = cudaMemcpyAsync(cov_dev, cov, sizeof(cov), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);                    // cov[] is a stack array
  if (e != cudaSuccess) { cudaFreeAsync(cov_dev, st); return cuda_fail(e, "cudaMemcpyAsync"); }
  int start = -1;   // a covering that holds ~4k + 16 index points on average
  if (ix->m > 0) {
    start = (int)std::floor(0.5 * std::log2((double)ix->m * 4.0 / (6.0 * (16.0 + 4.0 * max_results))));
    start = start < -1 ? -1 : (start > 29 ? 29 : start);
  }
  if (max_results == 0) start = 29;
  const double limit_angle = max_length2 >= 4.0 ? 4.0 : angle_of_length2(max_length2 > 0 ? max_length2 : 0.0);   // 4 > pi: "no limit"
  const IndexView v{ix->ids, ix->pts, ix->data, ix->first, (uint32_t)ix->m, 61 - 2 * ix->table_level};
  // Per-point minima over many targets (min_bits): a target's candidates are independent, so a group of neighbouring lanes
  // shares one target and strides its candidate ranges (coalesced, and parallelism no longer ends at the number of
  // targets). Group size from the expected number of candidates per target: m * (limit angle)^2 / 4 on a uniform sphere.
  int lane_shift = 0;
  if (min_bits) {
    const double la = limit_angle > 3.2 ? 3.2 : limit_angle;
    const double expected = (double)ix->m * la * la / 4.0;
    lane_shift = expected >= 256 ? 5 : (expected >= 32 ? 3 : 0);
  }
  const int grid = grid_for_pi(n << lane_shift, (const void*)closest_targets_kernel<kTargetEdge>);
  if (target_kind == kTargetPoint)
    closest_targets_kernel<kTargetPoint><<<grid, kThreads, 0, st>>>(v, targets_dev, n, max_results, start, max_length2, limit_angle, cov_dev, data_out_dev, length2_out_dev, count_out_dev, min_bits, lane_shift);
  else if (target_kind == kTargetEdge)
    closest_targets_kernel<kTargetEdge><<<grid, kThreads, 0, st>>>(v, targets_dev, n, max_results, start, max_length2, limit_angle, cov_dev, data_out_dev, length2_out_dev, count_out_dev, min_bits, lane_shift);
  else
    closest_targets_kernel<kTargetCell><<<grid, kThreads, 0, st>>>(v, targets_dev, n, max_results, start, max_length2, limit_angle, cov_dev, data_out_dev, length2_out_dev, count_out_dev, min_bits, lane_shift);
  count_launch();
  e = cudaGetLastError();
  cudaFreeAsync(cov_dev, st);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "closest_targets launch");
}

int s2b_closest_points_to_targets(const S2bPointIndex* ix, int target_kind, const void* targets, size_t n, int max_results, double max_length2,
                                  double max_error_radians, int32_t* data_out, double* length2_out, uint32_t* count_out) {
  if (target_kind < 0 || target_kind > 2) return set_error(S2B_EINVAL, "target_kind must be S2B_TARGET_POINT, _EDGE or _CELL");
  if (max_results < 0 || max_results > kMaxResults) return set_error(S2B_EINVAL, "max_results must be in [0, %d] (0 = count only)", kMaxResults);
  if (n == 0) return S2B_OK;
  if (!targets) return set_error(S2B_EINVAL, "null pointer");
  const size_t k = (size_t)max_results, tbytes = target_kind == kTargetPoint ? 24 : (target_kind == kTargetEdge ? 48 : 8);
  void *t = nullptr, *d = nullptr, *l = nullptr, *c = nullptr;
  cudaError_t e = cudaMalloc(&t, tbytes * n);
  if (e == cudaSuccess) e = cudaMalloc(&d, 4 * n * (k ? k : 1));
  if (e == cudaSuccess) e = cudaMalloc(&l, 8 * n * (k ? k : 1));
  if (e == cudaSuccess) e = cudaMalloc(&c, 4 * n);
  if (e == cudaSuccess) e = cudaMemcpy(t, targets, tbytes * n, cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? s2b_closest_points_to_targets_dev(ix, target_kind, t, n, max_results, max_length2, max_error_radians, (int32_t*)d, (double*)l, (uint32_t*)c, nullptr)
                            : cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK && k && data_out && (e = cudaMemcpy(data_out, d, 4 * n * k, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_OK && k && length2_out && (e = cudaMemcpy(length2_out, l, 8 * n * k, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_OK && count_out && (e = cudaMemcpy(count_out, c, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  cudaFree(t); cudaFree(d); cudaFree(l); cudaFree(c);
  return rc;
}

// Host-pointer forms (whole-array copies).
int s2b_closest_points(const S2bPointIndex* ix, const double* targets, size_t n, int max_results, double max_length2, int32_t* data_out,
                       double* length2_out, uint32_t* count_out) {
  if (max_results < 1 || max_results > kMaxResults) return set_error(S2B_EINVAL, "max_results must be in [1, %d]", kMaxResults);
  if (n == 0) return S2B_OK;
  if (!targets || !data_out) return set_error(S2B_EINVAL, "null pointer");
  const size_t k = (size_t)max_results;
  void *t = nullptr, *d = nullptr, *l = nullptr, *c = nullptr;
  cudaError_t e = cudaMalloc(&t, 24 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d, 4 * n * k);
  if (e == cudaSuccess) e = cudaMalloc(&l, 8 * n * k);
  if (e == cudaSuccess) e = cudaMalloc(&c, 4 * n);
  if (e == cudaSuccess) e = cudaMemcpy(t, targets, 24 * n, cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? s2b_closest_points_dev(ix, (const double*)t, n, max_results, max_length2, (int32_t*)d, (double*)l, (uint32_t*)c, nullptr)
                            : cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK && (e = cudaMemcpy(data_out, d, 4 * n * k, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_OK && length2_out && (e = cudaMemcpy(length2_out, l, 8 * n * k, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_OK && count_out && (e = cudaMemcpy(count_out, c, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  cudaFree(t); cudaFree(d); cudaFree(l); cudaFree(c);
  return rc;
}

int s2b_closest_points_count(const S2bPointIndex* ix, const double* targets, size_t n, double max_length2, uint32_t* counts_out) {
  if (n == 0) return S2B_OK;
  if (!targets || !counts_out) return set_error(S2B_EINVAL, "null pointer");
  void *t = nullptr, *c = nullptr;
  cudaError_t e = cudaMalloc(&t, 24 * n);
  if (e == cudaSuccess) e = cudaMalloc(&c, 4 * n);
  if (e == cudaSuccess) e = cudaMemcpy(t, targets, 24 * n, cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? s2b_closest_points_count_dev(ix, (const double*)t, n, max_length2, (uint32_t*)c, nullptr) : cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK && (e = cudaMemcpy(counts_out, c, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  cudaFree(t); cudaFree(c);
  return rc;
}

int s2b_closest_point(const S2bPointIndex* ix, const double* targets, size_t n, double max_length2, int32_t* data_out, double* length2_out) {
  if (n == 0) return S2B_OK;
  if (!targets || !data_out) return set_error(S2B_EINVAL, "null pointer");
  void *t = nullptr, *d = nullptr, *l = nullptr;
  cudaError_t e = cudaMalloc(&t, 24 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d, 4 * n);
  if (e == cudaSuccess) e = cudaMalloc(&l, 8 * n);
  if (e == cudaSuccess) e = cudaMemcpy(t, targets, 24 * n, cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? s2b_closest_point_dev(ix, (const double*)t, n, max_length2, (int32_t*)d, (double*)l, nullptr) : cuda_fail(e, "cudaMalloc/cudaMemcpy");
  if (rc == S2B_OK && (e = cudaMemcpy(data_out, d, 4 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  if (rc == S2B_OK && length2_out && (e = cudaMemcpy(length2_out, l, 8 * n, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy");
  cudaFree(t); cudaFree(d); cudaFree(l);
  return rc;
}


// FindClosestPoints(ShapeIndexTarget(index)) (s2closest_point_query.h:132-165, S2MinDistanceShapeIndexTarget in
// s2min_distance_targets.cc:196-255): the distance of an index point p to the target is what
// S2ClosestEdgeQuery(index).FindClosestEdge(PointTarget(p)) returns — 0 if include_interiors and a polygon of the index
// contains p (S2ContainsPointQuery, SEMI_OPEN), else the minimum over the index's edges of S2::UpdateMinDistance(p, v0, v1).
// Here the loops are turned inside out: every (clipped) edge of the device-resident index is an edge target whose
// candidates within the limit are scanned as for EdgeTarget, and each candidate's slot keeps the minimum (atomicMin on
// the bits of a non-negative double); points inside polygons are found by the index's own Contains query. The slots are
// then sorted by (distance, data). With max_results = k and no distance limit, the k index points nearest to one vertex
// of the target bound the k-th result's distance from above, which gives the edge scan a finite radius.
int s2b_closest_points_to_shape_index(const S2bPointIndex* pix, const S2bIndex* target, int include_interiors, int max_results, double max_length2,
                                      double max_error_radians, int32_t* data_out, double* length2_out, size_t capacity, size_t* total_out) {
  int rc = check_index_device(pix);
  if (rc != S2B_OK) return rc;
  if (!target || !total_out) return set_error(S2B_EINVAL, "null argument");
  if (target->device != pix->device) return set_error(S2B_EINVAL, "the point index (device %d) and the target index (device %d) must be on the same device", pix->device, target->device);
  if (max_results < 0 || std::isnan(max_length2) || !(max_error_radians >= 0)) return set_error(S2B_EINVAL, "bad max_results / max_length2 / max_error");
  if (capacity && (!data_out)) return set_error(S2B_EINVAL, "null output with non-zero capacity");
  *total_out = 0;
  const size_t m = pix->m;
  const size_t E = (size_t)target->num_edges;
  if (m == 0 || (E == 0 && target->view.num_cells == 0) || !(max_length2 > 0)) return S2B_OK;
  cudaStream_t st = nullptr;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t t_sort = 0, t_sort2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (const int32_t*)nullptr, (int32_t*)nullptr, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)m, 0, 32, st);
  cub::DeviceRadixSort::SortPairs(nullptr, t_sort2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)m, 0, 64, st);
  if (t_sort2 > t_sort) t_sort = t_sort2;
  const size_t o_bits = 0, o_bits2 = up(8 * m), o_data = o_bits2 + up(8 * m), o_data2 = o_data + up(4 * m), o_flags = o_data2 + up(4 * m),
               o_total = o_flags + up(m), o_tmp = o_total + 256;
  char* blob = nullptr;
  cudaError_t e = cudaMalloc((void**)&blob, o_tmp + up(t_sort));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  unsigned long long* bits = (unsigned long long*)(blob + o_bits);
  unsigned long long* bits2 = (unsigned long long*)(blob + o_bits2);
  int32_t* data = (int32_t*)(blob + o_data);
  int32_t* data2 = (int32_t*)(blob + o_data2);
  uint8_t* flags = (uint8_t*)(blob + o_flags);
  unsigned long long* total_dev = (unsigned long long*)(blob + o_total);
  const int blocks = (int)((m + 255) / 256);
  auto done = [&](int code) { cudaFree(blob); return code; };
  fill_u64_kernel<<<blocks, 256, 0, st>>>(bits, m, kInfBits);
  count_launch();
  // a finite radius for "the k nearest" without a distance limit (or a tighter one than the caller's)
  double limit2 = max_length2;
  if (max_results > 0 && max_results <= kMaxResults && m > (size_t)max_results && E > 0) {
    double* kd = (double*)bits2;                       // scratch: k distances
    int32_t* ki = data2;
    rc = s2b_closest_points_dev(pix, (const double*)target->view.edges, 1, max_results, HUGE_VAL, ki, kd, nullptr, st);
    if (rc != S2B_OK) return done(rc);
    double far = 0;
    e = cudaMemcpy(&far, kd + (max_results - 1), 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
    const double bound = far * (1 + 1e-12) + 1e-300;      // the limit is exclusive; at least k points lie within `far` of the target
    if (bound < limit2) limit2 = bound;
  }
  if (include_interiors && target->view.num_cells > 0) {
    rc = s2b_contains_dev(target, pix->pts, m, S2B_VERTEX_MODEL_SEMI_OPEN, flags, st);
    if (rc != S2B_OK) return done(rc);
    zero_where_flag_kernel<<<blocks, 256, 0, st>>>(bits, flags, m);
    count_launch();
  }
  if (E > 0) {
    rc = launch_closest_targets(pix, kTargetEdge, target->view.edges, E, 0, limit2, nullptr, nullptr, nullptr, bits, st);
    if (rc != S2B_OK) return done(rc);
  }
  // order by (distance, data): stable sort by data, then by the distance bits
  e = cub::DeviceRadixSort::SortPairs(blob + o_tmp, t_sort, pix->data, data2, (const unsigned long long*)bits, bits2, (int)m, 0, 32, st);
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(blob + o_tmp, t_sort, (const unsigned long long*)bits2, bits, (const int32_t*)data2, data, (int)m, 0, 64, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(total_dev, 0, 8, st);
  if (e != cudaSuccess) return done(cuda_fail(e, "result sort"));
  count_launch(); count_launch();
  unsigned long long limit_bits;
  { const double l = limit2 < HUGE_VAL ? limit2 : HUGE_VAL; memcpy(&limit_bits, &l, 8); }
  if (!(max_length2 < HUGE_VAL) && limit2 == max_length2) limit_bits = kInfBits;   // no limit: every point reached by a distance qualifies
  count_below_sorted_kernel<<<blocks, 256, 0, st>>>(bits, m, limit_bits, total_dev);
  count_launch();
  unsigned long long total = 0;
  e = cudaMemcpy(&total, total_dev, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
  if (max_results > 0 && total > (unsigned long long)max_results) total = (unsigned long long)max_results;
  *total_out = (size_t)total;
  const size_t w = total < capacity ? (size_t)total : capacity;
  if (w) {
    e = cudaMemcpy(data_out, data, 4 * w, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && length2_out) e = cudaMemcpy(length2_out, bits, 8 * w, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(cuda_fail(e, "cudaMemcpy"));
  }
  return done(total > capacity ? set_error(S2B_ECAPACITY, "%llu results, capacity %zu", total, capacity) : S2B_OK);
}


// S2ClosestEdgeQuery(index).GetDistance(PointTarget(points[i])) for a batch (s2closest_edge_query.h:301-318; options
// include_interiors and max_distance): the ShapeIndexTarget machinery above run over a temporary point index of the batch
// (cell-sorted points; every clipped edge of the index scans its candidates within the limit and keeps the per-point
// minimum), then scattered back to the caller's order. xyz / length2_out: both host or both device pointers.
static int point_distances(const S2bIndex* target, const double* xyz, size_t n, int include_interiors, double max_length2, double* out, bool device_out) {
  if (!target) return set_error(S2B_EINVAL, "null index");
  if (std::isnan(max_length2)) return set_error(S2B_EINVAL, "max_length2 is NaN");
  if (n == 0) return S2B_OK;
  if (!xyz || !out) return set_error(S2B_EINVAL, "null pointer");
  S2bPointIndex* pix = nullptr;
  int rc = s2b_point_index_create(xyz, n, &pix);
  if (rc != S2B_OK) return rc;
  if (pix->device != target->device) { s2b_point_index_destroy(pix); return set_error(S2B_EINVAL, "the current device (%d) is not the index's device (%d)", pix->device, target->device); }
  unsigned long long* bits = nullptr;
  double* out_dev = device_out ? out : nullptr;
  uint8_t* flags = nullptr;
  cudaError_t e = cudaMalloc((void**)&bits, 8 * n);
  if (e == cudaSuccess && !device_out) e = cudaMalloc((void**)&out_dev, 8 * n);
  if (e == cudaSuccess && include_interiors) e = cudaMalloc((void**)&flags, n);
  auto done = [&](int code) { cudaFree(bits); if (!device_out) cudaFree(out_dev); cudaFree(flags); s2b_point_index_destroy(pix); return code; };
  if (e != cudaSuccess) return done(cuda_fail(e, "cudaMalloc"));
  const int blocks = (int)((n + 255) / 256);
  cudaStream_t st = nullptr;
  fill_u64_kernel<<<blocks, 256, 0, st>>>(bits, n, kInfBits);
  count_launch();
  if (max_length2 > 0) {
    if (include_interiors && target->view.num_cells > 0) {
      rc = s2b_contains_dev(target, pix->pts, n, S2B_VERTEX_MODEL_SEMI_OPEN, flags, st);
      if (rc != S2B_OK) return done(rc);
      zero_where_flag_kernel<<<blocks, 256, 0, st>>>(bits, flags, n);
      count_launch();
    }
    if (target->num_edges > 0) {
      rc = launch_closest_targets(pix, kTargetEdge, target->view.edges, (size_t)target->num_edges, 0, max_length2, nullptr, nullptr, nullptr, bits, st);
      if (rc != S2B_OK) return done(rc);
    }
  }
  scatter_distances_kernel<<<blocks, 256, 0, st>>>(bits, pix->data, n, kInfBits, out_dev);
  count_launch();
  e = cudaGetLastError();
  if (e == cudaSuccess) e = device_out ? cudaStreamSynchronize(st) : cudaMemcpy(out, out_dev, 8 * n, cudaMemcpyDeviceToHost);
  return done(e == cudaSuccess ? S2B_OK : cuda_fail(e, "point distances"));
}

int s2b_index_point_distances(const S2bIndex* index, const double* xyz, size_t n, int include_interiors, double max_length2, double* length2_out) {
  return point_distances(index, xyz, n, include_interiors, max_length2, length2_out, false);
}
int s2b_index_point_distances_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int include_interiors, double max_length2, double* length2_out_dev) {
  return point_distances(index, xyz_dev, n, include_interiors, max_length2, length2_out_dev, true);
}

}  // extern "C"
