// s2b_region.cuh — S2ShapeIndexRegion's cell tests over the flattened index (host-compilable for tests/hostsim).
// Restates, per target cell:
//   S2ShapeIndexRegion::Contains(S2Cell)      s2shape_index_region.h:313-342
//   S2ShapeIndexRegion::MayIntersect(S2Cell)  s2shape_index_region.h:344-371
//   S2ShapeIndexRegion::AnyEdgeIntersects     s2shape_index_region.h:447-465
//   S2::ClipToPaddedFace / IntersectsRect     s2edge_clipping.cc:64-144,271-384
//   S2::RobustCrossProd (all stages)          s2edge_crossings.cc:97-359
//   S2Cell(S2CellId) bound / centre           s2cell.cc:64-71, s2cell_id.cc:395-405, s2cell.h:217
//   S2CellIterator::LocateImpl(S2CellId)      s2cell_iterator.h:159-189
// These are the calls an S2RegionCoverer makes on an index region, batched: one target cell per thread.
//
// RobustCrossProd: the double-precision stage decides unless the endpoints are closer than ~9 * DBL_EPSILON (or
// antipodal); a == b returns Ortho(a). The reference's two further stages run here too, on the device: the x87 long
// double product (emulated operation by operation, s2b_ext80.cuh) and exact arithmetic (the superaccumulator of
// s2b_predicates.cuh read back as ExactFloat::ToDouble would) with symbolic perturbation — no index is refused.
#pragma once
#include <float.h>

#include "s2b_contains.cuh"
#include "s2b_ext80.cuh"

namespace s2b {

struct Uv { double u, v; };
struct UvRect { double ulo, uhi, vlo, vhi; };

// kFaceClipErrorUVCoord + kIntersectsRectErrorUVDist (s2edge_clipping.h:101,113; summed at s2shape_index_region.h:450).
S2B_HD double region_max_error() { return 9 * 0.70710678118654752440 * DBL_EPSILON + 3 * 1.41421356237309504880 * DBL_EPSILON; }

S2B_HD P3 add(const P3& a, const P3& b) { return P3{a.x + b.x, a.y + b.y, a.z + b.z}; }

// S2::FaceXYZtoUVW (s2coords.cc:27-40).
S2B_HD P3 face_xyz_to_uvw(int face, const P3& p) {
  switch (face) {
    case 0: return P3{p.y, p.z, p.x};
    case 1: return P3{-p.x, p.z, p.y};
    case 2: return P3{-p.x, -p.y, p.z};
    case 3: return P3{-p.z, -p.y, -p.x};
    case 4: return P3{-p.z, p.x, -p.y};
    default: return P3{p.y, p.x, -p.z};
  }
}

// S2::GetFace (s2coords.h:409-413).
S2B_HD int get_face(const P3& p) {
  const int k = largest_abs_component(p);
  return comp(p, k) < 0 ? k + 3 : k;
}

// S2::ValidFaceXYZtoUV (s2coords.h:389-403) = (u, v) of FaceXYZtoUVW divided by w.
S2B_HD Uv valid_face_xyz_to_uv(int face, const P3& p) {
  const P3 q = face_xyz_to_uvw(face, p);
  // The reference divides by the signed major axis: faces 0-2 by +p[axis] (= w), faces 3-5 by p[axis] = -w with
  // the numerators' signs folded in. x / w == (-x) / (-w) exactly in IEEE arithmetic, so one form serves all six.
  return Uv{q.x / q.z, q.y / q.z};
}

// GetStableCrossProd<double> (s2edge_crossings.cc:97-136): (a - b) x (a + b), accepted when long enough.
S2B_HD bool stable_cross_prod(const P3& a, const P3& b, P3* out) {
  const double kSqrt3 = 1.7320508075688772935;
  const double kErr = DBL_EPSILON / 2;   // rounding_epsilon<double>() = DBL_ERR
  const double kMinNorm = (32 * kSqrt3 * kErr) / ((6 * kErr) / kErr - (1 + 2 * kSqrt3));
  *out = cross(sub(a, b), add(a, b));
  return norm2(*out) >= kMinNorm * kMinNorm;
}

// GetStableCrossProd<long double> (s2edge_crossings.cc:97-136) as the x87 evaluates it (s2b_ext80.cuh): every
// subtraction, addition and product of (a - b) x (a + b) and of its squared norm rounded to a 64-bit significand, the
// test against kMinNorm^2 for T = long double (the constant below is that expression evaluated by the same compiler:
// 0xAACA6DBF0CAB2C09 * 2^-185; tests/test_region_cpu.py recomputes it), and Vector3_d::Cast of the result.
S2B_HD_NOINLINE bool stable_cross_prod_ld(const P3& a, const P3& b, P3* out) {
  const X80 ax = x80_from_double(a.x), ay = x80_from_double(a.y), az = x80_from_double(a.z);
  const X80 bx = x80_from_double(b.x), by = x80_from_double(b.y), bz = x80_from_double(b.z);
  const X80 dx = x80_add(ax, bx, true), dy = x80_add(ay, by, true), dz = x80_add(az, bz, true);   // a - b
  const X80 sx = x80_add(ax, bx), sy = x80_add(ay, by), sz = x80_add(az, bz);                     // a + b
  // Vector3::CrossProd (util/math/vector.h:474-478): (y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x)
  const X80 nx = x80_add(x80_mul(dy, sz), x80_mul(dz, sy), true);
  const X80 ny = x80_add(x80_mul(dz, sx), x80_mul(dx, sz), true);
  const X80 nz = x80_add(x80_mul(dx, sy), x80_mul(dy, sx), true);
  // Norm2 = ((0 + x*x) + y*y) + z*z
  const X80 n2 = x80_add(x80_add(x80_mul(nx, nx), x80_mul(ny, ny)), x80_mul(nz, nz));
  *out = P3{x80_to_double(nx), x80_to_double(ny), x80_to_double(nz)};
  const X80 kMinNorm2{0xAACA6DBF0CAB2C09ull, -185 + 63, false};
  return x80_ge(n2, kMinNorm2);
}

// Exact (a x b)[k] = p*q - r*s read back the way ExactFloat reports it.
// A zero result carries the sign IEEE 754 (and ExactFloat::SignedSum) gives it: (-0) - (+0) = -0, every other
// combination of zero products and every exact cancellation +0; it is reported in `top` (1 = negative zero).
S2B_HD_NOINLINE ExactAcc::Read exact_minor(double p, double q, double r, double s) {
  ExactAcc acc;
  acc.clear();
  acc.add_product(p, q, 1.0, 1);
  acc.add_product(r, s, 1.0, -1);
  ExactAcc::Read v = acc.read();
  if (v.sign == 0) {
    const bool pq_zero = p == 0 || q == 0, rs_zero = r == 0 || s == 0;
    const bool pq_neg = signbit(p) != signbit(q), rs_neg = signbit(r) != signbit(s);
    v.top = (pq_zero && rs_zero && pq_neg && !rs_neg) ? 1 : 0;
  }
  return v;
}
S2B_HD double exact_read_to_double(const ExactAcc::Read& v, int scale_exp) {   // ToDouble(ldexp(value, -scale_exp))
  if (v.sign == 0) return v.top ? -0.0 : 0.0;
  const double d = ldexp((double)v.mant, v.exp2 - scale_exp);
  return v.sign < 0 ? -d : d;
}

// IsNormalizable / EnsureNormalizable (s2edge_crossings.cc:268-303).
S2B_HD bool is_normalizable(const P3& p) { return fmax(fabs(p.x), fmax(fabs(p.y), fabs(p.z))) >= 0x1p-242; }
S2B_HD P3 ensure_normalizable(const P3& p) {
  if (is_normalizable(p)) return p;
  const double p_max = fmax(fabs(p.x), fmax(fabs(p.y), fabs(p.z)));
  const double f = ldexp(2.0, -1 - ilogb(p_max));
  return P3{f * p.x, f * p.y, f * p.z};
}

// SymbolicCrossProdSorted (s2edge_crossings.cc:183-265), requires a < b lexicographically and a x b == 0 exactly.
S2B_HD P3 symbolic_cross_prod_sorted(const P3& a, const P3& b) {
  if (b.x != 0 || b.y != 0) return P3{-b.y, b.x, 0};
  if (b.z != 0) return P3{b.z, 0, 0};
  if (a.x != 0 || a.y != 0) return P3{a.y, -a.x, 0};
  return P3{1, 0, 0};
}

// internal::ExactCrossProd (s2edge_crossings.cc:345-358) incl. NormalizableFromExact (:305-330). Requires a != b.
S2B_HD_NOINLINE P3 exact_cross_prod(const P3& a, const P3& b) {
  const ExactAcc::Read x = exact_minor(a.y, b.z, a.z, b.y), y = exact_minor(a.z, b.x, a.x, b.z), z = exact_minor(a.x, b.y, a.y, b.x);
  if (x.sign != 0 || y.sign != 0 || z.sign != 0) {
    const P3 d{exact_read_to_double(x, 0), exact_read_to_double(y, 0), exact_read_to_double(z, 0)};
    if (is_normalizable(d)) return d;
    // scale so that the largest component magnitude is in [0.5, 1): exp = the largest ExactFloat::exp() of the components
    int e = -(1 << 30);
    if (x.sign != 0 && x.top > e) e = x.top;
    if (y.sign != 0 && y.top > e) e = y.top;
    if (z.sign != 0 && z.top > e) e = z.top;
    return P3{exact_read_to_double(x, e), exact_read_to_double(y, e), exact_read_to_double(z, e)};
  }
  // exactly proportional: symbolic perturbation (SymbolicCrossProd :334-343 applied as ExactCrossProd :352-357 does)
  if (lex_less(a, b)) return ensure_normalizable(ensure_normalizable(symbolic_cross_prod_sorted(a, b)));
  return neg(ensure_normalizable(ensure_normalizable(symbolic_cross_prod_sorted(b, a))));
}

// Statistic: true when RobustCrossProd(a, b) is decided by its double-precision stage (or a == b), i.e. for all but
// edges shorter than ~9 * DBL_EPSILON or nearly antipodal.
S2B_HD bool region_edge_is_plain(const P3& a, const P3& b) {
  P3 n;
  return stable_cross_prod(a, b, &n) || eq(a, b);
}

// S2::RobustCrossProd (s2edge_crossings.cc:147-177), every stage: stable double product; a == b -> Ortho(a); the x87
// long double product (the reference on x86-64: s2pred::kHasLongDouble); exact arithmetic; symbolic perturbation.
S2B_HD P3 robust_cross_prod(const P3& a, const P3& b) {
  P3 n;
  if (stable_cross_prod(a, b, &n)) return n;
  if (eq(a, b)) return ortho(a);
  if (stable_cross_prod_ld(a, b, &n)) return n;
  return exact_cross_prod(a, b);
}

// IntersectsFace / IntersectsOppositeEdges / GetExitAxis / GetExitPoint (s2edge_clipping.cc:71-136); n is the line's
// normal in the face's (u, v, w) frame.
S2B_HD bool clip_intersects_face(const P3& n) {
  const double u = fabs(n.x), v = fabs(n.y), w = fabs(n.z);
  return (v >= w - u) && (u >= w - v);
}
S2B_HD bool clip_intersects_opposite_edges(const P3& n) {
  const double u = fabs(n.x), v = fabs(n.y), w = fabs(n.z);
  if (fabs(u - v) != w) return fabs(u - v) >= w;
  return (u >= v) ? (u - w >= v) : (v - w >= u);
}
S2B_HD int clip_exit_axis(const P3& n) {
  if (clip_intersects_opposite_edges(n)) return (fabs(n.x) >= fabs(n.y)) ? 1 : 0;
  const int negatives = (signbit(n.x) ? 1 : 0) ^ (signbit(n.y) ? 1 : 0) ^ (signbit(n.z) ? 1 : 0);
  return negatives == 0 ? 1 : 0;
}
S2B_HD Uv clip_exit_point(const P3& n, int axis) {
  if (axis == 0) {
    const double u = (n.y > 0) ? 1.0 : -1.0;
    return Uv{u, (-u * n.x - n.z) / n.y};
  }
  const double v = (n.x < 0) ? 1.0 : -1.0;
  return Uv{(-v * n.y - n.z) / n.x, v};
}

// ClipDestination (s2edge_clipping.cc:271-321): clips endpoint B of AB to the padded face; returns the score.
S2B_HD int clip_destination(const P3& a, const P3& b, const P3& scaled_n, const P3& a_tangent, const P3& b_tangent,
                            double scale_uv, Uv* uv) {
  const double kMaxSafeUVCoord = 1 - 9 * 0.70710678118654752440 * DBL_EPSILON;
  if (b.z > 0) {
    *uv = Uv{b.x / b.z, b.y / b.z};
    const double mu = fabs(uv->u), mv = fabs(uv->v);
    if ((mu > mv ? mu : mv) <= kMaxSafeUVCoord) return 0;
  }
  const Uv e = clip_exit_point(scaled_n, clip_exit_axis(scaled_n));
  *uv = Uv{scale_uv * e.u, scale_uv * e.v};
  const P3 p{uv->u, uv->v, 1.0};
  int score = 0;
  if (dot(sub(p, a), a_tangent) < 0) score = 2;
  else if (dot(sub(p, b), b_tangent) < 0) score = 1;
  if (score > 0) {
    if (b.z <= 0) score = 3;
    else *uv = Uv{b.x / b.z, b.y / b.z};
  }
  return score;
}

// S2::ClipToPaddedFace (s2edge_clipping.cc:323-362).
S2B_HD bool clip_to_padded_face(const P3& a_xyz, const P3& b_xyz, int face, double padding, Uv* a_uv, Uv* b_uv) {
  if (get_face(a_xyz) == face && get_face(b_xyz) == face) {
    *a_uv = valid_face_xyz_to_uv(face, a_xyz);
    *b_uv = valid_face_xyz_to_uv(face, b_xyz);
    return true;
  }
  P3 n = face_xyz_to_uvw(face, robust_cross_prod(a_xyz, b_xyz));
  const P3 a = face_xyz_to_uvw(face, a_xyz);
  const P3 b = face_xyz_to_uvw(face, b_xyz);
  const double scale_uv = 1 + padding;
  const P3 scaled_n{scale_uv * n.x, scale_uv * n.y, n.z};
  if (!clip_intersects_face(scaled_n)) return false;
  n = normalize(n);
  const P3 a_tangent = cross(n, a);
  const P3 b_tangent = cross(b, n);
  const int a_score = clip_destination(b, a, neg(scaled_n), b_tangent, a_tangent, scale_uv, a_uv);
  const int b_score = clip_destination(a, b, scaled_n, a_tangent, b_tangent, scale_uv, b_uv);
  return a_score + b_score < 3;
}

// R1Interval::Intersects (r1interval.h:128-134).
S2B_HD bool interval_intersects(double lo, double hi, double ylo, double yhi) {
  if (lo <= ylo) return ylo <= hi && ylo <= yhi;
  return lo <= yhi && lo <= hi;
}

// S2::IntersectsRect (s2edge_clipping.cc:364-384).
S2B_HD bool intersects_rect(const Uv& a, const Uv& b, const UvRect& r) {
  const double xlo = a.u <= b.u ? a.u : b.u, xhi = a.u <= b.u ? b.u : a.u;
  const double ylo = a.v <= b.v ? a.v : b.v, yhi = a.v <= b.v ? b.v : a.v;
  if (!(interval_intersects(r.ulo, r.uhi, xlo, xhi) && interval_intersects(r.vlo, r.vhi, ylo, yhi))) return false;
  const double nx = -(b.v - a.v), ny = b.u - a.u;   // (b - a).Ortho()
  const bool i = nx >= 0, j = ny >= 0;
  const double mx = (i ? r.uhi : r.ulo) - a.u, my = (j ? r.vhi : r.vlo) - a.v;
  const double lx = (i ? r.ulo : r.uhi) - a.u, ly = (j ? r.vlo : r.vhi) - a.v;
  const double mx_dot = (0.0 + nx * mx) + ny * my;
  const double mn_dot = (0.0 + nx * lx) + ny * ly;
  return (mx_dot >= 0) && (mn_dot <= 0);
}

// What S2Cell(S2CellId) precomputes and the region tests read: face, (u, v) bound, centre.
struct TargetCell { uint64_t id; int face; UvRect bound; };

template <class TableT>
S2B_HD TargetCell target_cell(uint64_t id, const TableT* ij_table) {
  TargetCell t;
  t.id = id;
  int i, j;
  t.face = cellid_to_face_ij(id, ij_table, &i, &j);
  const int size = 1 << (30 - cellid_level(id));                 // GetSizeIJ(level)
  const int ilo = i & -size, jlo = j & -size;
  const double k = 1.0 / 1073741824.0;                           // IJtoSTMin (s2coords.h:340-343)
  t.bound = UvRect{st_to_uv(k * ilo), st_to_uv(k * (ilo + size)), st_to_uv(k * jlo), st_to_uv(k * (jlo + size))};
  return t;
}

// S2ShapeIndexRegion::AnyEdgeIntersects over one clipped shape's edges.
S2B_HD bool any_edge_intersects(const FlatIndexView& ix, const FlatClip& clip, const TargetCell& t) {
  const double max_error = region_max_error();
  const UvRect padded{t.bound.ulo - max_error, t.bound.uhi + max_error, t.bound.vlo - max_error, t.bound.vhi + max_error};
  const FlatEdge* edges = ix.edges + clip.edge_begin;
  for (uint32_t e = 0; e < clip.num_edges; ++e) {
    const FlatEdge ed = edges[e];
    Uv p0, p1;
    if (clip_to_padded_face(edge_v0(ed), edge_v1(ed), t.face, max_error, &p0, &p1) && intersects_rect(p0, p1, padded)) return true;
  }
  return false;
}

// Iterator::Locate(S2CellId) (s2cell_iterator.h:159-189). Returns 0 INDEXED (target inside index cell *at),
// 1 SUBDIVIDED (*at = first index cell inside the target), 2 DISJOINT.
S2B_HD int locate_target(const FlatIndexView& ix, uint64_t target, int32_t* at) {
  const uint32_t C = ix.num_cells;
  const uint64_t tmin = cellid_range_min(target), tmax = cellid_range_max(target);
  uint32_t lo = 0, hi = C;
  if (ix.id_first) {   // lower_bound(tmin) lies inside the bucket of tmin: cells before it are smaller, cells after it larger
    uint64_t b = tmin >> ix.id_shift;
    if (b >= ix.id_buckets) b = ix.id_buckets - 1;             // invalid ids beyond face 5
    lo = ix.id_first[b]; hi = ix.id_first[b + 1];
  }
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ix.cell_ids[mid] < tmin) lo = mid + 1; else hi = mid; }
  if (lo < C) {
    const uint64_t id = ix.cell_ids[lo];
    if (id >= target && cellid_range_min(id) <= target) { *at = (int32_t)lo; return 0; }
    if (id <= tmax) { *at = (int32_t)lo; return 1; }
  }
  if (lo > 0 && cellid_range_max(ix.cell_ids[lo - 1]) >= target) { *at = (int32_t)lo - 1; return 0; }
  *at = -1;
  return 2;
}

// ShapeContains(iter.id(), clipped, target.GetCenter()) with the SEMI_OPEN model of S2ShapeIndexRegion's query.
S2B_HD bool clip_contains_centre(const FlatIndexView& ix, const FlatClip& clip, const P3& cell_centre, const P3& centre,
                                 const P3& cross_ab) {
  return clip_contains(ix, clip, cell_centre, centre, cross_ab, kSemiOpen);
}

template <class TableT>
S2B_HD bool region_may_intersect(const FlatIndexView& ix, uint64_t target, const TableT* ij_table) {
  int32_t at;
  const int rel = locate_target(ix, target, &at);
  if (rel == 2) return false;
  if (rel == 1) return true;
  if (ix.cell_ids[at] == target) return true;
  const FlatCell fc = ix.cells[at];
  const TargetCell t = target_cell(target, ij_table);
  const P3 a{fc.cx, fc.cy, fc.cz};
  const P3 centre = cellid_to_point(target, ij_table);
  const P3 a_cross_b = cross(a, centre);
  for (uint32_t k = 0; k < fc.clip_count; ++k) {
    const FlatClip clip = ix.clips[fc.clip_begin + k];
    if (any_edge_intersects(ix, clip, t)) return true;
    if (clip_contains_centre(ix, clip, a, centre, a_cross_b)) return true;
  }
  return false;
}

template <class TableT>
S2B_HD bool region_contains_cell(const FlatIndexView& ix, uint64_t target, const TableT* ij_table) {
  int32_t at;
  if (locate_target(ix, target, &at) != 0) return false;
  const FlatCell fc = ix.cells[at];
  const bool same = ix.cell_ids[at] == target;
  TargetCell t{};
  P3 a{fc.cx, fc.cy, fc.cz}, centre{0, 0, 0}, a_cross_b{0, 0, 0};
  if (!same) {
    t = target_cell(target, ij_table);
    centre = cellid_to_point(target, ij_table);
    a_cross_b = cross(a, centre);
  }
  for (uint32_t k = 0; k < fc.clip_count; ++k) {
    const FlatClip clip = ix.clips[fc.clip_begin + k];
    if (same) {
      if (clip.num_edges == 0 && (clip.flags & 1u)) return true;
    } else if (((clip.flags >> 1) & 3u) == 2 && !any_edge_intersects(ix, clip, t) &&
               clip_contains_centre(ix, clip, a, centre, a_cross_b)) {
      return true;
    }
  }
  return false;
}

// One entry of VisitIntersectingShapeIds (s2shape_index_region.h:373-424) for the INDEXED case: 0 = the clipped shape
// is disjoint from the target, 1 = intersects, 2 = contains the target.
S2B_HD int region_clip_relation(const FlatIndexView& ix, const FlatClip& clip, bool same, const TargetCell& t, const P3& a,
                                const P3& centre, const P3& a_cross_b) {
  if (same) return (clip.num_edges == 0 && (clip.flags & 1u)) ? 2 : 1;
  if (any_edge_intersects(ix, clip, t)) return 1;
  return clip_contains_centre(ix, clip, a, centre, a_cross_b) ? 2 : 0;
}

// (S2CellId::GetCommonAncestorLevel is cellid_common_ancestor_level in s2b_core.cuh.)

// S2ShapeIndexRegion::GetCellUnionBound (s2shape_index_region.h:232-305) over the sorted index cell ids: at most six
// cells (one per face the index touches, or the up-to-four children of the smallest cell holding a single-face index,
// each shrunk to the index cells inside it). Host logic (a handful of binary searches); returns the number of cells.
inline int region_cell_union_bound(const uint64_t* ids, size_t num_cells, uint64_t out[6]) {
  if (num_cells == 0) return 0;
  int count = 0;
  auto cover_range = [&](uint64_t first, uint64_t last) {
    if (count >= 6) return;
    out[count++] = first == last ? first : cellid_parent(first, cellid_common_ancestor_level(first, last));
  };
  const uint64_t last_index_id = ids[num_cells - 1];
  size_t it = 0;
  if (ids[0] != last_index_id) {
    const int level = cellid_common_ancestor_level(ids[0], last_index_id) + 1;
    const uint64_t last_id = cellid_parent(last_index_id, level);
    for (uint64_t id = cellid_parent(ids[0], level); id != last_id; id += cellid_lsb(id) << 1) {
      if (cellid_range_max(id) < ids[it]) continue;
      const uint64_t first = ids[it];
      const uint64_t past = cellid_range_max(id) + 2;               // range_max().next(): the leaf after the cell
      size_t lo = it, hi = num_cells;                               // Seek(past)
      while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (ids[mid] < past) lo = mid + 1; else hi = mid; }
      cover_range(first, ids[lo - 1]);
      it = lo;
    }
  }
  cover_range(ids[it], last_index_id);
  return count;
}

}  // namespace s2b
