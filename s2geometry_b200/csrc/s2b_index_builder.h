// s2b_index_builder.h — host-side construction of the flattened shape index directly from shapes, without the
// reference library. It produces the same KIND of structure as MutableS2ShapeIndex (mutable_s2shape_index.cc:
// 1260-1925): disjoint S2 cells in Hilbert order, each with the clipped edge lists of the shapes that intersect it
// and a contains_center bit per shape. The two indexes need not be cell-for-cell identical: query results depend only
// on (a) every edge that can cross "cell centre -> query point" being listed in the cell, and (b) correct
// contains_center bits, because all predicates are exact. Design differences from the reference:
//
//  * edge/cell intersection is decided in 3-D, not by clipping in (u,v): for an edge A->B and a face, a point of the
//    chord P(t) = A + t(B-A) projects into the padded cell [u0,u1]x[v0,v1] iff  u0*w(t) <= pu(t) <= u1*w(t),
//    v0*w(t) <= pv(t) <= v1*w(t), w(t) > 0, where (pu,pv,w) are the face-frame coordinates, all LINEAR in t. The
//    surviving parameter interval [t0,t1] is handed down the recursion and only shrinks. Intervals are widened by a
//    slack, so the edge lists are conservative supersets (harmless), never subsets.
//  * one recursive descent per face in Hilbert child order emits cells already sorted by id; there is no ordered map.
//  * contains_center follows the reference's InteriorTracker idea (mutable_s2shape_index.cc:246-432): a focus point
//    walks entry vertex -> centre -> exit vertex of every edge-bearing cell (consecutive cells share that vertex, cells
//    without edges cannot change the state), toggling a shape whenever the step crosses one of its edges
//    (EdgeOrVertexCrossing with the exact predicates of s2b_predicates.cuh, compiled for the host).
//  * the initial state comes from a reference point per polygon, restating s2shapeutil::GetReferencePoint
//    (s2shapeutil_get_reference_point.cc:38-108) and S2ContainsVertexQuery::ContainsSign (s2contains_vertex_query.cc:30-49).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_contains.cuh"
#include "s2b_hilbert.h"

namespace s2b {

struct BuiltIndex {
  std::vector<uint64_t> cell_ids;
  std::vector<double> cell_center;       // C*3
  std::vector<uint32_t> cell_clip_begin; // C+1
  std::vector<int32_t> clip_shape_id;
  std::vector<uint8_t> clip_flags;
  std::vector<uint32_t> clip_edge_begin; // K+1
  std::vector<int32_t> edge_id;          // E (shape-local edge ids, for VisitIncidentEdges-style consumers)
  std::vector<double> edge_v0, edge_v1;  // E*3
  int32_t num_shape_ids = 0;
  S2bFlatIndex as_flat() const {
    S2bFlatIndex f;
    f.num_cells = cell_ids.size(); f.num_clips = clip_shape_id.size(); f.num_edges = edge_id.size();
    f.num_shape_ids = num_shape_ids;
    f.cell_ids = cell_ids.data(); f.cell_center = cell_center.data(); f.cell_clip_begin = cell_clip_begin.data();
    f.clip_shape_id = clip_shape_id.data(); f.clip_flags = clip_flags.data(); f.clip_edge_begin = clip_edge_begin.data();
    f.edge_v0 = edge_v0.data(); f.edge_v1 = edge_v1.data();
    return f;
  }
};

namespace builder_detail {

struct Edge {
  P3 v0, v1;
  int32_t shape_id, edge_id;
  int8_t dim;
  int8_t max_level;
};

// Face frame: (pu, pv, w) of a point for face f such that u = pu/w, v = pv/w  (s2coords.h:389-403).
inline void face_frame(int face, const P3& p, double* pu, double* pv, double* w) {
  switch (face) {
    case 0: *pu =  p.y; *pv =  p.z; *w =  p.x; break;
    case 1: *pu = -p.x; *pv =  p.z; *w =  p.y; break;
    case 2: *pu = -p.x; *pv = -p.y; *w =  p.z; break;
    case 3: *pu = -p.z; *pv = -p.y; *w = -p.x; break;   // z/x with x<0: (-z)/(-x)
    case 4: *pu = -p.z; *pv =  p.x; *w = -p.y; break;
    default: *pu =  p.y; *pv =  p.x; *w = -p.z; break;
  }
}

struct FrameEdge {      // an edge expressed in one face's frame: value(t) = a + t*(b - a)
  double au, av, aw, bu, bv, bw;
  int index;            // into the global edge array
  double t0, t1;        // parameter interval still alive
};

// Why the three constants below make every cell's edge list a SUPERSET of the edges that meet the cell (what point
// queries need: the segment from a cell's centre to a query point of that cell stays inside the cell's uv rectangle).
// With eps = 2^-53 and all frame coordinates, bounds and their products bounded by 1 in magnitude:
//  (1) Each constraint value f = pu - ulo*w (etc.) is computed with one product and one difference:
//      |f_computed - f| <= eps*|ulo*w| + eps*|f_computed| <= 3 eps < 4e-16 =: delta. The constraint is linear along the
//      chord, so the line through the computed end values lies within delta of the true one on all of [0,1].
//  (2) kSlack = 1e-13 > 250 delta is added to both end values (rounding of that sum: <= 2 eps, absorbed). Wherever the
//      TRUE f(t) >= 0, the slackened computed line is >= kSlack - delta - 2 eps > 9.9e-14 > 0. Its slope is at most
//      |f1 - f0| <= 4, so every truly feasible t lies at least 9.9e-14 / 4 > 2.4e-14 inside the computed half-line.
//  (3) The root tz = F0 / (F0 - F1) is one difference (F0 and F1 have opposite signs: no cancellation) and one
//      division: relative error <= 3 eps, absolute <= 3.4e-16 since tz is in [0,1] — 70 times below the 2.4e-14 margin of (2). The extra
//      1e-12 by which the live interval is widened is therefore not needed for correctness; it keeps the interval
//      conservative when children re-clip an interval inherited from their parent.
//  (4) kPad = 1e-14 widens the rectangle in uv before (1). It plays the part of the reference's kCellPadding
//      (3.9e-15, mutable_s2shape_index.cc:183-185) and covers what (1)-(3) do not: a query point is assigned to its cell
//      by the FLOATING-POINT chain XYZtoFaceUV -> UVtoST -> STtoIJ, whose uv error is a few eps (< 1e-15), so a point of
//      the cell can lie that far outside the exact rectangle [st_to_uv(i/2^30), ...] the clipping uses.
// tests/test_index_cpu.py::test_own_builder_edge_lists_are_supersets_of_the_exact_intersections checks the conclusion
// in exact rational arithmetic on sampled cells (no truly intersecting edge is ever missing, lists <= 1.5x the exact ones).
constexpr double kPad = 1e-14;       // uv padding, (4)
constexpr double kSlack = 1e-13;     // slack on the linear inequalities, (2)
constexpr double kRootSlack = 1e-12; // widening of the live parameter interval at a root, (3)
static_assert(kSlack > 100 * 4e-16 && kRootSlack > 1e3 * 3.4e-16 && kPad > 2 * 3.9e-15, "see the derivation above");

// Intersects [t0,t1] with { t : f0 + t*(f1 - f0) >= -slack }. Returns false if empty.
inline bool clip_halfline(double f0, double f1, double* t0, double* t1) {
  f0 += kSlack; f1 += kSlack;
  if (f0 >= 0 && f1 >= 0) return true;
  if (f0 < 0 && f1 < 0) {
    // negative at both ends of [0,1]; may still be non-negative nowhere in [t0,t1]
    return false;
  }
  double tz = f0 / (f0 - f1);          // root in [0,1]
  if (f0 < 0) { double lo = tz - kRootSlack; if (lo > *t0) *t0 = lo; }
  else        { double hi = tz + kRootSlack; if (hi < *t1) *t1 = hi; }
  return *t0 <= *t1;
}

// Restricts e's live interval to the padded rectangle [ulo,uhi]x[vlo,vhi] of its face. Returns false if it misses it.
inline bool clip_to_rect(const FrameEdge& e, double ulo, double uhi, double vlo, double vhi, double* t0, double* t1) {
  *t0 = e.t0; *t1 = e.t1;
  ulo -= kPad; uhi += kPad; vlo -= kPad; vhi += kPad;
  // pu - ulo*w >= 0 ; uhi*w - pu >= 0 ; pv - vlo*w >= 0 ; vhi*w - pv >= 0 ; w >= 0
  if (!clip_halfline(e.aw, e.bw, t0, t1)) return false;
  if (!clip_halfline(e.au - ulo * e.aw, e.bu - ulo * e.bw, t0, t1)) return false;
  if (!clip_halfline(uhi * e.aw - e.au, uhi * e.bw - e.bu, t0, t1)) return false;
  if (!clip_halfline(e.av - vlo * e.aw, e.bv - vlo * e.bw, t0, t1)) return false;
  if (!clip_halfline(vhi * e.aw - e.av, vhi * e.bw - e.bv, t0, t1)) return false;
  return true;
}

inline double ij_to_uv(uint32_t ij) {   // boundary coordinate of leaf index ij (0..2^30): STtoUV(IJtoSTMin(ij))
  return st_to_uv((1.0 / 1073741824.0) * (double)ij);
}

// S2::kAvgEdge.GetLevelForMaxValue(length) (s2metrics.h:169-183; kAvgEdge deriv for the quadratic projection).
inline int edge_max_level(const P3& a, const P3& b) {
  P3 d = sub(a, b);
  double value = sqrt(norm2(d));   // cell_size_to_long_edge_ratio = 1.0
  if (!(value > 0)) return 30;
  int level = std::ilogb(value / 1.459213746386106062);
  level = -level;
  return level < 0 ? 0 : (level > 30 ? 30 : level);
}

inline bool edge_or_vertex_crossing(const P3& a, const P3& b, const P3& c, const P3& d) {
  int s = crossing_sign(a, b, c, d);
  if (s < 0) return false;
  if (s > 0) return true;
  return vertex_crossing(a, b, c, d);
}

struct LexLess { bool operator()(const P3& a, const P3& b) const { return lex_less(a, b); } };

// S2ContainsVertexQuery (s2contains_vertex_query.cc): +1 if the shape contains vtest, -1 if not, 0 if every incident
// edge is matched by its reverse.
inline int contains_vertex_sign(const std::vector<Edge>& edges, size_t begin, size_t end, const P3& vtest) {
  std::map<P3, int, LexLess> dirs;
  for (size_t e = begin; e < end; ++e) {
    if (eq(edges[e].v0, vtest)) dirs[edges[e].v1] += 1;
    if (eq(edges[e].v1, vtest)) dirs[edges[e].v0] -= 1;
  }
  P3 ref_dir = ortho(vtest);
  P3 best = ref_dir;
  int best_sign = 0;
  for (const auto& kv : dirs) {
    if (kv.second == 0) continue;
    if (ordered_ccw(ref_dir, best, kv.first, vtest)) { best = kv.first; best_sign = kv.second; }
  }
  return best_sign;
}

struct RefPoint { P3 point; bool contained; };

// s2shapeutil::GetReferencePoint for the polygon whose edges are edges[begin,end). has_empty_chain: the shape has a
// chain with no vertices (the "full polygon" convention).
inline RefPoint reference_point(const std::vector<Edge>& edges, size_t begin, size_t end, bool has_any_chain, bool has_empty_chain) {
  const P3 origin{-0.0099994664350250197, 0.0025924542609324121, 0.99994664350250195};  // S2::Origin()
  if (begin == end) return RefPoint{origin, has_any_chain};
  int s = contains_vertex_sign(edges, begin, end, edges[begin].v0);
  if (s != 0) return RefPoint{edges[begin].v0, s > 0};
  struct E2 { P3 a, b; };
  auto less = [](const E2& x, const E2& y) { return lex_less(x.a, y.a) || (eq(x.a, y.a) && lex_less(x.b, y.b)); };
  std::vector<E2> fwd, rev;
  for (size_t e = begin; e < end; ++e) { fwd.push_back({edges[e].v0, edges[e].v1}); rev.push_back({edges[e].v1, edges[e].v0}); }
  std::sort(fwd.begin(), fwd.end(), less);
  std::sort(rev.begin(), rev.end(), less);
  for (size_t i = 0; i < fwd.size(); ++i) {
    if (less(fwd[i], rev[i])) { int s2 = contains_vertex_sign(edges, begin, end, fwd[i].a); return RefPoint{fwd[i].a, s2 > 0}; }
    if (less(rev[i], fwd[i])) { int s2 = contains_vertex_sign(edges, begin, end, rev[i].a); return RefPoint{rev[i].a, s2 > 0}; }
  }
  return RefPoint{origin, has_empty_chain};
}

struct Tracker {                       // which polygons contain the focus point
  P3 focus;
  std::vector<int32_t> inside;         // sorted shape ids
  uint64_t next_leaf = 1;              // first leaf id of the cell whose entry vertex the focus sits on (InteriorTracker::at_cellid)
  void toggle(int32_t id) {
    auto it = std::lower_bound(inside.begin(), inside.end(), id);
    if (it != inside.end() && *it == id) inside.erase(it); else inside.insert(it, id);
  }
};

struct Builder {
  const std::vector<Edge>& edges;
  int max_edges_per_cell;
  BuiltIndex* out;
  Tracker tracker;
  std::vector<std::vector<FrameEdge>> pool;   // per-depth scratch lists

  Builder(const std::vector<Edge>& e, int max_edges, BuiltIndex* o) : edges(e), max_edges_per_cell(max_edges), out(o), pool(32 * 4 + 8) {}

  static P3 vertex_at(int face, uint32_t i, uint32_t j) {          // FaceSiTitoXYZ(face, 2i, 2j).Normalize()
    return normalize(face_uv_to_xyz(face, ij_to_uv(i), ij_to_uv(j)));
  }

  void move_focus(const P3& to, const std::vector<FrameEdge>& cell_edges) {   // DrawTo + TestAllEdges
    const P3 from = tracker.focus;
    if (!eq(from, to)) {
      for (const FrameEdge& fe : cell_edges) {
        const Edge& e = edges[fe.index];
        if (e.dim != 2) continue;
        if (edge_or_vertex_crossing(from, to, e.v0, e.v1)) tracker.toggle(e.shape_id);
      }
    }
    tracker.focus = to;
  }

  void emit_cell(uint64_t id, const P3& center, const std::vector<FrameEdge>& cell_edges, const std::vector<int32_t>& containing) {
    out->cell_ids.push_back(id);
    out->cell_center.push_back(center.x); out->cell_center.push_back(center.y); out->cell_center.push_back(center.z);
    // merge the (shape-sorted) edge list with the containing-shape set
    size_t en = 0, cn = 0;
    while (en < cell_edges.size() || cn < containing.size()) {
      int32_t es = en < cell_edges.size() ? edges[cell_edges[en].index].shape_id : INT32_MAX;
      int32_t cs = cn < containing.size() ? containing[cn] : INT32_MAX;
      int32_t sid = es < cs ? es : cs;
      uint8_t flags = 0;
      int dim = 2;
      out->clip_shape_id.push_back(sid);
      if (cs == sid) { flags |= 1; ++cn; }
      while (en < cell_edges.size() && edges[cell_edges[en].index].shape_id == sid) {
        const Edge& e = edges[cell_edges[en].index];
        dim = e.dim;
        out->edge_id.push_back(e.edge_id);
        out->edge_v0.push_back(e.v0.x); out->edge_v0.push_back(e.v0.y); out->edge_v0.push_back(e.v0.z);
        out->edge_v1.push_back(e.v1.x); out->edge_v1.push_back(e.v1.y); out->edge_v1.push_back(e.v1.z);
        ++en;
      }
      flags |= (uint8_t)(dim << 1);
      out->clip_flags.push_back(flags);
      out->clip_edge_begin.push_back((uint32_t)out->edge_id.size());
    }
    out->cell_clip_begin.push_back((uint32_t)out->clip_shape_id.size());
  }

  // cell = (face, level, i_lo, j_lo) with Hilbert orientation `orient`; id_prefix holds the position bits so far.
  void build(int face, int level, uint32_t i_lo, uint32_t j_lo, int orient, uint64_t pos_bits, std::vector<FrameEdge>& cell_edges, int depth) {
    const uint32_t size = 1u << (30 - level);
    const uint64_t lsb = 1ull << (2 * (30 - level));
    const uint64_t id = ((((uint64_t)face << 60) | pos_bits) << 1) | lsb;   // pos_bits already aligned to the top
    if (cell_edges.empty()) {
      if (tracker.inside.empty()) return;
      P3 center = cellid_to_point(id, kHilbert.ij);
      emit_cell(id, center, cell_edges, tracker.inside);     // interior cell: no edges, containment = tracker state
      return;
    }
    bool subdivide = false;
    if ((int)cell_edges.size() > max_edges_per_cell && level < 30) {
      int max_short = std::max(max_edges_per_cell, (int)(0.2 * (double)(cell_edges.size() + tracker.inside.size())));
      int count = 0;
      for (const FrameEdge& fe : cell_edges) {
        count += level < edges[fe.index].max_level;
        if (count > max_short) { subdivide = true; break; }
      }
    }
    if (!subdivide) {
      // entry vertex: (i_lo, j_lo), or the opposite corner when the curve is inverted (s2padded_cell.cc:100-111)
      uint32_t ei = i_lo, ej = j_lo;
      if (orient & 2) { ei += size; ej += size; }
      if (tracker.next_leaf != cellid_range_min(id)) tracker.focus = vertex_at(face, ei, ej);   // MoveTo: skipped cells have no edges
      P3 center = cellid_to_point(id, kHilbert.ij);
      move_focus(center, cell_edges);
      std::vector<int32_t> containing = tracker.inside;
      emit_cell(id, center, cell_edges, containing);
      uint32_t xi = i_lo, xj = j_lo;                                              // exit vertex (s2padded_cell.cc:113-125)
      if (orient == 0 || orient == 3) xi += size; else xj += size;
      move_focus(vertex_at(face, xi, xj), cell_edges);
      tracker.next_leaf = id + lsb + 1;                                           // range_min of S2CellId::next()
      return;
    }
    // children in Hilbert order: pos -> (di,dj) for this orientation; child orientation = orient ^ modifier[pos]
    static const int kPosToIJ[4][4] = {{0, 1, 3, 2}, {0, 2, 3, 1}, {3, 2, 0, 1}, {3, 1, 0, 2}};
    static const int kPosToOrient[4] = {1, 0, 0, 3};
    const uint32_t half = size >> 1;
    for (int pos = 0; pos < 4; ++pos) {
      int ij = kPosToIJ[orient][pos];
      uint32_t ci = i_lo + ((ij >> 1) ? half : 0), cj = j_lo + ((ij & 1) ? half : 0);
      double ulo = ij_to_uv(ci), uhi = ij_to_uv(ci + half), vlo = ij_to_uv(cj), vhi = ij_to_uv(cj + half);
      std::vector<FrameEdge>& child = pool[depth + 1];
      child.clear();
      for (const FrameEdge& fe : cell_edges) {
        double t0, t1;
        if (clip_to_rect(fe, ulo, uhi, vlo, vhi, &t0, &t1)) { FrameEdge c = fe; c.t0 = t0; c.t1 = t1; child.push_back(c); }
      }
      uint64_t child_pos = pos_bits | ((uint64_t)pos << (2 * (29 - level)));
      build(face, level + 1, ci, cj, orient ^ kPosToOrient[pos], child_pos, child, depth + 1);
    }
  }
};

}  // namespace builder_detail

// Shape soup (see s2b_capi.h s2b_index_build). Returns S2B_OK or S2B_EINVAL with *err set.
inline int build_index_from_shapes(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                                   const uint32_t* chain_vertex_begin, const double* vertices, int max_edges_per_cell,
                                   BuiltIndex* out, std::string* err) {
  using namespace builder_detail;
  if (num_shapes < 0 || (num_shapes && (!shape_dim || !shape_chain_begin || !chain_vertex_begin))) { *err = "null shape arrays"; return S2B_EINVAL; }
  if (max_edges_per_cell <= 0) max_edges_per_cell = 10;   // --s2shape_index_default_max_edges_per_cell
  *out = BuiltIndex();
  out->num_shape_ids = num_shapes;
  out->cell_clip_begin.push_back(0);
  out->clip_edge_begin.push_back(0);
  auto V = [&](uint32_t k) { return P3{vertices[3 * (size_t)k], vertices[3 * (size_t)k + 1], vertices[3 * (size_t)k + 2]}; };

  std::vector<Edge> edges;
  std::vector<size_t> shape_edge_begin(num_shapes + 1, 0);
  std::vector<char> any_chain(num_shapes, 0), empty_chain(num_shapes, 0);
  for (int s = 0; s < num_shapes; ++s) {
    int dim = shape_dim[s];
    if (dim < 0 || dim > 2) { *err = "shape dimension must be 0, 1 or 2"; return S2B_EINVAL; }
    shape_edge_begin[s] = edges.size();
    int32_t eid = 0;
    for (uint32_t c = shape_chain_begin[s]; c < shape_chain_begin[s + 1]; ++c) {
      uint32_t b = chain_vertex_begin[c], e = chain_vertex_begin[c + 1];
      if (e < b) { *err = "chain_vertex_begin is not monotone"; return S2B_EINVAL; }
      uint32_t n = e - b;
      any_chain[s] = 1;
      if (n == 0) empty_chain[s] = 1;
      uint32_t ne = dim == 2 ? n : (dim == 1 ? (n ? n - 1 : 0) : n);
      for (uint32_t k = 0; k < ne; ++k) {
        Edge ed;
        ed.v0 = V(b + k);
        ed.v1 = dim == 0 ? ed.v0 : V(b + ((k + 1 == n) ? 0 : k + 1));
        ed.shape_id = s; ed.edge_id = eid++; ed.dim = (int8_t)dim;
        ed.max_level = (int8_t)edge_max_level(ed.v0, ed.v1);
        edges.push_back(ed);
      }
    }
  }
  shape_edge_begin[num_shapes] = edges.size();

  Builder b(edges, max_edges_per_cell, out);
  // Start of the Hilbert curve: the (-1,-1) corner of face 0; initial containment by crossing parity from each
  // polygon's reference point (ContainsBruteForce, s2shapeutil_contains_brute_force.cc:26-40).
  b.tracker.focus = normalize(face_uv_to_xyz(0, -1.0, -1.0));
  b.tracker.next_leaf = 1;   // S2CellId::Begin(kMaxLevel)
  for (int s = 0; s < num_shapes; ++s) {
    if (shape_dim[s] != 2) continue;
    RefPoint rp = reference_point(edges, shape_edge_begin[s], shape_edge_begin[s + 1], any_chain[s] != 0, empty_chain[s] != 0);
    bool inside = rp.contained;
    if (!eq(rp.point, b.tracker.focus)) {
      for (size_t e = shape_edge_begin[s]; e < shape_edge_begin[s + 1]; ++e)
        inside ^= edge_or_vertex_crossing(rp.point, b.tracker.focus, edges[e].v0, edges[e].v1);
    }
    if (inside) b.tracker.inside.push_back(s);
  }

  // Per face: the edges that can touch the (padded) face, in its frame.
  for (int face = 0; face < 6; ++face) {
    std::vector<FrameEdge>& list = b.pool[0];
    list.clear();
    for (size_t k = 0; k < edges.size(); ++k) {
      FrameEdge fe;
      face_frame(face, edges[k].v0, &fe.au, &fe.av, &fe.aw);
      face_frame(face, edges[k].v1, &fe.bu, &fe.bv, &fe.bw);
      fe.index = (int)k; fe.t0 = 0; fe.t1 = 1;
      double t0, t1;
      if (clip_to_rect(fe, -1.0, 1.0, -1.0, 1.0, &t0, &t1)) { fe.t0 = t0; fe.t1 = t1; list.push_back(fe); }
    }
    // face cells alternate orientation: face & kSwapMask (s2cell_id.cc:274-279)
    b.build(face, 0, 0, 0, face & 1, 0, list, 0);
  }
  return S2B_OK;
}

}  // namespace s2b
