// s2b_contains.cuh — Locate + containment over the flattened, device-resident shape index (host-compilable for
// tests/hostsim). Restates, per point:
//   S2CellIterator::LocateImpl          s2cell_iterator.h:137-155  (+ Iterator::Seek, mutable_s2shape_index.h:737-739)
//   S2ContainsPointQuery::ShapeContains s2contains_point_query.h:352-389
//   Contains / VisitContainingShapeIds  s2contains_point_query.h:255-295
//
// Flat layout (built on the host from the index's public iteration API, see s2b_capi.h S2bFlatIndex):
//   cell_ids[C]  sorted S2CellIds of the index cells (disjoint)         — the search array
//   cells[C]     cell centre (= S2CellId::ToPoint(), precomputed) + range of clipped shapes   32 B
//   clips[K]     shape id, first clipped edge, edge count, flags (bit0 contains_center, bits1-2 dimension)   16 B
//   edges[E]     both endpoints of every clipped edge, materialised so no shape is dereferenced          48 B
//   lut[B]       bucket = the level-lut_level cell containing the leaf, addressed by (face, i >> s, j >> s) — NOT by
//                Hilbert position, so the table is read before (and mostly instead of) the Hilbert encoding —
//                -> bucket kind + cell index; turns Seek's O(log C) binary search into one table read plus, for
//                buckets holding several cells, one range read and 0-3 probes.
//   mixed[M]     candidate cell range [lo, hi] of every bucket of kind "mixed"                           8 B
#pragma once
#include "s2b_core.cuh"
#include "s2b_predicates.cuh"

namespace s2b {

struct FlatCell { double cx, cy, cz; uint32_t clip_begin, clip_count; };
struct FlatClip { int32_t shape_id; uint32_t edge_begin; uint32_t num_edges; uint32_t flags; };
struct FlatEdge { double v0x, v0y, v0z, v1x, v1y, v1z; };

struct LutRange { uint32_t lo, hi; };
struct FlatIndexView {
  const uint64_t* cell_ids;
  const FlatCell* cells;
  const FlatClip* clips;
  const FlatEdge* edges;
  const uint32_t* lut;
  const LutRange* mixed;
  const int32_t* interior_info;    // per cell: 0 = has edges; pure interior cell (no edges, all clips contain the centre):
                                   //   -(shape_id + 1) if it has exactly one clip, else the number of clips
  uint32_t num_cells;
  int lut_level;                   // 2..10: buckets are the cells of this level
  int num_shape_ids;
  // Optional (device indexes): a Hilbert-order bucket table over cell_ids for Locate(S2CellId) — id_first[b] = number of
  // cells whose id >> id_shift is below b (id_buckets + 1 entries) — so that a target cell is located with one table read and
  // a search inside one bucket instead of a binary search over all cells.
  const uint32_t* id_first = nullptr;
  uint32_t id_buckets = 0;
  int id_shift = 0;
};

enum VertexModel { kOpen = 0, kSemiOpen = 1, kClosed = 2 };

// Lookup-table entry: bits 31..30 = what the bucket (a cell of level lut_level, i.e. a contiguous range of leaf ids)
// looks like; bits 29..0 = payload. Everything a query needs for a point in an "inside" bucket is in the entry itself,
// so that the pre-filter answers it with ONE load (a dependent second load — the cell's info — stalled every warp with
// at least one hit: 80 % of the warps on the 10k-polygon join):
//   kLutMixed  several index cells and/or gaps start inside the bucket: payload = index into mixed[], whose entry is
//              the range of cell_ids to probe
//   kLutEmpty  no index cell intersects the bucket: every leaf of it is a miss
//   kLutInside the whole bucket lies inside index cell (payload & kLutCellMask); bit kLutHasEdges says whether that cell
//              has clipped edges (-> crossing tests) or is a pure interior cell with several clipped shapes
//   kLutInterior  the whole bucket lies inside a pure interior cell with a SINGLE clipped shape (no edges, the shape
//              contains the centre): payload = that shape's id. "Is the point contained" / "by how many shapes" / "by
//              which shape" need nothing else; the cell's position, where a query asks for it, is found by searching
constexpr uint32_t kLutIndexMask = 0x3FFFFFFFu;
constexpr uint32_t kLutMixed = 0u, kLutEmpty = 1u << 30, kLutInside = 2u << 30, kLutInterior = 3u << 30;
constexpr uint32_t kLutHasEdges = 1u << 29, kLutCellMask = kLutHasEdges - 1u;

S2B_HD uint32_t lut_bucket(int lut_level, int face, uint32_t i, uint32_t j) {
  const int s = 30 - lut_level;
  return ((uint32_t)face << (2 * lut_level)) | ((i >> s) << lut_level) | (j >> s);
}

// Index of the cell containing the leaf cell (face, i, j), or -1. Equivalent to LocateImpl (s2cell_iterator.h:137-155):
// the index cells are disjoint, so the only candidate is the last cell whose range_min <= leaf. `leaf_id()` yields
// the leaf's S2CellId (the Hilbert encoding); it is only evaluated for mixed buckets.
template <class LeafFn>
S2B_HD int locate_cell(const FlatIndexView& ix, int face, uint32_t i, uint32_t j, LeafFn leaf_id) {
  if (ix.num_cells == 0) return -1;
  const uint32_t e0 = ix.lut[lut_bucket(ix.lut_level, face, i, j)];
  const uint32_t kind = e0 & ~kLutIndexMask;
  if (kind == kLutEmpty) return -1;
  if (kind == kLutInside) return (int)(e0 & kLutCellMask);
  // kLutInterior carries the shape, not the cell: search all cells (only callers that want the cell itself get here)
  LutRange r{0u, ix.num_cells - 1u};
  if (kind == kLutMixed) r = ix.mixed[e0 & kLutIndexMask];
  const uint64_t leaf = leaf_id();
  uint32_t lo = r.lo, hi = r.hi;
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (cellid_range_min(ix.cell_ids[mid]) <= leaf) lo = mid; else hi = mid - 1;
  }
  uint64_t id = ix.cell_ids[lo];
  return (cellid_range_min(id) <= leaf && leaf <= cellid_range_max(id)) ? (int)lo : -1;
}

// Locate for a point (4-bit Hilbert table), used by the host-side simulation and the generic visitor.
template <class HilbertT>
S2B_HD int locate_point(const FlatIndexView& ix, const P3& p, const HilbertT* hilbert_pos) {
  double fi, fj;
  const int face = point_to_face_fij(p, &fi, &fj);
  const uint32_t i = fij_to_ij(fi), j = fij_to_ij(fj);
  return locate_cell(ix, face, i, j, [&] { return from_face_ij(face, i, j, hilbert_pos); });
}

S2B_HD P3 edge_v0(const FlatEdge& e) { return P3{e.v0x, e.v0y, e.v0z}; }
S2B_HD P3 edge_v1(const FlatEdge& e) { return P3{e.v1x, e.v1y, e.v1z}; }

// Everything past the S2EdgeCrosser fast path for one edge, out of line and with by-value arguments so that the hot
// loop keeps its state in registers. Returns 0: no crossing, 1: crossing (toggle), 2: p is a vertex and the model is
// OPEN (result false), 3: p is a vertex and the model is CLOSED (result true)  (s2contains_point_query.h:375-385).
S2B_HD_NOINLINE int edge_slow_path(P3 a, P3 b, P3 c, P3 d, int acb, int bda, int model) {
  int sign = crossing_sign_slow(a, b, c, d, acb, bda);
  if (sign < 0) return 0;
  if (sign == 0) {
    if (model != kSemiOpen && (eq(c, b) || eq(d, b))) return model == kClosed ? 3 : 2;
    sign = vertex_crossing(a, b, c, d) ? 1 : 0;
  }
  return sign;
}

// ShapeContains(cell_id, clipped, p): a = cell centre, b = p, a_cross_b = a x b (hoisted: one per located cell).
S2B_HD bool clip_contains(const FlatIndexView& ix, const FlatClip& clip, const P3& a, const P3& b, const P3& a_cross_b, int model) {
  bool inside = (clip.flags & 1u) != 0;
  const uint32_t num_edges = clip.num_edges;
  if (num_edges == 0) return inside;
  const FlatEdge* edges = ix.edges + clip.edge_begin;
  const int dim = (int)((clip.flags >> 1) & 3u);
  if (dim < 2) {
    if (model != kClosed) return false;
    for (uint32_t i = 0; i < num_edges; ++i) {
      FlatEdge e = edges[i];
      if (eq(edge_v0(e), b) || eq(edge_v1(e), b)) return true;
    }
    return false;
  }
  // Pass 1: the S2EdgeCrosser fast path (s2edge_crosser.h:365-388) for every edge — C and D strictly on the same side
  // of AB means no crossing. Edges it cannot dismiss (~9 %) are only REMEMBERED here (bitmask per 64 edges) and
  // resolved in pass 2, so a warp's lanes go through the expensive path together instead of one at a time at
  // arbitrary iterations. The order of evaluation does not matter: the parity is a XOR, and the vertex-model early
  // return yields the same value whichever qualifying edge triggers it (s2contains_point_query.h:377-382).
  for (uint32_t base = 0; base < num_edges; base += 64) {
    const uint32_t cnt = num_edges - base < 64 ? num_edges - base : 64;
    uint64_t pending = 0;
    for (uint32_t i = 0; i < cnt; ++i) {
      const FlatEdge e = edges[base + i];
      const int acb = -triage_sign(a_cross_b, edge_v0(e));
      const int bda = triage_sign(a_cross_b, edge_v1(e));
      if (!(acb == -bda && bda != 0)) pending |= 1ull << i;
    }
    while (pending) {
      const int i = ctz64(pending);
      pending &= pending - 1;
      const FlatEdge e = edges[base + i];
      const P3 c = edge_v0(e), d = edge_v1(e);
      const int acb = -triage_sign(a_cross_b, c);
      const int bda = triage_sign(a_cross_b, d);
      // Common case of a pending edge: both signs certain and equal (C and D strictly on opposite sides of AB). Then
      // s2edge_crosser.cc:76-96 reduces to: shared vertex -> 0, degenerate edge -> -1, else compare with the signs
      // of B and A against CD. Done inline when those two are certain too; anything uncertain goes out of line.
      int r = -1;
      bool resolved = false;
      if (acb != 0 && bda != 0) {
        const bool shared = eq(a, c) || eq(a, d) || eq(b, c) || eq(b, d);
        const bool degenerate = eq(a, b) || eq(c, d);
        if (!shared && !degenerate) {
          const P3 c_cross_d = cross(c, d);
          const int cbd = -triage_sign(c_cross_d, b);
          const int dac = triage_sign(c_cross_d, a);
          if (cbd != 0 && cbd != acb) { r = 0; resolved = true; }
          else if (cbd != 0 && dac != 0) { r = (dac == acb) ? 1 : 0; resolved = true; }
        }
      }
      if (!resolved) r = edge_slow_path(a, b, c, d, acb, bda, model);
      if (r >= 2) return r == 3;
      inside ^= (r != 0);
    }
  }
  return inside;
}

// Calls visit(shape_id) for every shape that contains p, in ascending shape id order, until visit returns false.
// Returns the located cell index (or -1).
template <class HilbertT, class Visit>
S2B_HD int visit_containing_shapes(const FlatIndexView& ix, const P3& p, const HilbertT* hilbert_pos, int model, Visit visit) {
  int cell = locate_point(ix, p, hilbert_pos);
  if (cell < 0) return cell;
  FlatCell fc = ix.cells[cell];
  P3 a{fc.cx, fc.cy, fc.cz};
  P3 a_cross_b = cross(a, p);
  for (uint32_t k = 0; k < fc.clip_count; ++k) {
    FlatClip clip = ix.clips[fc.clip_begin + k];
    if (clip_contains(ix, clip, a, p, a_cross_b, model)) {
      if (!visit(clip.shape_id)) break;
    }
  }
  return cell;
}

}  // namespace s2b
