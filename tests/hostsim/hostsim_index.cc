// tests/hostsim/hostsim_index.cc — TEST-ONLY host execution of the flat-index query logic (s2b_contains.cuh) so the
// Locate / lookup-table / crossing-parity code is checked against the oracle without a GPU.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <map>

#include "s2b_flat_host.h"
#include "s2b_region.cuh"

using namespace s2b;

extern "C" {

// mode 0: Contains (any), 1: ShapeContains(shape_id), 2: histogram into counts[num_shape_ids] (u64),
// 3: located cell index per point (int32, -1 = miss) into out32, 4: per-point containing-shape count into out32.
int hs_index_query(const S2bFlatIndex* f, const double* xyz, size_t n, int model, int mode, int shape_id, uint8_t* out8,
                   uint64_t* counts, int32_t* out32, uint64_t* n_exact_out, char* err, size_t errlen) {
  HostFlat hf;
  std::string e;
  int rc = pack_flat_index(*f, &hf, &e);
  if (rc) { if (err) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; } return rc; }
  FlatIndexView ix = hf.view(f->cell_ids, f->num_shape_ids);
  unsigned long long before = host_exact_count();
  for (size_t i = 0; i < n; ++i) {
    P3 p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    if (mode == 0) {
      bool any = false;
      visit_containing_shapes(ix, p, kHilbert.pos, model, [&](int) { any = true; return false; });
      out8[i] = any;
    } else if (mode == 1) {
      bool hit = false;
      visit_containing_shapes(ix, p, kHilbert.pos, model, [&](int id) { if (id == shape_id) hit = true; return id < shape_id; });
      out8[i] = hit;
    } else if (mode == 2) {
      visit_containing_shapes(ix, p, kHilbert.pos, model, [&](int id) { ++counts[id]; return true; });
    } else if (mode == 3) {
      out32[i] = locate_point(ix, p, kHilbert.pos);
    } else {
      int c = 0;
      visit_containing_shapes(ix, p, kHilbert.pos, model, [&](int) { ++c; return true; });
      out32[i] = c;
    }
  }
  if (n_exact_out) *n_exact_out = host_exact_count() - before;
  return 0;
}

void hs_cellid_to_point(const uint64_t* ids, size_t n, double* xyz) {
  for (size_t i = 0; i < n; ++i) { P3 p = cellid_to_point(ids[i], kHilbert.ij); xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z; }
}

// S2ShapeIndexRegion cell tests (s2b_region.cuh). mode 0: MayIntersect(S2Cell), 1: Contains(S2Cell).
int hs_region_cells(const S2bFlatIndex* f, const uint64_t* ids, size_t n, int mode, uint8_t* out, uint64_t* unsupported_edges) {
  HostFlat hf;
  std::string e;
  int rc = pack_flat_index(*f, &hf, &e);
  if (rc) return rc;
  FlatIndexView ix = hf.view(f->cell_ids, f->num_shape_ids);
  uint64_t bad = 0;
  for (const FlatEdge& ed : hf.edges) if (!region_edge_is_plain(edge_v0(ed), edge_v1(ed))) ++bad;   // statistic only
  if (unsupported_edges) *unsupported_edges = bad;
  for (size_t i = 0; i < n; ++i)
    out[i] = mode == 0 ? region_may_intersect(ix, ids[i], kHilbert.ij) : region_contains_cell(ix, ids[i], kHilbert.ij);
  return 0;
}

// VisitIntersectingShapeIds with the same per-target decisions as region_shapes_kernel; the sort/reduce of the raw
// tuples is a std::map here. Returns the number of pairs; writes them when capacity suffices.
size_t hs_region_intersecting_shapes(const S2bFlatIndex* f, const uint64_t* ids, size_t n, uint32_t* offsets, int32_t* shape_ids,
                                     uint8_t* contains, size_t capacity) {
  HostFlat hf;
  std::string e;
  if (pack_flat_index(*f, &hf, &e)) return 0;
  FlatIndexView ix = hf.view(f->cell_ids, f->num_shape_ids);
  size_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    offsets[i] = (uint32_t)total;
    std::map<int, bool> not_contains;
    int32_t at;
    const int rel = locate_target(ix, ids[i], &at);
    if (rel == 1) {
      const uint64_t tmax = cellid_range_max(ids[i]);
      for (uint32_t c = (uint32_t)at; c < ix.num_cells && ix.cell_ids[c] <= tmax; ++c)
        for (uint32_t k = 0; k < ix.cells[c].clip_count; ++k) {
          const FlatClip clip = ix.clips[ix.cells[c].clip_begin + k];
          not_contains[clip.shape_id] |= clip.num_edges > 0 || !(clip.flags & 1u);
        }
    } else if (rel == 0) {
      const FlatCell fc = ix.cells[at];
      const bool same = ix.cell_ids[at] == ids[i];
      const TargetCell tc = target_cell(ids[i], kHilbert.ij);
      const P3 a{fc.cx, fc.cy, fc.cz}, centre = cellid_to_point(ids[i], kHilbert.ij), axb = cross(a, centre);
      for (uint32_t k = 0; k < fc.clip_count; ++k) {
        const FlatClip clip = ix.clips[fc.clip_begin + k];
        const int r = region_clip_relation(ix, clip, same, tc, a, centre, axb);
        if (r) not_contains[clip.shape_id] = r != 2;
      }
    }
    for (auto& kv : not_contains) {
      if (total < capacity) { shape_ids[total] = kv.first; contains[total] = !kv.second; }
      ++total;
    }
  }
  offsets[n] = (uint32_t)total;
  return total;
}

void hs_clip_to_padded_face(const double* a, const double* b, const int32_t* face, size_t n, double padding, double* out_uv, uint8_t* ok,
                            uint8_t* supported) {
  for (size_t i = 0; i < n; ++i) {
    const P3 pa{a[3 * i], a[3 * i + 1], a[3 * i + 2]}, pb{b[3 * i], b[3 * i + 1], b[3 * i + 2]};
    Uv p0{0, 0}, p1{0, 0};
    supported[i] = region_edge_is_plain(pa, pb);   // 0: RobustCrossProd went beyond its double-precision stage
    ok[i] = clip_to_padded_face(pa, pb, face[i], padding, &p0, &p1);
    out_uv[4 * i] = p0.u; out_uv[4 * i + 1] = p0.v; out_uv[4 * i + 2] = p1.u; out_uv[4 * i + 3] = p1.v;
  }
}
void hs_intersects_rect(const double* uv4, const double* rect4, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    const double* r = rect4 + 4 * i;
    out[i] = intersects_rect(Uv{uv4[4 * i], uv4[4 * i + 1]}, Uv{uv4[4 * i + 2], uv4[4 * i + 3]}, UvRect{r[0], r[1], r[2], r[3]});
  }
}

// S2::RobustCrossProd, all stages. stage[i]: 0 double, 1 a == b, 2 long double, 3 exact / symbolic.
void hs_robust_cross_prod(const double* a, const double* b, size_t n, double* out, uint8_t* stage) {
  for (size_t i = 0; i < n; ++i) {
    const P3 pa{a[3 * i], a[3 * i + 1], a[3 * i + 2]}, pb{b[3 * i], b[3 * i + 1], b[3 * i + 2]};
    const P3 r = robust_cross_prod(pa, pb);
    out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    P3 t;
    stage[i] = stable_cross_prod(pa, pb, &t) ? 0 : (eq(pa, pb) ? 1 : (stable_cross_prod_ld(pa, pb, &t) ? 2 : 3));
  }
}

int hs_region_cell_union_bound(const uint64_t* ids, size_t n, uint64_t* out6) { return region_cell_union_bound(ids, n, out6); }

}  // extern "C"
