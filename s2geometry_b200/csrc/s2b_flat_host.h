// s2b_flat_host.h — host-side validation and packing of an S2bFlatIndex (include/s2b_capi.h) into the array-of-
// struct form the kernels read (s2b_contains.cuh), including the bucket lookup table that replaces
// MutableS2ShapeIndex::Iterator::Seek's ordered-map search (mutable_s2shape_index.h:737-739).
#pragma once
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_contains.cuh"
#include "s2b_hilbert.h"

namespace s2b {

struct HostFlat {
  std::vector<FlatCell> cells;
  std::vector<FlatClip> clips;
  std::vector<FlatEdge> edges;
  std::vector<uint32_t> lut;   // num_buckets entries, (face, i, j) order
  std::vector<LutRange> mixed; // one per mixed bucket (at least one entry, so the array is never empty)
  std::vector<int32_t> interior_info;  // per cell, see FlatIndexView
  uint64_t num_buckets = 0;
  int lut_level = 0;
  FlatIndexView view(const uint64_t* cell_ids, int num_shape_ids) const {
    FlatIndexView v;
    v.cell_ids = cell_ids; v.cells = cells.data(); v.clips = clips.data(); v.edges = edges.data(); v.lut = lut.data(); v.mixed = mixed.data(); v.interior_info = interior_info.data();
    v.num_cells = (uint32_t)cells.size(); v.lut_level = lut_level; v.num_shape_ids = num_shape_ids;
    return v;
  }
};

inline int pack_flat_index(const S2bFlatIndex& f, HostFlat* out, std::string* err) {
  const uint64_t C = f.num_cells, K = f.num_clips, E = f.num_edges;
  auto fail = [&](int code, const std::string& m) { *err = m; return code; };
  if (C > 0xFFFFFFF0ull || K > 0xFFFFFFF0ull || E > 0xFFFFFFF0ull) return fail(S2B_EINVAL, "index too large for 32-bit offsets");
  if (C && (!f.cell_ids || !f.cell_clip_begin)) return fail(S2B_EINVAL, "null cell arrays");
  if (K && (!f.clip_shape_id || !f.clip_flags || !f.clip_edge_begin)) return fail(S2B_EINVAL, "null clip arrays");
  if (E && (!f.edge_v0 || !f.edge_v1)) return fail(S2B_EINVAL, "null edge arrays");
  for (uint64_t c = 0; c < C; ++c) {
    uint64_t id = f.cell_ids[c];
    if (id == 0 || (id >> 61) > 5 || !(cellid_lsb(id) & 0x1555555555555555ull))
      return fail(S2B_EINVAL, "cell_ids[" + std::to_string(c) + "] is not a valid S2CellId");
    if (c && cellid_range_max(f.cell_ids[c - 1]) >= cellid_range_min(id))
      return fail(S2B_EINVAL, "cell_ids must be ascending and disjoint (at " + std::to_string(c) + ")");
    if (f.cell_clip_begin[c] > f.cell_clip_begin[c + 1] || f.cell_clip_begin[c + 1] > K) return fail(S2B_EINVAL, "cell_clip_begin is not monotone");
  }
  for (uint64_t k = 0; k < K; ++k) {
    if (f.clip_edge_begin[k] > f.clip_edge_begin[k + 1] || f.clip_edge_begin[k + 1] > E) return fail(S2B_EINVAL, "clip_edge_begin is not monotone");
    if (f.clip_shape_id[k] < 0 || f.clip_shape_id[k] >= f.num_shape_ids) return fail(S2B_EINVAL, "clip_shape_id out of range");
  }
  // lookup table level: ~kLutFactor buckets per cell, between level 2 (96 buckets) and level 10 (6.3M buckets, 25 MB).
  // Finer buckets mean fewer "mixed" ones (a bucket at or below the level of the cells around it is never mixed), i.e.
  // more points decided by the single-precision pre-filter and fewer probes in the exact Locate.
  uint64_t lut_factor = 64;   // measured on B200: 4 -> 64 raises the dense ("near") workload from 32 to 44 Gpts/s at no cost to the others
  if (const char* e = getenv("S2B_LUT_FACTOR")) lut_factor = (uint64_t)atoi(e);
  int lut_level = 2;
  while (lut_level < 10 && (6ull << (2 * lut_level)) < lut_factor * C) ++lut_level;
  const uint64_t nb = 6ull << (2 * lut_level);
  out->num_buckets = nb;
  out->lut_level = lut_level;
  const int lut_shift = 61 - 2 * lut_level;   // Hilbert-order bucket of a leaf id = id >> lut_shift
  try {
    out->cells.resize(C); out->clips.resize(K); out->edges.resize(E); out->lut.resize(nb); out->interior_info.resize(C + 1);
  } catch (const std::bad_alloc&) {
    return fail(S2B_ENOMEM, "host allocation failed");
  }
  for (uint64_t c = 0; c < C; ++c) {
    P3 ctr;
    if (f.cell_center) ctr = P3{f.cell_center[3 * c], f.cell_center[3 * c + 1], f.cell_center[3 * c + 2]};
    else ctr = cellid_to_point(f.cell_ids[c], kHilbert.ij);
    out->cells[c] = FlatCell{ctr.x, ctr.y, ctr.z, f.cell_clip_begin[c], f.cell_clip_begin[c + 1] - f.cell_clip_begin[c]};
  }
  for (uint64_t k = 0; k < K; ++k)
    out->clips[k] = FlatClip{f.clip_shape_id[k], f.clip_edge_begin[k], f.clip_edge_begin[k + 1] - f.clip_edge_begin[k], f.clip_flags[k]};
  for (uint64_t e = 0; e < E; ++e)
    out->edges[e] = FlatEdge{f.edge_v0[3 * e], f.edge_v0[3 * e + 1], f.edge_v0[3 * e + 2], f.edge_v1[3 * e], f.edge_v1[3 * e + 1], f.edge_v1[3 * e + 2]};
  // Cells whose clips have no edges and all contain the centre (pure interior cells) are answered without looking
  // at clips or edges when only "any"/"how many" is asked.
  for (uint64_t c = 0; c < C; ++c) {
    uint32_t b = f.cell_clip_begin[c], e = f.cell_clip_begin[c + 1];
    bool pure = e > b;
    for (uint32_t k = b; k < e && pure; ++k)
      pure = f.clip_edge_begin[k + 1] == f.clip_edge_begin[k] && (f.clip_flags[k] & 1);
    out->interior_info[c] = !pure ? 0 : (e - b == 1 ? -(f.clip_shape_id[b] + 1) : (int32_t)(e - b));
  }
  // Walk the buckets in Hilbert order (b = leaf id >> lut_shift) with c = number of cells whose range_max < first leaf
  // of the bucket (the first cell a leaf of bucket b can be in), classify each (s2b_contains.cuh: empty / entirely
  // inside cell c / mixed) and store the entry at the bucket's (face, i, j) address.
  if (C > kLutCellMask) return fail(S2B_EINVAL, "index too large for the 29-bit cell field of the lookup-table entries");
  out->mixed.clear();
  uint64_t c = 0;
  for (uint64_t b = 0; b < nb; ++b) {
    const uint64_t first = b << lut_shift, last = ((b + 1) << lut_shift) - 1;
    while (c < C && cellid_range_max(f.cell_ids[c]) < first) ++c;
    uint32_t entry;
    if (c == C || cellid_range_min(f.cell_ids[c]) > last) entry = kLutEmpty;
    else if (cellid_range_min(f.cell_ids[c]) <= (first | 1) && cellid_range_max(f.cell_ids[c]) >= last)   // first | 1: the bucket's first LEAF id (leaf ids are odd)
      entry = out->interior_info[c] < 0 ? (kLutInterior | (uint32_t)(-out->interior_info[c] - 1))
                                        : (kLutInside | (uint32_t)c | (out->interior_info[c] == 0 ? kLutHasEdges : 0u));
    else {
      uint64_t h = c;   // last cell starting inside the bucket
      while (h + 1 < C && cellid_range_min(f.cell_ids[h + 1]) <= last) ++h;
      entry = kLutMixed | (uint32_t)out->mixed.size();
      out->mixed.push_back(LutRange{(uint32_t)c, (uint32_t)h});
    }
    int bi, bj;
    const int face = cellid_to_face_ij(first | 1, kHilbert.ij, &bi, &bj);   // the bucket's first leaf cell
    out->lut[lut_bucket(lut_level, face, (uint32_t)bi, (uint32_t)bj)] = entry;
  }
  if (out->mixed.empty()) out->mixed.push_back(LutRange{0, 0});
  return S2B_OK;
}

}  // namespace s2b
