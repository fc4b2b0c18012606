// s2b_cellindex.cuh — flat S2CellIndex (s2cell_index.h:113-441, s2cell_index.cc:71-146) and the traversal every batched
// query over it shares. Integer work only, bit-exact with the reference.
//
// The reference keeps the (cell_id, label) pairs as a tree (`cell_tree_`, parent = the nearest enclosing pair) in the order
// its Build() pushes them: by range_min, larger cells first, then by label (the `Delta` order, s2cell_index.cc:79-99) — a
// preorder of the containment forest — plus a list of leaf-cell ranges pointing into the tree. Here the tree is three
// parallel arrays in that same order and the range list is replaced by a bucket table over the Hilbert curve:
//
//   ids[m], labels[m]   the pairs in tree order
//   parent[m]           nearest enclosing pair (an equal cell with a smaller label counts), -1 for roots
//   first[B + 2]        first[b] = number of pairs whose range_min lies before bucket b (cells of table_level in Hilbert order)
//
// The pairs that intersect a target cell T (what VisitIntersectingCells reports for it, s2cell_index.h:619-647) are
//   (a) T itself and its descendants: one contiguous run [lo, hi) of the tree order, found with two searches inside a bucket;
//   (b) T's proper ancestors: the parent chain that starts at the deepest pair containing T. The pair just before `lo` lies
//       inside that ancestor (or is it), so the chain is reached by walking up from lo - 1 until a pair contains T.
// For a target union (cells sorted and non-overlapping, as S2CellUnion requires) each pair is reported once: runs of
// different target cells are disjoint, and a chain stops at the first pair whose position is not above the start of an
// earlier chain — the preorder argument of ContentsIterator::StartUnion (s2cell_index.cc:38-66).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "s2b_core.cuh"

namespace s2b {

struct CellIndexView {
  const uint64_t* ids;
  const int32_t* labels;
  const int32_t* parent;
  const uint32_t* first;
  uint32_t m;
  int shift;             // 61 - 2 * table_level
  uint32_t num_buckets;  // 6 << (2 * table_level)
};

S2B_HD bool cellid_valid(uint64_t id) {   // S2CellId::is_valid, s2cell_id.h:583-585
  return (id >> 61) < 6 && (cellid_lsb(id) & 0x1555555555555555ull) != 0;
}

// Tree order: pair (id) sorts before the key (kmin, kid) when its range_min is smaller, or equal with a larger cell.
S2B_HD bool ci_before(uint64_t id, uint64_t kmin, uint64_t kid) {
  const uint64_t rmin = cellid_range_min(id);
  return rmin < kmin || (rmin == kmin && id > kid);
}

// First position whose pair does not sort before (kmin, kid); kmin is a leaf id or End(30) (the bucket index is then B).
S2B_HD uint32_t ci_lower_bound(const CellIndexView& ix, uint64_t kmin, uint64_t kid) {
  const uint64_t b = kmin >> ix.shift;
  uint32_t lo = ix.first[b], hi = ix.first[b + 1];
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (ci_before(ix.ids[mid], kmin, kid)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// State carried across the cells of one target union.
struct CiUnionState { int32_t cutoff = -1; };

// Reports every pair intersecting target cell `t` that no earlier cell of the same union reported: emit(position).
// Order: proper ancestors deepest first, then the run of t and its descendants in tree order. Returns the number emitted.
template <class Emit>
S2B_HD uint32_t ci_visit_cell(const CellIndexView& ix, uint64_t t, CiUnionState& st, Emit&& emit) {
  if (!cellid_valid(t) || ix.m == 0) return 0;
  const uint64_t tmin = cellid_range_min(t), tmax = cellid_range_max(t);
  const uint32_t lo = ci_lower_bound(ix, tmin, t);
  // the run ends before the first pair whose range_min exceeds tmax; (tmax + 2, ~0) precedes every pair that starts there.
  // A leaf target (a point) has no descendants: its run is the copies of the leaf itself.
  uint32_t hi = lo;
  if (t & 1) { while (hi < ix.m && ix.ids[hi] == t) ++hi; } else { hi = ci_lower_bound(ix, tmax + 2, ~0ull); }
  uint32_t n = 0;
  int32_t a = (int32_t)lo - 1;
  while (a >= 0 && !cellid_contains(ix.ids[a], t)) a = ix.parent[a];
  const int32_t start = a;
  for (; a > st.cutoff; a = ix.parent[a]) { emit((uint32_t)a); ++n; }
  if (start > st.cutoff) st.cutoff = start;
  for (uint32_t e = lo; e < hi; ++e) { emit(e); ++n; }
  return n;
}

// pair cell ∩ target cell (one contains the other): the smaller cell. Its S2CellId::lsb is the measure the sharder sums
// (s2region_sharder.cc:133-136).
S2B_HD uint64_t ci_overlap_cell(uint64_t a, uint64_t b) { return cellid_lsb(a) < cellid_lsb(b) ? a : b; }

// S2CellUnion::Intersects(S2CellId) (s2cell_union.cc:302-309) over cells[0, n): sorted, non-overlapping.
S2B_HD bool ci_union_intersects(const uint64_t* cells, uint32_t n, uint64_t id) {
  const uint64_t lo_id = cellid_range_min(id), hi_id = cellid_range_max(id);
  uint32_t lo = 0, hi = n;
  while (lo < hi) {   // first cell whose range_max reaches id's range_min
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (cellid_range_max(cells[mid]) < lo_id) lo = mid + 1; else hi = mid;
  }
  return lo < n && cellid_range_min(cells[lo]) <= hi_id;
}

// The sharder's choice (s2region_sharder.cc:57-60,112-140) from (label, weight) pairs sorted by label: a single candidate
// label wins outright; otherwise the largest positive total, the smaller label on ties; `fallback` if there is none.
S2B_HD int32_t ci_best_label(const int32_t* sorted, const uint64_t* weight, uint32_t begin, uint32_t end, int32_t fallback) {
  if (begin == end) return fallback;
  if (sorted[begin] == sorted[end - 1]) return sorted[begin];
  int32_t best = fallback;
  long long best_sum = 0;
  uint32_t k = begin;
  while (k < end) {
    const int32_t label = sorted[k];
    long long sum = 0;
    for (; k < end && sorted[k] == label; ++k) sum += (long long)weight[k];
    if (sum > best_sum) { best = label; best_sum = sum; }   // labels ascend, so ties keep the smaller one
  }
  return best;
}

}  // namespace s2b

// ---- host side: the builder (plain C++; also compiled by tests/hostsim)
#include <algorithm>
#include <vector>
namespace s2b {

struct CellIndexHost {
  std::vector<uint64_t> ids;
  std::vector<int32_t> labels, parent;
  std::vector<uint32_t> first;
  int table_level = 0;
  int32_t num_labels = 0;   // max label + 1
  CellIndexView view() const {
    return CellIndexView{ids.data(), labels.data(), parent.data(), first.data(), (uint32_t)ids.size(), 61 - 2 * table_level,
                         (uint32_t)(6u << (2 * table_level))};
  }
};

// S2CellIndex::Add x m + Build (s2cell_index.cc:71-146): sort into tree order, link every pair to its nearest enclosing pair
// with the reference's push/pop stack, and tabulate the buckets. Returns false on an invalid cell id or a negative label.
inline bool build_cell_index_host(const uint64_t* cell_ids, const int32_t* labels, size_t m, CellIndexHost* out) {
  std::vector<uint32_t> order(m);
  for (size_t i = 0; i < m; ++i) {
    if (!cellid_valid(cell_ids[i]) || labels[i] < 0) return false;
    order[i] = (uint32_t)i;
  }
  std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    const uint64_t a = cell_ids[x], b = cell_ids[y];
    const uint64_t ra = cellid_range_min(a), rb = cellid_range_min(b);
    if (ra != rb) return ra < rb;
    if (a != b) return a > b;
    if (labels[x] != labels[y]) return labels[x] < labels[y];
    return x < y;
  });
  out->ids.resize(m); out->labels.resize(m); out->parent.resize(m);
  int32_t top = -1, max_label = -1;
  for (size_t e = 0; e < m; ++e) {
    const uint64_t id = cell_ids[order[e]];
    out->ids[e] = id;
    out->labels[e] = labels[order[e]];
    if (out->labels[e] > max_label) max_label = out->labels[e];
    while (top >= 0 && !cellid_contains(out->ids[top], id)) top = out->parent[top];   // pop the cells that ended before id
    out->parent[e] = top;
    top = (int32_t)e;
  }
  out->num_labels = max_label + 1;
  int level = 0;   // about two buckets per pair, between level 0 (6 buckets) and level 12 (100M buckets)
  while (level < 12 && (6ull << (2 * (level + 1))) <= 2 * (uint64_t)m) ++level;
  out->table_level = level;
  const int shift = 61 - 2 * level;
  const uint32_t B = 6u << (2 * level);
  out->first.assign((size_t)B + 2, (uint32_t)m);
  uint64_t next = 0;   // buckets below `next` are final
  for (size_t e = 0; e < m; ++e) {
    const uint64_t b = cellid_range_min(out->ids[e]) >> shift;
    for (; next <= b; ++next) out->first[next] = (uint32_t)e;
  }
  return true;
}

}  // namespace s2b
