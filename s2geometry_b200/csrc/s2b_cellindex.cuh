// s2b_cellindex.cuh — flat S2CellIndex (s2cell_index.h:113-441, s2cell_index.cc:71-146) and the traversal every batched
// query over it shares. Integer work only, bit-exact with the reference.
//
// The reference keeps the (cell_id, label) pairs as a tree (`cell_tree_`, parent = the nearest enclosing pair) in the order
// its Build() pushes them: by range_min, larger cells first, then by label (the `Delta` order, s2cell_index.cc:79-99) — a
// preorder of the containment forest — plus a list of leaf-cell ranges pointing into the tree. Here the tree is three
// parallel arrays in that same order and the range list is replaced by a bucket table over the Hilbert curve:
//
//   nodes[m]            {cell id, label, parent} in tree order, 16 bytes = one load per step of a walk;
//                       parent = nearest enclosing pair (an equal cell with a smaller label counts), -1 for roots
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

struct alignas(16) CiNode {
  uint64_t id;
  int32_t label;
  int32_t parent;
};

S2B_HD CiNode ci_load(const CiNode* p) {
#if defined(__CUDA_ARCH__)
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));   // read-only path, one 128-bit load
  return CiNode{((uint64_t)v.y << 32) | v.x, (int32_t)v.z, (int32_t)v.w};
#else
  return *p;
#endif
}

struct CellIndexView {
  const CiNode* nodes;
  const uint32_t* first;
  uint32_t m;
  int shift;             // 61 - 2 * table_level
  uint32_t num_buckets;  // 6 << (2 * table_level)
  // Leaf-cell ranges, the reference's `range_nodes_` (s2cell_index.h:419-441): range r = leaves [rstart[r], rstart[r+1]);
  // rfirst = bucket table over rstart (rshift like shift). rrec[r] describes the pairs covering the range (see CiRangeRec).
  const uint64_t* rstart;
  const uint32_t* rfirst;
  const struct CiRangeRec* rrec;
  const int32_t* llabels;
  uint32_t num_ranges;
  int rshift;
};

// What covers a leaf range, in one 32-byte sector: the number of covering pairs and the labels of the first six inline
// (deepest pair first), the rest at llabels[off ...]. n == kCiChain: the lists of this index would be too large, walk the
// chain of tree nodes that starts at node `off` instead (the reference's ContentsIterator, s2cell_index.h:589-598).
constexpr uint32_t kCiChain = 0xFFFFFFFFu;
constexpr int kCiInline = 6;
struct alignas(32) CiRangeRec {
  uint32_t off;
  uint32_t n;
  int32_t label[kCiInline];
};

S2B_HD CiRangeRec ci_load_rec(const CiRangeRec* p) {
#if defined(__CUDA_ARCH__)
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  return CiRangeRec{a.x, a.y, {(int32_t)a.z, (int32_t)a.w, (int32_t)b.x, (int32_t)b.y, (int32_t)b.z, (int32_t)b.w}};
#else
  return *p;
#endif
}


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
    if (ci_before(ix.nodes[mid].id, kmin, kid)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// State carried across the cells of one target union.
struct CiUnionState { int32_t cutoff = -1; };

// Reports every pair intersecting target cell `t` that no earlier cell of the same union reported: emit(node).
// Order: proper ancestors deepest first, then the run of t and its descendants in tree order. Returns the number emitted.
template <class Emit>
S2B_HD uint32_t ci_visit_cell(const CellIndexView& ix, uint64_t t, CiUnionState& st, Emit&& emit) {
  if (!cellid_valid(t) || ix.m == 0) return 0;
  const uint64_t tmin = cellid_range_min(t), tmax = cellid_range_max(t);
  const uint32_t lo = ci_lower_bound(ix, tmin, t);
  // the run ends before the first pair whose range_min exceeds tmax; (tmax + 2, ~0) precedes every pair that starts there.
  // A leaf target (a point) has no descendants: its run is the copies of the leaf itself.
  uint32_t hi = lo;
  if (t & 1) { while (hi < ix.m && ix.nodes[hi].id == t) ++hi; } else { hi = ci_lower_bound(ix, tmax + 2, ~0ull); }
  uint32_t n = 0;
  int32_t a = (int32_t)lo - 1;
  CiNode node{0, 0, -1};
  while (a >= 0) {   // up from the pair before the run until one contains t
    node = ci_load(ix.nodes + a);
    if (cellid_contains(node.id, t)) break;
    a = node.parent;
  }
  const int32_t start = a;
  while (a > st.cutoff) {
    emit(node); ++n;
    a = node.parent;
    if (a > st.cutoff) node = ci_load(ix.nodes + a);
  }
  if (start > st.cutoff) st.cutoff = start;
  for (uint32_t e = lo; e < hi; ++e) { emit(ci_load(ix.nodes + e)); ++n; }
  return n;
}

// The leaf range containing `leaf` (RangeIterator::Seek, s2cell_index.cc:32-36): upper_bound over rstart, minus one.
S2B_HD uint32_t ci_seek_range(const CellIndexView& ix, uint64_t leaf) {
  const uint64_t b = leaf >> ix.rshift;
  uint32_t lo = ix.rfirst[b], hi = ix.rfirst[b + 1];
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (ix.rstart[mid] <= leaf) lo = mid + 1; else hi = mid;
  }
  return lo - 1;   // the first range starts at the first leaf of face 0, so lo >= 1 for every valid leaf
}

// Reports the labels a range record stands for.
template <class Emit>
S2B_HD void ci_emit_rec_labels(const CellIndexView& ix, const CiRangeRec& rec, Emit&& emit) {
  if (rec.n == kCiChain) {
    for (int32_t a = (int32_t)rec.off; a >= 0;) {
      const CiNode nd = ci_load(ix.nodes + a);
      emit(nd.label);
      a = nd.parent;
    }
    return;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < kCiInline; ++k) if ((uint32_t)k < rec.n) emit(rec.label[k]);
  for (uint32_t k = kCiInline; k < rec.n; ++k) emit(ix.llabels[rec.off + (k - kCiInline)]);
}

// The labels of the pairs whose cell contains a valid leaf cell (what VisitIntersectingCells reports for a point).
template <class Emit>
S2B_HD void ci_visit_leaf_labels(const CellIndexView& ix, uint64_t leaf, Emit&& emit) {
  ci_emit_rec_labels(ix, ci_load_rec(ix.rrec + ci_seek_range(ix, leaf)), emit);
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
  std::vector<CiNode> nodes;
  std::vector<uint32_t> first;
  int table_level = 0;
  int32_t num_labels = 0;   // max label + 1
  std::vector<uint64_t> rstart;
  std::vector<int32_t> rcontents;   // deepest pair covering each range (-1: none); host only, the records derive from it
  std::vector<uint32_t> rfirst;
  std::vector<CiRangeRec> rrec;
  std::vector<int32_t> llabels;
  int range_table_level = 0;
  bool has_lists = false;
  CellIndexView view() const {
    return CellIndexView{nodes.data(), first.data(), (uint32_t)nodes.size(), 61 - 2 * table_level,
                         (uint32_t)(6u << (2 * table_level)), rstart.data(), rfirst.data(), rrec.data(), llabels.data(),
                         (uint32_t)rstart.size(), 61 - 2 * range_table_level};
  }
  // Range records from rcontents; with_lists = false forces the chain-walk form (tests).
  void make_records(bool with_lists) {
    const size_t R = rstart.size(), m = nodes.size();
    std::vector<uint32_t> depth(m);
    for (size_t e = 0; e < m; ++e) depth[e] = nodes[e].parent < 0 ? 1 : depth[nodes[e].parent] + 1;
    uint64_t overflow = 0;
    for (size_t r = 0; r < R; ++r) {
      const uint32_t d = rcontents[r] < 0 ? 0 : depth[rcontents[r]];
      if (d > (uint32_t)kCiInline) overflow += d - kCiInline;
    }
    has_lists = with_lists && overflow <= 64 * (uint64_t)m + 1024 && overflow < 0xFFFFFFF0ull;
    rrec.assign(R, CiRangeRec{0, 0, {0, 0, 0, 0, 0, 0}});
    llabels.clear();
    if (has_lists) llabels.reserve((size_t)overflow);
    for (size_t r = 0; r < R; ++r) {
      CiRangeRec& rec = rrec[r];
      const uint32_t d = rcontents[r] < 0 ? 0 : depth[rcontents[r]];
      if (!has_lists && d > 0) { rec.off = (uint32_t)rcontents[r]; rec.n = kCiChain; continue; }
      rec.off = (uint32_t)llabels.size();
      rec.n = d;
      uint32_t k = 0;
      for (int32_t a = rcontents[r]; a >= 0; a = nodes[a].parent, ++k) {
        if (k < (uint32_t)kCiInline) rec.label[k] = nodes[a].label; else llabels.push_back(nodes[a].label);
      }
    }
  }
};


// Bucket table over ascending leaf-aligned keys: first[b] = number of keys before bucket b; level = about `per_key` buckets per key.
inline int ci_bucket_table(const std::vector<uint64_t>& keys, double per_key, int max_level, std::vector<uint32_t>* first) {
  int level = 0;
  while (level < max_level && (double)(6ull << (2 * (level + 1))) <= per_key * (double)keys.size()) ++level;
  const int shift = 61 - 2 * level;
  const uint32_t B = 6u << (2 * level);
  first->assign((size_t)B + 2, (uint32_t)keys.size());
  uint64_t next = 0;   // buckets below `next` are final
  for (size_t e = 0; e < keys.size(); ++e) {
    const uint64_t b = keys[e] >> shift;
    for (; next <= b && next < (uint64_t)B + 2; ++next) (*first)[next] = (uint32_t)e;
  }
  return level;
}

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
  out->nodes.resize(m);
  int32_t top = -1, max_label = -1;
  for (size_t e = 0; e < m; ++e) {
    const uint64_t id = cell_ids[order[e]];
    const int32_t label = labels[order[e]];
    if (label > max_label) max_label = label;
    while (top >= 0 && !cellid_contains(out->nodes[top].id, id)) top = out->nodes[top].parent;   // pop the cells that ended before id
    out->nodes[e] = CiNode{id, label, top};
    top = (int32_t)e;
  }
  out->num_labels = max_label + 1;
  std::vector<uint64_t> keys(m);
  for (size_t e = 0; e < m; ++e) keys[e] = cellid_range_min(out->nodes[e].id);
  out->table_level = ci_bucket_table(keys, 2.0, 12, &out->first);
  // Leaf ranges (s2cell_index.cc:101-145): a range starts wherever a cell starts or ends, plus the two ends of the curve.
  std::vector<uint64_t>& rs = out->rstart;
  rs.clear();
  rs.reserve(2 * m + 2);
  rs.push_back(1);                           // S2CellId::Begin(kMaxLevel)
  rs.push_back(0xC000000000000001ull);       // S2CellId::End(kMaxLevel)
  for (size_t e = 0; e < m; ++e) { rs.push_back(keys[e]); rs.push_back(cellid_range_max(out->nodes[e].id) + 2); }
  std::sort(rs.begin(), rs.end());
  rs.erase(std::unique(rs.begin(), rs.end()), rs.end());
  const size_t R = rs.size();
  out->rcontents.assign(R, -1);
  {
    size_t e = 0;
    top = -1;
    for (size_t r = 0; r < R; ++r) {
      const uint64_t s = rs[r];
      while (top >= 0 && cellid_range_max(out->nodes[top].id) < s) top = out->nodes[top].parent;   // cells that ended
      for (; e < m && keys[e] == s; ++e) top = (int32_t)e;                                          // cells that start here
      out->rcontents[r] = top;
    }
  }
  // fine buckets: most hold no range start, so no probes. Level <= 10: a bucket is named by the first three 4-bit steps of
  // FromFaceIJ (24 position bits), which is all the point kernel computes before it looks the bucket up.
  out->range_table_level = ci_bucket_table(rs, 8.0, 10, &out->rfirst);
  out->make_records(true);
  return true;
}

}  // namespace s2b
