// tests/hostsim/hostsim_cellindex.cc — TEST-ONLY host compilation of s2b_cellindex.cuh (builder + traversal), so the
// S2CellIndex logic the kernels inline is checked against the reference on the CPU. Nothing in the product loads this.
#include <algorithm>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

#include "s2b_cellindex.cuh"

using namespace s2b;

extern "C" {

// VisitIntersectingCells for T targets (CSR, offsets nullable = single cells); each target's pairs sorted by (cell, label).
// Returns the total; writes at most `capacity` pairs. Returns ~0 if the builder rejects the input.
uint64_t hs_cell_index_visit(const uint64_t* ids, const int32_t* labels, size_t m, const uint64_t* target_ids, const uint32_t* target_offsets,
                             size_t T, uint64_t* out_offsets, uint64_t* out_cells, int32_t* out_labels, uint64_t capacity) {
  CellIndexHost h;
  if (!build_cell_index_host(ids, labels, m, &h)) return ~0ull;
  const CellIndexView ix = h.view();
  uint64_t total = 0;
  out_offsets[0] = 0;
  std::vector<std::pair<uint64_t, int32_t>> got;
  for (size_t t = 0; t < T; ++t) {
    const uint32_t b = target_offsets ? target_offsets[t] : (uint32_t)t, e = target_offsets ? target_offsets[t + 1] : (uint32_t)t + 1;
    CiUnionState st;
    got.clear();
    for (uint32_t k = b; k < e; ++k) ci_visit_cell(ix, target_ids[k], st, [&](const CiNode& nd) { got.push_back({nd.id, nd.label}); });
    std::sort(got.begin(), got.end());
    for (auto& g : got) {
      if (total < capacity) { out_cells[total] = g.first; out_labels[total] = g.second; }
      ++total;
    }
    out_offsets[t + 1] = total;
  }
  return total;
}

// The sharder's choice per region (bound covering + optional region union as the MayIntersect filter), through the same
// device functions as the kernels: pairs -> (label, weight) sorted by label -> ci_best_label.
void hs_cell_index_best_label(const uint64_t* ids, const int32_t* labels, size_t m, const uint64_t* bound_ids, const uint32_t* bound_offsets,
                              const uint64_t* region_ids, const uint32_t* region_offsets, size_t T, int32_t fallback, int32_t* out) {
  CellIndexHost h;
  if (!build_cell_index_host(ids, labels, m, &h)) return;
  const CellIndexView ix = h.view();
  for (size_t t = 0; t < T; ++t) {
    CiUnionState st;
    std::vector<std::pair<int32_t, uint64_t>> got;
    for (uint32_t k = bound_offsets[t]; k < bound_offsets[t + 1]; ++k) {
      const uint64_t target = bound_ids[k];
      st.cutoff = -1;   // every (shard cell, bound cell) overlap counts
      ci_visit_cell(ix, target, st, [&](const CiNode& nd) {
        const uint64_t c = ci_overlap_cell(nd.id, target);
        const bool hit = !region_ids || ci_union_intersects(region_ids + region_offsets[t], region_offsets[t + 1] - region_offsets[t], c);
        got.push_back({nd.label, hit ? cellid_lsb(c) : 0});
      });
    }
    std::stable_sort(got.begin(), got.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<int32_t> sl;
    std::vector<uint64_t> sw;
    for (auto& g : got) { sl.push_back(g.first); sw.push_back(g.second); }
    out[t] = ci_best_label(sl.data(), sw.data(), 0, (uint32_t)sl.size(), fallback);
  }
}

// The point path: labels of the pairs containing each leaf (range table + flat lists or chain walk), sorted per leaf.
// force_chain != 0 drops the flat lists. Returns the total; writes at most `capacity` labels.
uint64_t hs_cell_index_leaf_labels(const uint64_t* ids, const int32_t* labels, size_t m, const uint64_t* leaves, size_t T, int force_chain,
                                   uint64_t* out_offsets, int32_t* out_labels, uint64_t capacity) {
  CellIndexHost h;
  if (!build_cell_index_host(ids, labels, m, &h)) return ~0ull;
  if (force_chain) h.make_records(false);
  const CellIndexView ix = h.view();
  uint64_t total = 0;
  out_offsets[0] = 0;
  std::vector<int32_t> got;
  for (size_t t = 0; t < T; ++t) {
    got.clear();
    ci_visit_leaf_labels(ix, leaves[t], [&](int32_t l) { got.push_back(l); });
    std::sort(got.begin(), got.end());
    for (int32_t l : got) { if (total < capacity) out_labels[total] = l; ++total; }
    out_offsets[t + 1] = total;
  }
  return total;
}

}  // extern "C"
