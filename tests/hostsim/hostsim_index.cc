// tests/hostsim/hostsim_index.cc — TEST-ONLY host execution of the flat-index query logic (s2b_contains.cuh) so the
// Locate / lookup-table / crossing-parity code is checked against the oracle without a GPU.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include "s2b_flat_host.h"

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

}  // extern "C"
