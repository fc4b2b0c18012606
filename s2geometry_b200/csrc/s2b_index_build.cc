// s2b_index_build.cc — C ABI of the host-side index builder (s2b_index_builder.h). Plain C++ (compiled as host code only).
#include <new>
#include <string>

#include "../../include/s2b_capi.h"
#include "s2b_index_builder.h"

namespace s2b { int set_error(int code, const char* fmt, ...); }

struct S2bBuiltIndex { s2b::BuiltIndex data; };

extern "C" {

int s2b_index_build(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                    const uint32_t* chain_vertex_begin, const double* vertices, int max_edges_per_cell, S2bBuiltIndex** out) {
  if (!out) return s2b::set_error(S2B_EINVAL, "null output");
  *out = nullptr;
  S2bBuiltIndex* b = new (std::nothrow) S2bBuiltIndex();
  if (!b) return s2b::set_error(S2B_ENOMEM, "host allocation failed");
  std::string err;
  int rc;
  try {
    rc = s2b::build_index_from_shapes(num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices, max_edges_per_cell, &b->data, &err);
  } catch (const std::bad_alloc&) {
    delete b;
    return s2b::set_error(S2B_ENOMEM, "host allocation failed while building the index");
  }
  if (rc != S2B_OK) { delete b; return s2b::set_error(rc, "%s", err.c_str()); }
  *out = b;
  return S2B_OK;
}

int s2b_built_index_flat(const S2bBuiltIndex* built, S2bFlatIndex* flat_out, const int32_t** edge_id_out) {
  if (!built || !flat_out) return s2b::set_error(S2B_EINVAL, "null argument");
  *flat_out = built->data.as_flat();
  if (edge_id_out) *edge_id_out = built->data.edge_id.data();
  return S2B_OK;
}

void s2b_built_index_destroy(S2bBuiltIndex* built) { delete built; }

}  // extern "C"
