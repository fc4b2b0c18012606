// s2b_index_build.cc — C ABI of the host-side index builder (s2b_index_builder.h). Plain C++ (compiled as host code only).
#include <new>
#include <string>

#include "../../include/s2b_capi.h"
#include "s2b_index_builder.h"
#include "s2b_index_wire.h"
#include "s2b_shapes_wire.h"

namespace s2b { int set_error(int code, const char* fmt, ...); }

struct S2bBuiltIndex { s2b::BuiltIndex data; };
struct S2bShapeSoup { s2b::ShapeSoupData data; };

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

int s2b_index_decode(const void* bytes, size_t len, int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                     const uint32_t* chain_vertex_begin, const double* vertices, S2bBuiltIndex** out, int* max_edges_per_cell_out,
                     size_t* consumed_out) {
  if (!out) return s2b::set_error(S2B_EINVAL, "null output");
  *out = nullptr;
  S2bBuiltIndex* b = new (std::nothrow) S2bBuiltIndex();
  if (!b) return s2b::set_error(S2B_ENOMEM, "host allocation failed");
  std::string err;
  int rc;
  try {
    rc = s2b::decode_index_wire((const uint8_t*)bytes, len, num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices,
                                &b->data, max_edges_per_cell_out, consumed_out, &err);
  } catch (const std::bad_alloc&) {
    delete b;
    return s2b::set_error(S2B_ENOMEM, "host allocation failed while decoding the index");
  }
  if (rc != S2B_OK) { delete b; return s2b::set_error(rc, "%s", err.c_str()); }
  *out = b;
  return S2B_OK;
}

int s2b_index_encode(const S2bFlatIndex* flat, const int32_t* edge_id, int max_edges_per_cell, void* out, size_t capacity,
                     size_t* size_out) {
  if (!flat || !size_out) return s2b::set_error(S2B_EINVAL, "null argument");
  std::vector<uint8_t> bytes;
  std::string err;
  int rc;
  try {
    rc = s2b::encode_index_wire(*flat, edge_id, max_edges_per_cell, &bytes, &err);
  } catch (const std::bad_alloc&) {
    return s2b::set_error(S2B_ENOMEM, "host allocation failed while encoding the index");
  }
  if (rc != S2B_OK) return s2b::set_error(rc, "%s", err.c_str());
  *size_out = bytes.size();
  if (!out || capacity < bytes.size()) return s2b::set_error(S2B_ECAPACITY, "index encode: %zu bytes needed", bytes.size());
  if (!bytes.empty()) std::memcpy(out, bytes.data(), bytes.size());
  return S2B_OK;
}

int s2b_shapes_decode(const void* bytes, size_t len, S2bShapeSoup** out, size_t* consumed_out) {
  if (!out) return s2b::set_error(S2B_EINVAL, "null output");
  *out = nullptr;
  S2bShapeSoup* soup = new (std::nothrow) S2bShapeSoup();
  if (!soup) return s2b::set_error(S2B_ENOMEM, "host allocation failed");
  std::string err;
  int rc;
  try {
    rc = s2b::decode_tagged_shapes((const uint8_t*)bytes, len, &soup->data, consumed_out, &err);
  } catch (const std::bad_alloc&) {
    delete soup;
    return s2b::set_error(S2B_ENOMEM, "host allocation failed while decoding shapes");
  }
  if (rc != S2B_OK) { delete soup; return s2b::set_error(rc, "%s", err.c_str()); }
  *out = soup;
  return S2B_OK;
}

int s2b_shape_soup_arrays(const S2bShapeSoup* soup, int* num_shapes, const int8_t** shape_dim, const uint32_t** shape_chain_begin,
                          const uint32_t** chain_vertex_begin, const double** vertices, size_t* num_vertices) {
  if (!soup) return s2b::set_error(S2B_EINVAL, "null soup");
  const s2b::ShapeSoupData& d = soup->data;
  if (num_shapes) *num_shapes = (int)d.shape_dim.size();
  if (shape_dim) *shape_dim = d.shape_dim.data();
  if (shape_chain_begin) *shape_chain_begin = d.shape_chain_begin.data();
  if (chain_vertex_begin) *chain_vertex_begin = d.chain_vertex_begin.data();
  if (vertices) *vertices = d.vertices.data();
  if (num_vertices) *num_vertices = d.vertices.size() / 3;
  return S2B_OK;
}

void s2b_shape_soup_destroy(S2bShapeSoup* soup) { delete soup; }

int s2b_shapes_encode_hint(int coding_hint, int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                           const uint32_t* chain_vertex_begin, const double* vertices, void* out, size_t capacity, size_t* size_out) {
  if (coding_hint != S2B_CODING_FAST && coding_hint != S2B_CODING_COMPACT) return s2b::set_error(S2B_EINVAL, "coding_hint must be S2B_CODING_FAST or S2B_CODING_COMPACT");
  if (!size_out || num_shapes < 0 || (num_shapes && (!shape_dim || !shape_chain_begin || !chain_vertex_begin))) return s2b::set_error(S2B_EINVAL, "null argument");
  std::vector<uint8_t> bytes;
  std::string err;
  int rc;
  try {
    rc = s2b::encode_tagged_shapes(num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices, &bytes, &err, coding_hint == S2B_CODING_COMPACT);
  } catch (const std::bad_alloc&) {
    return s2b::set_error(S2B_ENOMEM, "host allocation failed while encoding shapes");
  }
  if (rc != S2B_OK) return s2b::set_error(rc, "%s", err.c_str());
  *size_out = bytes.size();
  if (!out || capacity < bytes.size()) return s2b::set_error(S2B_ECAPACITY, "shape encode: %zu bytes needed", bytes.size());
  if (!bytes.empty()) std::memcpy(out, bytes.data(), bytes.size());
  return S2B_OK;
}

int s2b_shapes_encode(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin, const uint32_t* chain_vertex_begin,
                      const double* vertices, void* out, size_t capacity, size_t* size_out) {
  return s2b_shapes_encode_hint(S2B_CODING_FAST, num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices, out, capacity, size_out);
}

}  // extern "C"
