// s2b_index_wire.h — reader and writer for the reference's serialized shape index (host C++, no CUDA).
//
// Format (all from the reference, restated from its documentation and decoders; nothing is linked):
//   MutableS2ShapeIndex::Encode / Init        mutable_s2shape_index.cc:1988-2040
//     varint64  max_edges_per_cell << 2 | version(0)
//     EncodedS2CellIdVector  cell ids          encoded_s2cell_id_vector.cc:30-140
//     EncodedStringVector    one string per cell (encoded_string_vector.cc:25-60), each an
//     S2ShapeIndexCell::Encode record          s2shape_index.cc:69-320
//   EncodedUintVector<uint64>                  encoded_uint_vector.h:130-300
//   varints                                    util/coding/varint.h:262-284
// The encoding stores only the cells (ids, clipped shapes, edge ids); geometry travels separately
// (s2shapeutil::EncodeTaggedShapes in the reference). Here the geometry is the same "shape soup" the builder takes,
// and edge ids resolve to endpoints with S2LaxPolygonShape / S2LaxPolylineShape / S2PointVectorShape numbering.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "s2b_index_builder.h"

namespace s2b {
namespace wire {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  size_t avail() const { return (size_t)(end - p); }
  bool get8(uint32_t* v) { if (p >= end) return false; *v = *p++; return true; }
  bool varint64(uint64_t* out) {
    uint64_t r = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p >= end) return false;
      const uint64_t b = *p++;
      if (shift == 63 && b > 1) return false;
      r |= (b & 127) << shift;
      if (b < 128) { *out = r; return true; }
    }
    return false;
  }
  bool varint32(uint32_t* out) {   // at most 5 bytes, the last one < 16 (varint.h:262-284)
    uint32_t r = 0;
    for (int k = 0; k < 5; ++k) {
      if (p >= end) return false;
      const uint32_t b = *p++;
      r |= (b & 127) << (7 * k);
      if (k == 4 ? b < 16 : b < 128) { *out = r; return true; }
      if (k == 4) return false;
    }
    return false;
  }
  bool uint_with_length(int length, uint64_t* out) {   // little endian, `length` bytes
    if (avail() < (size_t)length) return false;
    uint64_t r = 0;
    for (int k = 0; k < length; ++k) r |= (uint64_t)p[k] << (8 * k);
    p += length;
    *out = r;
    return true;
  }
};

struct UintVector {   // EncodedUintVector<uint64_t>
  const uint8_t* data = nullptr;
  uint64_t size = 0;
  int len = 1;
  bool init(Reader* r) {
    uint64_t size_len;
    if (!r->varint64(&size_len)) return false;
    size = size_len / 8;
    len = (int)(size_len & 7) + 1;
    if (size > 0xFFFFFFFFull) return false;   // the reference keeps the size in 32 bits
    const uint64_t bytes = size * (uint64_t)len;
    if (r->avail() < bytes) return false;
    data = r->p;
    r->p += bytes;
    return true;
  }
  uint64_t operator[](uint64_t i) const {
    uint64_t v = 0;
    const uint8_t* q = data + i * len;
    for (int k = 0; k < len; ++k) v |= (uint64_t)q[k] << (8 * k);
    return v;
  }
};

struct Writer {
  std::vector<uint8_t>* out;
  void put8(uint32_t v) { out->push_back((uint8_t)v); }
  void varint64(uint64_t v) {
    while (v >= 128) { out->push_back((uint8_t)(v | 128)); v >>= 7; }
    out->push_back((uint8_t)v);
  }
  void varint32(uint32_t v) { varint64(v); }
  void uint_with_length(uint64_t v, int length) { for (int k = 0; k < length; ++k) { out->push_back((uint8_t)v); v >>= 8; } }
  void uint_vector(const std::vector<uint64_t>& v) {   // EncodeUintVector<uint64_t>
    uint64_t one_bits = 1;
    for (uint64_t x : v) one_bits |= x;
    const int len = ((63 - __builtin_clzll(one_bits)) >> 3) + 1;
    varint64(((uint64_t)v.size() * 8) | (uint64_t)(len - 1));
    for (uint64_t x : v) uint_with_length(x, len);
  }
};

// Shape soup -> (shape, edge id) -> endpoints. Edge numbering: chains in order; a dimension-2 chain of n vertices has
// n edges (closed), a dimension-1 chain n-1, a dimension-0 chain n degenerate edges (s2lax_polygon_shape.cc:255,
// s2lax_polyline_shape.h, s2point_vector_shape.h).
struct SoupEdges {
  int num_shapes;
  const int8_t* dim;
  const uint32_t* shape_chain_begin;
  const uint32_t* chain_vertex_begin;
  const double* vertices;
  std::vector<uint64_t> chain_edge_begin;   // per chain, shape-local; one extra entry per shape is not needed:
  std::vector<uint64_t> shape_num_edges;
  void init() {
    const uint32_t num_chains = num_shapes ? shape_chain_begin[num_shapes] : 0;
    chain_edge_begin.assign(num_chains, 0);
    shape_num_edges.assign(num_shapes, 0);
    for (int s = 0; s < num_shapes; ++s) {
      uint64_t e = 0;
      for (uint32_t c = shape_chain_begin[s]; c < shape_chain_begin[s + 1]; ++c) {
        chain_edge_begin[c] = e;
        const uint32_t n = chain_vertex_begin[c + 1] - chain_vertex_begin[c];
        e += dim[s] == 2 ? n : (dim[s] == 1 ? (n ? n - 1 : 0) : n);
      }
      shape_num_edges[s] = e;
    }
  }
  void edge(int s, uint64_t e, double* v0, double* v1) const {
    uint32_t lo = shape_chain_begin[s], hi = shape_chain_begin[s + 1] - 1;   // last chain with chain_edge_begin <= e
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (chain_edge_begin[mid] <= e) lo = mid; else hi = mid - 1;
    }
    const uint32_t b = chain_vertex_begin[lo], n = chain_vertex_begin[lo + 1] - b;
    const uint32_t k = (uint32_t)(e - chain_edge_begin[lo]);
    const uint32_t k1 = dim[s] == 0 ? k : (k + 1 == n ? 0 : k + 1);
    std::memcpy(v0, vertices + 3 * (size_t)(b + k), 24);
    std::memcpy(v1, vertices + 3 * (size_t)(b + k1), 24);
  }
};

// `limit` = number of edges of the shape: counts and ids beyond it are rejected BEFORE anything is materialised, so a
// malformed record cannot force large allocations.
inline bool decode_edges(Reader* r, uint32_t num_edges, uint64_t limit, std::vector<int32_t>* ids) {   // s2shape_index.cc:325-375
  if (num_edges > limit) return false;
  int64_t edge_id = 0;
  for (uint32_t i = 0; i < num_edges;) {
    uint32_t delta = 0;
    if (!r->varint32(&delta)) return false;
    if (i + 1 == num_edges) {
      edge_id += delta;
      if (edge_id > INT32_MAX || (uint64_t)edge_id >= limit) return false;
      ids->push_back((int32_t)edge_id);
      ++i;
    } else {
      uint32_t count = (delta & 7) + 1;
      delta >>= 3;
      if (count == 8) {
        count = delta + 8;
        if (!r->varint32(&delta)) return false;
      }
      if ((uint64_t)i + count > num_edges) return false;
      edge_id += delta;
      if (edge_id + count > INT32_MAX || (uint64_t)edge_id + count > limit) return false;
      for (uint32_t k = 0; k < count; ++k) ids->push_back((int32_t)(edge_id + k));
      i += count;
      edge_id += count;
    }
  }
  return true;
}

}  // namespace wire

// Decodes MutableS2ShapeIndex::Encode output into the flat arrays. Returns S2B_OK or S2B_EINVAL (message in *err).
inline int decode_index_wire(const uint8_t* bytes, size_t len, int num_shapes, const int8_t* shape_dim,
                             const uint32_t* shape_chain_begin, const uint32_t* chain_vertex_begin, const double* vertices,
                             BuiltIndex* out, int* max_edges_per_cell, size_t* consumed, std::string* err) {
  using namespace wire;
  auto fail = [&](const std::string& m) { *err = "index decode: " + m; return S2B_EINVAL; };
  if (!bytes && len) return fail("null buffer");
  if (num_shapes < 0 || (num_shapes && (!shape_dim || !shape_chain_begin || !chain_vertex_begin))) return fail("null shape arrays");
  for (int s = 0; s < num_shapes; ++s)
    if (shape_dim[s] < 0 || shape_dim[s] > 2) return fail("shape dimension must be 0, 1 or 2");
  *out = BuiltIndex();
  out->num_shape_ids = num_shapes;
  Reader r{bytes, bytes + len};
  uint64_t max_edges_version;
  if (!r.varint64(&max_edges_version)) return fail("truncated header");
  if ((max_edges_version & 3) != 0) return fail("unsupported encoding version " + std::to_string(max_edges_version & 3));
  if (max_edges_per_cell) *max_edges_per_cell = (int)(max_edges_version >> 2);

  // EncodedS2CellIdVector::Init (encoded_s2cell_id_vector.cc:100-125)
  if (r.avail() < 2) return fail("truncated cell id vector");
  uint32_t code_plus_len = 0, extra = 0;
  r.get8(&code_plus_len);
  int shift_code = (int)(code_plus_len >> 3);
  if (shift_code == 31) {
    r.get8(&extra);
    shift_code = 29 + (int)extra;
    if (shift_code > 56) return fail("bad cell id shift");
  }
  const int base_len = (int)(code_plus_len & 7);
  uint64_t base = 0;
  if (!r.uint_with_length(base_len, &base)) return fail("truncated cell id base");
  base <<= 64 - 8 * (base_len > 1 ? base_len : 1);
  int shift;
  if (shift_code >= 29) {
    shift = 2 * (shift_code - 29) + 1;
    base |= 1ull << (shift - 1);
  } else {
    shift = 2 * shift_code;
  }
  UintVector deltas, offsets;
  if (!deltas.init(&r)) return fail("truncated cell id deltas");
  // EncodedStringVector::Init
  if (!offsets.init(&r)) return fail("truncated cell offsets");
  if (offsets.size != deltas.size) return fail("cell id / cell record count mismatch");
  const uint64_t C = deltas.size;
  const uint64_t data_len = C ? offsets[C - 1] : 0;
  if (r.avail() < data_len) return fail("truncated cell records");
  const uint8_t* data = r.p;
  r.p += data_len;
  if (consumed) *consumed = (size_t)(r.p - bytes);

  SoupEdges soup{num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices};
  soup.init();
  out->cell_ids.resize(C);
  out->cell_center.resize(3 * C);
  out->cell_clip_begin.assign(1, 0);
  out->clip_edge_begin.assign(1, 0);
  uint64_t start = 0;
  for (uint64_t c = 0; c < C; ++c) {
    const uint64_t id = (deltas[c] << shift) + base;
    if (id == 0 || (id >> 61) > 5 || !(cellid_lsb(id) & 0x1555555555555555ull)) return fail("invalid cell id at " + std::to_string(c));
    if (c && cellid_range_max(out->cell_ids[c - 1]) >= cellid_range_min(id)) return fail("cell ids not ascending and disjoint at " + std::to_string(c));
    out->cell_ids[c] = id;
    const P3 ctr = cellid_to_point(id, kHilbert.ij);
    out->cell_center[3 * c] = ctr.x; out->cell_center[3 * c + 1] = ctr.y; out->cell_center[3 * c + 2] = ctr.z;
    const uint64_t limit = offsets[c];
    if (limit < start || limit > data_len) return fail("bad cell record offsets");
    Reader cr{data + start, data + limit};
    start = limit;
    // S2ShapeIndexCell::Decode (s2shape_index.cc:197-290)
    auto add_clip = [&](int64_t shape_id, bool contains_center, const std::vector<int32_t>& ids) -> bool {
      if (shape_id >= num_shapes) return false;
      for (int32_t e : ids) if ((uint64_t)e >= soup.shape_num_edges[shape_id]) return false;
      out->clip_shape_id.push_back((int32_t)shape_id);
      out->clip_flags.push_back((uint8_t)((contains_center ? 1 : 0) | (shape_dim[shape_id] << 1)));
      for (int32_t e : ids) {
        double v0[3], v1[3];
        soup.edge((int)shape_id, (uint64_t)e, v0, v1);
        out->edge_id.push_back(e);
        out->edge_v0.insert(out->edge_v0.end(), v0, v0 + 3);
        out->edge_v1.insert(out->edge_v1.end(), v1, v1 + 3);
      }
      out->clip_edge_begin.push_back((uint32_t)out->edge_id.size());
      return true;
    };
    std::vector<int32_t> ids;
    const std::string where = " in cell " + std::to_string(c);
    if (num_shapes == 1) {
      uint64_t header = 0;
      if (!cr.varint64(&header)) return fail("truncated record" + where);
      bool cc;
      if ((header & 1) == 0) {
        const int n = (int)((header >> 2) & 15) + 2;
        cc = (header & 2) != 0;
        const uint64_t e0 = header >> 6;
        if (e0 + n > (uint64_t)INT32_MAX) return fail("edge id overflow" + where);
        for (int k = 0; k < n; ++k) ids.push_back((int32_t)(e0 + k));
      } else if ((header & 2) == 0) {
        cc = (header & 4) != 0;
        const uint64_t e0 = header >> 3;
        if (e0 > (uint64_t)INT32_MAX) return fail("edge id overflow" + where);
        ids.push_back((int32_t)e0);
      } else {
        cc = (header & 4) != 0;
        const uint64_t n = header >> 3;
        if (n > (uint64_t)INT32_MAX || !decode_edges(&cr, (uint32_t)n, soup.shape_num_edges[0], &ids)) return fail("bad edge list" + where);
      }
      if (!add_clip(0, cc, ids)) return fail("shape or edge id out of range" + where);
    } else {
      uint32_t header = 0;
      if (!cr.varint32(&header)) return fail("truncated record" + where);
      uint32_t num_clipped = 1;
      if ((header & 7) == 3) {
        num_clipped = header >> 3;
        if (!cr.varint32(&header)) return fail("truncated record" + where);
      }
      int64_t shape_id = 0;
      for (uint32_t j = 0; j < num_clipped; ++j, ++shape_id) {
        if (j > 0 && !cr.varint32(&header)) return fail("truncated record" + where);
        ids.clear();
        bool cc;
        if ((header & 1) == 0) {
          uint32_t shape_id_count = 0;
          if (!cr.varint32(&shape_id_count)) return fail("truncated record" + where);
          shape_id += shape_id_count >> 4;
          const int n = (int)(shape_id_count & 15) + 1;
          cc = (header & 2) != 0;
          const uint64_t e0 = header >> 2;
          if (e0 + n > (uint64_t)INT32_MAX) return fail("edge id overflow" + where);
          for (int k = 0; k < n; ++k) ids.push_back((int32_t)(e0 + k));
        } else if ((header & 7) == 7) {
          shape_id += header >> 4;
          cc = (header & 8) != 0;
        } else {
          if ((header & 3) != 1) return fail("bad clipped-shape tag" + where);
          uint32_t shape_delta = 0;
          if (!cr.varint32(&shape_delta)) return fail("truncated record" + where);
          shape_id += shape_delta;
          const uint32_t n = (header >> 3) + 1;
          cc = (header & 4) != 0;
          if (shape_id >= num_shapes) return fail("shape id out of range" + where);
          if (!decode_edges(&cr, n, soup.shape_num_edges[shape_id], &ids)) return fail("bad edge list" + where);
        }
        if (!add_clip(shape_id, cc, ids)) return fail("shape or edge id out of range" + where);
      }
    }
    out->cell_clip_begin.push_back((uint32_t)out->clip_shape_id.size());
  }
  return S2B_OK;
}

// Encodes flat arrays (with shape-local edge ids) exactly as MutableS2ShapeIndex::Encode would encode an index with
// these cells: the output is byte-identical for identical cells (tests/test_index_wire.py).
inline int encode_index_wire(const S2bFlatIndex& f, const int32_t* edge_id, int max_edges_per_cell, std::vector<uint8_t>* bytes,
                             std::string* err) {
  using namespace wire;
  if (max_edges_per_cell <= 0) max_edges_per_cell = 10;
  if (f.num_edges && !edge_id) { *err = "index encode: edge ids are required"; return S2B_EINVAL; }
  bytes->clear();
  Writer w{bytes};
  w.varint64((uint64_t)max_edges_per_cell << 2 | 0);
  const uint64_t C = f.num_cells;
  // EncodeS2CellIdVector (encoded_s2cell_id_vector.cc:42-95)
  uint64_t v_or = 0, v_and = ~0ull, v_min = ~0ull, v_max = 0;
  for (uint64_t c = 0; c < C; ++c) {
    const uint64_t id = f.cell_ids[c];
    v_or |= id; v_and &= id;
    if (id < v_min) v_min = id;
    if (id > v_max) v_max = id;
  }
  uint64_t e_base = 0;
  int e_base_len = 0, e_shift = 0, e_max_delta_msb = 0;
  if (v_or > 0) {
    e_shift = __builtin_ctzll(v_or) & ~1;
    if (e_shift > 56) e_shift = 56;
    if (v_and & (1ull << e_shift)) ++e_shift;
    uint64_t e_bytes = ~0ull;
    for (int len = 0; len < 8; ++len) {
      const uint64_t t_base = v_min & ~(~0ull >> (8 * len));
      const uint64_t span = (v_max - t_base) >> e_shift;
      const int width = span ? 64 - __builtin_clzll(span) : 0;
      const int t_max_delta_msb = width - 1 > 0 ? width - 1 : 0;
      const uint64_t t_bytes = len + C * ((t_max_delta_msb >> 3) + 1);
      if (t_bytes < e_bytes) { e_base = t_base; e_base_len = len; e_max_delta_msb = t_max_delta_msb; e_bytes = t_bytes; }
    }
    if ((e_shift & 1) && (e_max_delta_msb & 7) != 7) --e_shift;
  }
  {  // EncodeBaseShift
    int shift_code = e_shift >> 1;
    if (e_shift & 1) shift_code = shift_code + 29 < 31 ? shift_code + 29 : 31;
    w.put8((uint32_t)(shift_code << 3) | (uint32_t)e_base_len);
    if (shift_code == 31) w.put8((uint32_t)(e_shift >> 1));
    const uint64_t base_bytes = e_base >> (64 - 8 * (e_base_len > 1 ? e_base_len : 1));
    w.uint_with_length(base_bytes, e_base_len);
  }
  std::vector<uint64_t> tmp(C);
  for (uint64_t c = 0; c < C; ++c) tmp[c] = (f.cell_ids[c] - e_base) >> e_shift;
  w.uint_vector(tmp);

  // one record per cell (S2ShapeIndexCell::Encode, s2shape_index.cc:69-195), gathered by a StringVectorEncoder
  std::vector<uint8_t> records;
  Writer rw{&records};
  auto encode_edges = [&](const int32_t* ids, int n) {   // s2shape_index.cc:292-323
    int base = 0;
    for (int i = 0; i < n; ++i) {
      const int e = ids[i], delta = e - base;
      if (i + 1 == n) {
        rw.varint32((uint32_t)delta);
      } else {
        int count = 1;
        for (; i + 1 < n && ids[i + 1] == e + count; ++i) ++count;
        if (count < 8) rw.varint32((uint32_t)delta << 3 | (uint32_t)(count - 1));
        else { rw.varint32((uint32_t)(count - 8) << 3 | 7); rw.varint32((uint32_t)delta); }
        base = e + count;
      }
    }
  };
  for (uint64_t c = 0; c < C; ++c) {
    const uint32_t kb = f.cell_clip_begin[c], ke = f.cell_clip_begin[c + 1];
    if (f.num_shape_ids == 1) {
      if (ke - kb != 1 || f.clip_shape_id[kb] != 0) { *err = "index encode: a single-shape index must have one clip of shape 0 per cell"; return S2B_EINVAL; }
      const int32_t* ids = edge_id + f.clip_edge_begin[kb];
      const int n = (int)(f.clip_edge_begin[kb + 1] - f.clip_edge_begin[kb]);
      const uint64_t cc = f.clip_flags[kb] & 1;
      if (n >= 2 && n <= 17 && ids[n - 1] - ids[0] == n - 1) rw.varint64((uint64_t)ids[0] << 6 | (uint64_t)(n - 2) << 2 | cc << 1 | 0);
      else if (n == 1) rw.varint64((uint64_t)ids[0] << 3 | cc << 2 | 1);
      else { rw.varint64((uint64_t)n << 3 | cc << 2 | 3); encode_edges(ids, n); }
    } else {
      if (ke - kb > 1) rw.varint32((ke - kb) << 3 | 3);
      int shape_id_base = 0;
      for (uint32_t k = kb; k < ke; ++k) {
        const int shape_delta = f.clip_shape_id[k] - shape_id_base;
        if (shape_delta < 0) { *err = "index encode: clip shape ids must ascend within a cell"; return S2B_EINVAL; }
        shape_id_base = f.clip_shape_id[k] + 1;
        const int32_t* ids = edge_id + f.clip_edge_begin[k];
        const int n = (int)(f.clip_edge_begin[k + 1] - f.clip_edge_begin[k]);
        const uint32_t cc = f.clip_flags[k] & 1;
        if (n >= 1 && n <= 16 && ids[n - 1] - ids[0] == n - 1) {
          rw.varint32((uint32_t)ids[0] << 2 | cc << 1 | 0);
          rw.varint32((uint32_t)shape_delta << 4 | (uint32_t)(n - 1));
        } else if (n == 0) {
          rw.varint32((uint32_t)shape_delta << 4 | cc << 3 | 7);
        } else {
          rw.varint32((uint32_t)(n - 1) << 3 | cc << 2 | 1);
          rw.varint32((uint32_t)shape_delta);
          encode_edges(ids, n);
        }
      }
    }
    tmp[c] = records.size();   // end offset of record c
  }
  w.uint_vector(tmp);
  bytes->insert(bytes->end(), records.begin(), records.end());
  return S2B_OK;
}

}  // namespace s2b
