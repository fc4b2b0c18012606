// s2b_shapes_wire.h — reader / writer for the reference's serialized shapes (host C++, no CUDA): the geometry half of
// a serialized index (the cells are s2b_index_wire.h).
//   s2shapeutil::EncodeTaggedShapes / TaggedShapeFactory   s2shapeutil_coding.cc:134-172  (StringVector of [varint tag, shape])
//   S2PointVectorShape   (tag 3)  s2point_vector_shape.h:62-73     EncodedS2PointVector
//   S2LaxPolylineShape   (tag 4)  s2lax_polyline_shape.cc:74-86    EncodedS2PointVector
//   S2LaxPolygonShape    (tag 5)  s2lax_polygon_shape.cc:177-252   version, num_loops, EncodedS2PointVector, loop starts
//   EncodedS2PointVector           encoded_s2point_vector.cc:121-134 (format), :205-250 (UNCOMPRESSED), :836-942 (CELL_IDS)
//   S2Polygon::Shape     (tag 1)  s2polygon.cc:787-860 (lossless, version 1), :1485-1531 (compressed, version 4);
//                                 loops: s2loop.cc:640-690, :1377-1436; edges are ORIENTED (holes reversed): s2polygon.cc Shape::chain_edge
//   S2Polyline::Shape    (tag 2)  s2polyline.cc:425-475 (lossless, version 1), :527-560 (compressed, version 2)
//   S2DecodePointsCompressed       s2point_compression.cc:330-391 (face runs, 2nd-order delta coded (pi, qi), off-centre points)
// Decoding handles both point-vector formats (the CELL_IDS decoder is restated from the format description and the
// reference decoder's arithmetic) and both encodings of S2Polygon / S2Polyline; encoding writes the lax shapes with
// either coding hint (FAST = UNCOMPRESSED point vectors, COMPACT = CELL_IDS), byte-identical to the reference. Of an S2Polygon only what the shape exposes is kept:
// its loops as oriented vertex chains (the depth decides the orientation); bounds and origin_inside are skipped.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "s2b_index_wire.h"

namespace s2b {

struct ShapeSoupData {
  std::vector<int8_t> shape_dim;
  std::vector<uint32_t> shape_chain_begin{0};
  std::vector<uint32_t> chain_vertex_begin{0};
  std::vector<double> vertices;
};

namespace wire {

inline void deinterleave_bit_pairs(uint64_t code, uint32_t* val0, uint32_t* val1) {   // encoded_s2point_vector.cc:71-92
  uint64_t v0 = code, v1 = code >> 2;
  v0 &= 0x3333333333333333ull; v0 |= v0 >> 2;  v1 &= 0x3333333333333333ull; v1 |= v1 >> 2;
  v0 &= 0x0f0f0f0f0f0f0f0full; v0 |= v0 >> 4;  v1 &= 0x0f0f0f0f0f0f0f0full; v1 |= v1 >> 4;
  v0 &= 0x00ff00ff00ff00ffull; v0 |= v0 >> 8;  v1 &= 0x00ff00ff00ff00ffull; v1 |= v1 >> 8;
  v0 &= 0x0000ffff0000ffffull; v0 |= v0 >> 16; v1 &= 0x0000ffff0000ffffull; v1 |= v1 >> 16;
  *val0 = (uint32_t)v0; *val1 = (uint32_t)v1;
}

inline uint64_t load_le(const uint8_t* p, int n) { uint64_t v = 0; for (int k = 0; k < n; ++k) v |= (uint64_t)p[k] << (8 * k); return v; }

// EncodedS2PointVector::Init + Decode: appends the points (3 doubles each) to *out.
inline bool decode_point_vector(Reader* r, std::vector<double>* out, std::string* why) {
  auto fail = [&](const char* m) { *why = m; return false; };
  if (r->avail() < 1) return fail("truncated point vector");
  const int format = *r->p & 7;
  if (format == 0) {   // UNCOMPRESSED: varint64 (size << 3 | 0), then size * 24 raw bytes
    uint64_t size;
    if (!r->varint64(&size)) return fail("truncated point vector size");
    size >>= 3;
    if (size > 0x7FFFFFFFull || r->avail() < size * 24) return fail("truncated uncompressed points");
    const size_t at = out->size();
    out->resize(at + 3 * size);
    if (size) std::memcpy(out->data() + at, r->p, size * 24);
    r->p += size * 24;
    return true;
  }
  if (format != 1) return fail("unknown point vector format");
  // CELL_IDS: two header bytes, base, then a string vector of 16-point blocks
  if (r->avail() < 2) return fail("truncated cell-id point vector");
  uint32_t h1, h2;
  r->get8(&h1); r->get8(&h2);
  const bool have_exceptions = (h1 & 8) != 0;
  const int last_block_count = (int)(h1 >> 4) + 1;
  const int base_bytes = (int)(h2 & 7);
  const int level = (int)(h2 >> 3);
  if (level > 30) return fail("bad cell-id level");
  uint64_t base;
  if (!r->uint_with_length(base_bytes, &base)) return fail("truncated base");
  const int base_shift = 2 * level + 3 - 8 * base_bytes;
  base <<= base_shift > 0 ? base_shift : 0;
  UintVector offsets;
  if (!offsets.init(r)) return fail("truncated block offsets");
  const uint64_t nblocks = offsets.size;
  const uint64_t data_len = nblocks ? offsets[nblocks - 1] : 0;
  if (r->avail() < data_len) return fail("truncated blocks");
  const uint8_t* data = r->p;
  r->p += data_len;
  if (nblocks == 0) return fail("cell-id point vector without blocks");
  const uint64_t size = 16 * (nblocks - 1) + (uint64_t)last_block_count;
  if (size > 0x7FFFFFFFull) return fail("point vector too large");
  const size_t at = out->size();
  out->resize(at + 3 * size);
  double* dst = out->data() + at;
  for (uint64_t i = 0; i < size; ++i) {
    const uint64_t b = i >> 4;
    const uint64_t bstart = b ? offsets[b - 1] : 0, bend = offsets[b];
    if (bend < bstart || bend > data_len || bend == bstart) return fail("invalid block");
    const uint8_t* ptr = data + bstart;
    const uint8_t* end = data + bend;
    const uint8_t header = *ptr++;
    const int overlap_nibbles = (header >> 3) & 1;
    const int offset_bytes = (header & 7) + overlap_nibbles;
    const int delta_nibbles = (header >> 4) + 1;
    uint64_t offset = 0;
    if (offset_bytes > 0) {
      const int offset_shift = (delta_nibbles - overlap_nibbles) << 2;
      if (offset_shift >= 64 || ptr + offset_bytes > end) return fail("invalid block offset");
      offset = load_le(ptr, offset_bytes) << offset_shift;
      ptr += offset_bytes;
    }
    const int delta_nibble_offset = (int)(i & 15) * delta_nibbles;
    const int delta_bytes = (delta_nibbles + 1) >> 1;
    const uint8_t* delta_ptr = ptr + (delta_nibble_offset >> 1);
    if (delta_ptr + delta_bytes > end) return fail("invalid block delta");
    uint64_t delta = load_le(delta_ptr, delta_bytes);
    delta >>= (delta_nibble_offset & 1) << 2;
    delta &= (~0ull >> (64 - (delta_nibbles << 2)));
    if (have_exceptions) {
      if (delta < 16) {
        const uint64_t block_size = size - (i & ~15ull) < 16 ? size - (i & ~15ull) : 16;
        ptr += (block_size * delta_nibbles + 1) >> 1;
        ptr += delta * 24;
        if (ptr + 24 > end) return fail("invalid exception");
        std::memcpy(dst + 3 * i, ptr, 24);
        continue;
      }
      delta -= 16;
    }
    const uint64_t value = base + offset + delta;
    const int shift = 30 - level;
    uint32_t sj, tj;
    deinterleave_bit_pairs(value, &sj, &tj);
    const int si = (int)((((sj << 1) | 1u) << shift) & 0x7fffffffu);
    const int ti = (int)((((tj << 1) | 1u) << shift) & 0x7fffffffu);
    const int face = (int)(((sj << shift) >> 30) | (((tj << (shift + 1)) >> 29) & 4u));
    if (face > 5) return fail("invalid face in cell-id point");
    const P3 p = normalize(face_uv_to_xyz(face, st_to_uv((1.0 / 2147483648.0) * (unsigned)si), st_to_uv((1.0 / 2147483648.0) * (unsigned)ti)));
    dst[3 * i] = p.x; dst[3 * i + 1] = p.y; dst[3 * i + 2] = p.z;
  }
  return true;
}


// ---- EncodeS2PointVectorCompact (encoded_s2point_vector.cc:323-594): the CELL_IDS writer -------------------------------
// The format is restated from its description there (two header bytes, base, a string vector of 16-value blocks, each
// block = header byte, offset, nibble-packed deltas, exceptions); the parameter choices (level, base, per-block delta /
// offset / overlap sizes) have to be the reference's own for the bytes to be identical, so ChooseBestLevel (:597-632),
// ConvertCellsToValues (:639-664), ChooseBase (:666-712), CanEncode (:716-731) and GetBlockCode (:736-830) are followed.

inline uint64_t bit_mask(int n) { return n == 0 ? 0 : (~0ull >> (64 - n)); }
inline int bit_width64(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }

inline uint64_t interleave_bit_pairs(uint32_t val0, uint32_t val1) {   // encoded_s2point_vector.cc:53-66
  uint64_t v0 = val0, v1 = val1;
  v0 = (v0 | (v0 << 16)) & 0x0000ffff0000ffffull; v1 = (v1 | (v1 << 16)) & 0x0000ffff0000ffffull;
  v0 = (v0 | (v0 << 8)) & 0x00ff00ff00ff00ffull;  v1 = (v1 | (v1 << 8)) & 0x00ff00ff00ff00ffull;
  v0 = (v0 | (v0 << 4)) & 0x0f0f0f0f0f0f0f0full;  v1 = (v1 | (v1 << 4)) & 0x0f0f0f0f0f0f0f0full;
  v0 = (v0 | (v0 << 2)) & 0x3333333333333333ull;  v1 = (v1 | (v1 << 2)) & 0x3333333333333333ull;
  return v0 | (v1 << 2);
}

// S2::XYZtoFaceSiTi (s2coords.cc:43-66): the level at which p is exactly a cell centre, or -1.
inline int xyz_to_face_si_ti(const P3& p, int* face, uint32_t* si, uint32_t* ti) {
  double u, v;
  *face = xyz_to_face_uv(p, &u, &v);
  auto st_to_si_ti = [](double s) {   // s2coords.h:363-366 with MathUtil::Round<int64_t> (mathutil.h:71-79)
    const double x = s * 2147483648.0;
    return (uint32_t)(int64_t)(x < 0 ? x - 0.5 : x + 0.5);
  };
  *si = st_to_si_ti(uv_to_st(u));
  *ti = st_to_si_ti(uv_to_st(v));
  const int level = 30 - __builtin_ctz(*si | 0x80000000u);
  if (level < 0 || level != 30 - __builtin_ctz(*ti | 0x80000000u)) return -1;
  const P3 c = normalize(face_uv_to_xyz(*face, st_to_uv((1.0 / 2147483648.0) * *si), st_to_uv((1.0 / 2147483648.0) * *ti)));
  return eq(p, c) ? level : -1;
}

struct BlockCode { int delta_bits, offset_bits, overlap_bits; };

inline bool can_encode(uint64_t d_min, uint64_t d_max, int delta_bits, int overlap_bits, bool have_exceptions) {
  d_min &= ~bit_mask(delta_bits - overlap_bits);
  uint64_t max_delta = bit_mask(delta_bits);
  if (have_exceptions) {
    if (max_delta < 16) return false;
    max_delta -= 16;
  }
  return (d_min > ~max_delta) || (d_min + max_delta >= d_max);
}

inline BlockCode block_code(const uint64_t* values, int count, uint64_t base, bool have_exceptions) {
  uint64_t b_min = ~0ull, b_max = 0;
  for (int k = 0; k < count; ++k)
    if (values[k] != ~0ull) { b_min = std::min(b_min, values[k]); b_max = std::max(b_max, values[k]); }
  if (b_min == ~0ull) return BlockCode{4, 0, 0};
  b_min -= base; b_max -= base;
  int delta_bits = (std::max(1, bit_width64(b_max - b_min) - 1) + 3) & ~3;
  int overlap_bits = 0;
  if (!can_encode(b_min, b_max, delta_bits, 0, have_exceptions)) {
    if (can_encode(b_min, b_max, delta_bits, 4, have_exceptions)) {
      overlap_bits = 4;
    } else {
      delta_bits += 4;
      if (!can_encode(b_min, b_max, delta_bits, 0, have_exceptions)) overlap_bits = 4;
    }
  }
  if (count == 1 && !have_exceptions) delta_bits = 8;
  const uint64_t max_delta = bit_mask(delta_bits) - (have_exceptions ? 16 : 0);
  int offset_bits = 0;
  if (b_max > max_delta) {
    const int offset_shift = delta_bits - overlap_bits;
    const uint64_t mask = bit_mask(offset_shift);
    const uint64_t min_offset = (b_max - max_delta + mask) & ~mask;
    offset_bits = (bit_width64(min_offset) - offset_shift + 7) & ~7;
    if (offset_bits == 64) overlap_bits = 4;
  }
  return BlockCode{delta_bits, offset_bits, overlap_bits};
}

inline void encode_point_vector_fast(const double* pts, uint64_t n, std::vector<uint8_t>* out) {   // :205-213
  Writer w{out};
  w.varint64(n << 3 | 0);
  const uint8_t* src = (const uint8_t*)pts;
  out->insert(out->end(), src, src + n * 24);
}

inline void encode_point_vector_compact(const double* pts, uint64_t n, std::vector<uint8_t>* out) {
  // 1. the level at which most points are cell centres (or give up: fewer than 5 % are)
  struct CellPoint { int8_t level, face; uint32_t si, ti; };
  std::vector<CellPoint> cps((size_t)n);
  int level_counts[31] = {0};
  for (uint64_t k = 0; k < n; ++k) {
    int face; uint32_t si, ti;
    const int level = xyz_to_face_si_ti(P3{pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]}, &face, &si, &ti);
    cps[k] = CellPoint{(int8_t)level, (int8_t)face, si, ti};
    if (level >= 0) ++level_counts[level];
  }
  int level = 0;
  for (int l = 1; l <= 30; ++l) if (level_counts[l] > level_counts[level]) level = l;
  if (level_counts[level] <= 0.05 * (double)n) return encode_point_vector_fast(pts, n, out);
  // 2. the 64-bit values: (face, si, ti) at that level with bit pairs interleaved; others are exceptions
  std::vector<uint64_t> values((size_t)n);
  bool have_exceptions = false;
  const int shift = 30 - level;
  for (uint64_t k = 0; k < n; ++k) {
    const CellPoint& cp = cps[k];
    if (cp.level != level) { values[k] = ~0ull; have_exceptions = true; continue; }
    const uint32_t sj = ((((uint32_t)cp.face & 3u) << 30) | (cp.si >> 1)) >> shift;
    const uint32_t tj = ((((uint32_t)cp.face & 4u) << 29) | cp.ti) >> (shift + 1);
    values[k] = interleave_bit_pairs(sj, tj);
  }
  // 3. base: the prefix shared by all values, at most 7 bytes of it
  const int max_bits = 2 * level + 3;
  auto base_shift_for = [&](int base_bits) { return std::max(0, max_bits - base_bits); };
  uint64_t v_min = ~0ull, v_max = 0;
  for (uint64_t v : values) if (v != ~0ull) { v_min = std::min(v_min, v); v_max = std::max(v_max, v); }
  uint64_t base = 0;
  int base_bits = 0;
  if (v_min != ~0ull) {
    const int min_delta_bits = (have_exceptions || n == 1) ? 8 : 4;
    const int excluded_bits = std::max(bit_width64(v_min ^ v_max), std::max(min_delta_bits, base_shift_for(56)));
    const uint64_t prefix = v_min & ~bit_mask(excluded_bits);
    if (prefix != 0) base_bits = (max_bits - __builtin_ctzll(prefix) + 7) & ~7;
    base = v_min & ~bit_mask(base_shift_for(base_bits));
  }
  const uint64_t num_blocks = (n + 15) >> 4;
  const int base_bytes = base_bits >> 3;
  const int last_block_count = (int)((int64_t)n - 16 * ((int64_t)num_blocks - 1));
  Writer w{out};
  w.put8(1u | ((uint32_t)have_exceptions << 3) | ((uint32_t)(last_block_count - 1) << 4));
  w.put8((uint32_t)base_bytes | ((uint32_t)level << 3));
  w.uint_with_length(base >> base_shift_for(base_bits), base_bytes);
  // 4. the blocks
  std::vector<uint8_t> blocks;
  std::vector<uint64_t> ends;
  Writer b{&blocks};
  for (uint64_t i = 0; i < n; i += 16) {
    const int block_size = (int)std::min<uint64_t>(16, n - i);
    const BlockCode code = block_code(&values[i], block_size, base, have_exceptions);
    const int offset_bytes = code.offset_bits >> 3, delta_nibbles = code.delta_bits >> 2, overlap_nibbles = code.overlap_bits >> 2;
    b.put8((uint32_t)(offset_bytes - overlap_nibbles) | ((uint32_t)overlap_nibbles << 3) | ((uint32_t)(delta_nibbles - 1) << 4));
    uint64_t offset = ~0ull;
    int num_exceptions = 0;
    for (int j = 0; j < block_size; ++j) {
      if (values[i + j] == ~0ull) ++num_exceptions;
      else offset = std::min(offset, values[i + j] - base);
    }
    if (num_exceptions == block_size) offset = 0;
    const int offset_shift = code.delta_bits - code.overlap_bits;
    offset &= ~bit_mask(offset_shift);
    if (offset > 0) b.uint_with_length(offset >> offset_shift, offset_bytes);
    const int delta_bytes = (delta_nibbles + 1) >> 1;
    int exceptions = 0;
    for (int j = 0; j < block_size; ++j) {
      uint64_t delta;
      if (values[i + j] == ~0ull) {
        delta = (uint64_t)exceptions++;
      } else {
        delta = values[i + j] - (offset + base);
        if (have_exceptions) delta += 16;
      }
      if ((delta_nibbles & 1) && (j & 1)) {   // odd nibble count: share a byte with the previous delta's high nibble
        const uint8_t last = blocks.back();
        blocks.pop_back();
        delta = (delta << 4) | (last & 0xf);
      }
      b.uint_with_length(delta, delta_bytes);
    }
    for (int j = 0; j < block_size; ++j)
      if (values[i + j] == ~0ull) {
        const uint8_t* src = (const uint8_t*)(pts + 3 * (i + j));
        blocks.insert(blocks.end(), src, src + 24);
      }
    ends.push_back(blocks.size());
  }
  w.uint_vector(ends);          // StringVectorEncoder::Encode: offsets (EncodeUintVector<uint64_t>), then the data
  out->insert(out->end(), blocks.begin(), blocks.end());
}

}  // namespace wire

namespace wire {

// S2DecodePointsCompressed (s2point_compression.cc:330-391): n points snapped to `level`, appended to *out.
inline bool decode_points_compressed(Reader* r, int level, uint64_t n, std::vector<double>* out, std::string* why) {
  auto fail = [&](const char* m) { *why = m; return false; };
  if (level > 30) return fail("bad snap level");
  if (n > (uint64_t)r->avail() + 1) return fail("more points than bytes");   // every point but the first takes >= 1 byte
  std::vector<uint8_t> faces;   // one face per point, from the run-length list (Faces::Decode :131-142)
  faces.reserve((size_t)(n < (1u << 20) ? n : (1u << 20)));
  while (faces.size() < n) {
    uint64_t face_and_count;
    if (!r->varint64(&face_and_count)) return fail("truncated face runs");
    const uint64_t count = face_and_count / 6;
    if (count == 0 || count > 0x7FFFFFFFull) return fail("bad face run");
    if (count > n - faces.size() && r->avail() == 0) return fail("face run beyond the points");
    // the reference accepts a last run that overshoots; only the faces of the n points are used
    const uint64_t take = count < n - faces.size() ? count : n - faces.size();
    faces.insert(faces.end(), (size_t)take, (uint8_t)(face_and_count % 6));
  }
  // NthDerivativeCoder(2)::Decode (util/coding/nth-derivative.h:122-128): the order in use grows 1, 2, 2, ... with the
  // calls; for (i = used - 1 .. 0) k = memory[i] = memory[i] + k, in wrap-around 32-bit arithmetic.
  struct Nth { int32_t m[2] = {0, 0}; int used = 0; } pm, qm;
  auto nth_decode = [](Nth& c, int32_t k) {
    if (c.used < 2) ++c.used;
    for (int i = c.used - 1; i >= 0; --i) { k = (int32_t)((uint32_t)k + (uint32_t)c.m[i]); c.m[i] = k; }
    return k;
  };
  auto deinterleave = [](uint64_t code, uint32_t* even, uint32_t* odd) {   // util_bits::DeinterleaveUint32: val0 = even bits, val1 = odd bits
    auto compact = [](uint64_t x) {
      x &= 0x5555555555555555ull; x = (x | (x >> 1)) & 0x3333333333333333ull; x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
      x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull; x = (x | (x >> 8)) & 0x0000ffff0000ffffull; x = (x | (x >> 16)) & 0xffffffffull;
      return (uint32_t)x;
    };
    *even = compact(code); *odd = compact(code >> 1);
  };
  const size_t at = out->size();
  out->resize(at + 3 * (size_t)n);
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t a, b;
    int32_t pi, qi;
    if (i == 0) {   // DecodeFirstPointFixedLength: (level + 7) / 8 * 2 bytes, interleaved, little endian
      const int bytes = (level + 7) / 8 * 2;
      uint64_t code;
      if (!r->uint_with_length(bytes, &code)) return fail("truncated first point");
      deinterleave(code, &a, &b);
      pi = nth_decode(pm, (int32_t)a); qi = nth_decode(qm, (int32_t)b);
    } else {        // DecodePointCompressed: varint64 of interleaved zig-zag coded second differences
      uint64_t code;
      if (!r->varint64(&code)) return fail("truncated point");
      deinterleave(code, &a, &b);
      auto zz = [](uint32_t n) { return (int32_t)((n >> 1) ^ (uint32_t)(-(int32_t)(n & 1))); };
      pi = nth_decode(pm, zz(a)); qi = nth_decode(qm, zz(b));
    }
    // FacePiQitoXYZ: FaceUVtoXYZ(face, STtoUV((pi + 0.5) / 2^level), STtoUV((qi + 0.5) / 2^level)).Normalize()
    const double scale = (double)(1 << level);
    const P3 p = normalize(face_uv_to_xyz(faces[i], st_to_uv((pi + 0.5) / scale), st_to_uv((qi + 0.5) / scale)));
    double* q = out->data() + at + 3 * i;
    q[0] = p.x; q[1] = p.y; q[2] = p.z;
  }
  uint32_t num_off_center;
  if (!r->varint32(&num_off_center) || num_off_center > n) return fail("bad off-centre count");
  for (uint32_t k = 0; k < num_off_center; ++k) {
    uint32_t index;
    if (!r->varint32(&index) || index >= n) return fail("bad off-centre index");
    if (r->avail() < 24) return fail("truncated off-centre point");
    std::memcpy(out->data() + at + 3 * (size_t)index, r->p, 24);
    r->p += 24;
  }
  return true;
}

inline bool skip_latlng_rect(Reader* r) {   // S2LatLngRect::Encode: version byte + 4 doubles (s2latlng_rect.cc)
  if (r->avail() < 33) return false;
  r->p += 33;
  return true;
}

// One decoded S2Loop -> an oriented chain appended to the soup (S2Polygon::Shape: holes run backwards, the full loop is a
// chain without vertices, empty loops are dropped as S2Polygon::Decode drops them).
inline void append_oriented_loop(ShapeSoupData* out, std::vector<double>* loop, uint32_t depth) {
  const size_t nv = loop->size() / 3;
  if (nv == 0) return;
  if (nv == 1) {                                  // S2Loop::is_empty_or_full(): empty if z >= 0, full if z < 0
    if (!((*loop)[2] < 0)) return;
    out->chain_vertex_begin.push_back((uint32_t)(out->vertices.size() / 3));
    return;
  }
  if (depth & 1)                                  // is_hole(): oriented_vertex(i) = vertex(n - 1 - i)
    for (size_t i = 0; i < nv / 2; ++i)
      for (int c = 0; c < 3; ++c) std::swap((*loop)[3 * i + c], (*loop)[3 * (nv - 1 - i) + c]);
  out->vertices.insert(out->vertices.end(), loop->begin(), loop->end());
  out->chain_vertex_begin.push_back((uint32_t)(out->vertices.size() / 3));
}

// S2Polygon::Decode (tag 1 payload).
inline bool decode_s2polygon(Reader* r, ShapeSoupData* out, std::string* why) {
  auto fail = [&](const char* m) { *why = m; return false; };
  uint32_t version;
  if (!r->get8(&version)) return fail("truncated polygon");
  std::vector<double> loop;
  if (version == 1) {                             // DecodeUncompressed
    uint32_t owns, holes;
    uint64_t num_loops;
    if (!r->get8(&owns) || !r->get8(&holes) || !r->uint_with_length(4, &num_loops)) return fail("truncated polygon header");
    if (num_loops > 10000000) return fail("too many loops");
    for (uint64_t l = 0; l < num_loops; ++l) {    // S2Loop::Decode
      uint32_t lv, origin_inside;
      uint64_t nv, depth;
      if (!r->get8(&lv) || lv != 1) return fail("bad loop version");
      if (!r->uint_with_length(4, &nv) || nv > 50000000 || r->avail() < nv * 24 + 5) return fail("truncated loop");
      loop.resize(3 * (size_t)nv);
      if (nv) std::memcpy(loop.data(), r->p, (size_t)nv * 24);
      r->p += nv * 24;
      r->get8(&origin_inside);
      if (!r->uint_with_length(4, &depth)) return fail("truncated loop");
      if (!skip_latlng_rect(r)) return fail("truncated loop bound");
      append_oriented_loop(out, &loop, (uint32_t)depth);
    }
    if (!skip_latlng_rect(r)) return fail("truncated polygon bound");
    return true;
  }
  if (version != 4) return fail("unknown polygon encoding version");
  uint32_t snap_level, num_loops;                 // DecodeCompressed
  if (!r->get8(&snap_level) || snap_level > 30) return fail("bad snap level");
  if (!r->varint32(&num_loops) || num_loops > 10000000) return fail("bad loop count");
  for (uint32_t l = 0; l < num_loops; ++l) {      // S2Loop::DecodeCompressed
    uint32_t nv, properties, depth;
    if (!r->varint32(&nv) || nv == 0 || nv > 50000000) return fail("bad vertex count");
    loop.clear();
    if (!decode_points_compressed(r, (int)snap_level, nv, &loop, why)) return false;
    if (!r->varint32(&properties) || !r->varint32(&depth)) return fail("truncated loop properties");
    if ((properties & 2) && !skip_latlng_rect(r)) return fail("truncated loop bound");   // kBoundEncoded = bit 1
    append_oriented_loop(out, &loop, depth);
  }
  return true;
}

// S2Polyline::Decode (tag 2 payload): the vertices as one chain.
inline bool decode_s2polyline(Reader* r, ShapeSoupData* out, std::string* why) {
  auto fail = [&](const char* m) { *why = m; return false; };
  uint32_t version;
  if (!r->get8(&version)) return fail("truncated polyline");
  if (version == 1) {
    uint64_t nv;
    if (!r->uint_with_length(4, &nv) || nv > 50000000 || r->avail() < nv * 24) return fail("truncated polyline");
    const size_t at = out->vertices.size();
    out->vertices.resize(at + 3 * (size_t)nv);
    if (nv) std::memcpy(out->vertices.data() + at, r->p, (size_t)nv * 24);
    r->p += nv * 24;
    return true;
  }
  if (version != 2) return fail("unknown polyline encoding version");
  uint32_t snap_level, nv;
  if (!r->get8(&snap_level) || snap_level > 30) return fail("bad snap level");
  if (!r->varint32(&nv) || nv > 50000000) return fail("bad vertex count");
  return nv == 0 || decode_points_compressed(r, (int)snap_level, nv, &out->vertices, why);
}

}  // namespace wire


// Decodes s2shapeutil::{Fast,Compact}EncodeTaggedShapes output into a shape soup (dimension-2 shapes: one chain per loop).
inline int decode_tagged_shapes(const uint8_t* bytes, size_t len, ShapeSoupData* out, size_t* consumed, std::string* err) {
  using namespace wire;
  auto fail = [&](const std::string& m) { *err = "shape decode: " + m; return S2B_EINVAL; };
  if (!bytes && len) return fail("null buffer");
  *out = ShapeSoupData();
  Reader r{bytes, bytes + len};
  UintVector offsets;   // EncodedStringVector
  if (!offsets.init(&r)) return fail("truncated shape offsets");
  const uint64_t S = offsets.size;
  const uint64_t data_len = S ? offsets[S - 1] : 0;
  if (r.avail() < data_len) return fail("truncated shapes");
  const uint8_t* data = r.p;
  r.p += data_len;
  if (consumed) *consumed = (size_t)(r.p - bytes);
  if (S > 0x7FFFFFFFull) return fail("too many shapes");
  uint64_t start = 0;
  for (uint64_t s = 0; s < S; ++s) {
    const uint64_t limit = offsets[s];
    if (limit < start || limit > data_len) return fail("bad shape offsets");
    Reader sr{data + start, data + limit};
    start = limit;
    const std::string where = " (shape " + std::to_string(s) + ")";
    std::string why;
    auto end_chain = [&]() { out->chain_vertex_begin.push_back((uint32_t)(out->vertices.size() / 3)); };
    if (sr.avail() == 0) {               // a removed (null) shape keeps its id: an empty point set
      out->shape_dim.push_back(0);
      end_chain();
    } else {
      uint32_t tag;
      if (!sr.varint32(&tag)) return fail("truncated type tag" + where);
      if (tag == 3 || tag == 4) {        // S2PointVectorShape / S2LaxPolylineShape: one point vector
        out->shape_dim.push_back(tag == 3 ? 0 : 1);
        if (!decode_point_vector(&sr, &out->vertices, &why)) return fail(why + where);
        end_chain();
      } else if (tag == 5) {             // S2LaxPolygonShape
        uint32_t version, num_loops;
        if (!sr.get8(&version) || version != 1) return fail("bad lax polygon version" + where);
        if (!sr.varint32(&num_loops) || num_loops > 0x7FFFFFFEu) return fail("bad loop count" + where);
        const size_t v0 = out->vertices.size() / 3;
        if (!decode_point_vector(&sr, &out->vertices, &why)) return fail(why + where);
        const size_t nv = out->vertices.size() / 3 - v0;
        out->shape_dim.push_back(2);
        if (num_loops == 0) {
          out->vertices.resize(3 * v0);
        } else if (num_loops == 1) {
          end_chain();
        } else {
          UintVector starts;   // EncodedUintVector<uint32_t>: the size field counts bytes of uint32
          uint64_t size_len;
          if (!sr.varint64(&size_len)) return fail("truncated loop starts" + where);
          starts.size = size_len / 4; starts.len = (int)(size_len & 3) + 1;
          if (starts.size != (uint64_t)num_loops + 1 || sr.avail() < starts.size * (uint64_t)starts.len) return fail("bad loop starts" + where);
          starts.data = sr.p;
          uint64_t prev = 0;
          for (uint32_t k = 0; k <= num_loops; ++k) {
            const uint64_t v = starts[k];
            if (v < prev || v > nv || (k == 0 && v != 0)) return fail("loop starts out of range" + where);
            if (k) out->chain_vertex_begin.push_back((uint32_t)(v0 + v));
            prev = v;
          }
          if (prev != nv) return fail("loop starts do not cover the vertices" + where);
        }
      } else if (tag == 1) {             // S2Polygon::Shape
        out->shape_dim.push_back(2);
        if (!decode_s2polygon(&sr, out, &why)) return fail(why + where);
      } else if (tag == 2) {             // S2Polyline::Shape
        out->shape_dim.push_back(1);
        if (!decode_s2polyline(&sr, out, &why)) return fail(why + where);
        end_chain();
      } else {
        return fail("unsupported shape type tag " + std::to_string(tag) + where);
      }
    }
    out->shape_chain_begin.push_back((uint32_t)(out->chain_vertex_begin.size() - 1));
  }
  return S2B_OK;
}

// Encodes a shape soup as s2shapeutil::FastEncodeTaggedShapes (compact = false: uncompressed point vectors) or
// s2shapeutil::CompactEncodeTaggedShapes (compact = true: CELL_IDS point vectors where at least 5 % of a vector's points
// are cell centres) would encode the corresponding S2PointVectorShape / S2LaxPolylineShape / S2LaxPolygonShape objects.
inline int encode_tagged_shapes(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin, const uint32_t* chain_vertex_begin,
                                const double* vertices, std::vector<uint8_t>* bytes, std::string* err, bool compact = false) {
  using namespace wire;
  auto fail = [&](const std::string& m) { *err = "shape encode: " + m; return S2B_EINVAL; };
  std::vector<uint8_t> records;
  std::vector<uint64_t> ends((size_t)num_shapes);
  Writer w{&records};
  auto point_vector = [&](uint32_t vbegin, uint32_t vend) {   // EncodeS2PointVector(points, hint)
    if (compact) encode_point_vector_compact(vertices + 3 * (size_t)vbegin, vend - vbegin, &records);
    else encode_point_vector_fast(vertices + 3 * (size_t)vbegin, vend - vbegin, &records);
  };
  for (int s = 0; s < num_shapes; ++s) {
    const uint32_t cb = shape_chain_begin[s], ce = shape_chain_begin[s + 1];
    const uint32_t vb = chain_vertex_begin[cb], ve = chain_vertex_begin[ce];
    if (shape_dim[s] == 0) {
      w.varint32(3);
      point_vector(vb, ve);                       // all chains of a point shape form one point list (edge ids are kept)
    } else if (shape_dim[s] == 1) {
      if (ce - cb > 1) return fail("a polyline shape with several chains has no S2LaxPolylineShape form (shape " + std::to_string(s) + ")");
      w.varint32(4);
      point_vector(vb, ve);
    } else if (shape_dim[s] == 2) {
      w.varint32(5);
      w.put8(1);
      w.varint32(ce - cb);
      point_vector(vb, ve);
      if (ce - cb > 1) {                          // EncodeUintVector<uint32_t>(loop_starts, num_loops + 1)
        uint32_t one_bits = 1;
        for (uint32_t c = cb; c <= ce; ++c) one_bits |= chain_vertex_begin[c] - vb;
        const int len = ((31 - __builtin_clz(one_bits)) >> 3) + 1;
        w.varint64(((uint64_t)(ce - cb + 1) * 4) | (uint64_t)(len - 1));
        for (uint32_t c = cb; c <= ce; ++c) w.uint_with_length(chain_vertex_begin[c] - vb, len);
      }
    } else {
      return fail("shape dimension must be 0, 1 or 2");
    }
    ends[s] = records.size();
  }
  bytes->clear();
  Writer o{bytes};
  o.uint_vector(ends);
  bytes->insert(bytes->end(), records.begin(), records.end());
  return S2B_OK;
}

}  // namespace s2b
