// s2b_core.cuh — scalar building blocks of the S2 hot path, written once for device code (nvcc, sm_100a,
// -fmad=false) and also compilable as plain C++ (g++ -ffp-contract=off) so tests/hostsim can check the
// LOGIC against the oracle on a machine without a GPU. The product only ever runs the CUDA build.
//
// Arithmetic must reproduce the reference bit for bit: no FMA contraction except where
// fma() is spelled out (exact two-products), IEEE division and square root, and the reference's evaluation order.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define S2B_HD __host__ __device__ __forceinline__
#define S2B_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define S2B_HD inline
#define S2B_HD_NOINLINE inline
#endif

namespace s2b {

struct P3 { double x, y, z; };

S2B_HD bool eq(const P3& a, const P3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }  // -0 == +0 (vector.h operator==)
// Vector3::DotProd accumulates ((0 + a0*b0) + a1*b1) + a2*b2 (util/math/vector.h:315-318).
S2B_HD double dot(const P3& a, const P3& b) { return ((0.0 + a.x * b.x) + a.y * b.y) + a.z * b.z; }
// Vector3::CrossProd (util/math/vector.h:474-478).
S2B_HD P3 cross(const P3& a, const P3& b) { return P3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
S2B_HD P3 sub(const P3& a, const P3& b) { return P3{a.x - b.x, a.y - b.y, a.z - b.z}; }
S2B_HD P3 neg(const P3& a) { return P3{-a.x, -a.y, -a.z}; }
S2B_HD double norm2(const P3& a) { return dot(a, a); }
// Vector3::Normalize: one reciprocal then three multiplies (util/math/vector.h:191-198).
S2B_HD P3 normalize(const P3& a) {
  double n = sqrt(norm2(a));
  if (n != 0) n = 1.0 / n;
  return P3{a.x * n, a.y * n, a.z * n};
}
// Vector3::LargestAbsComponent (util/math/vector.h:508-513): strict '>' decides ties.
S2B_HD int largest_abs_component(const P3& p) {
  double ax = fabs(p.x), ay = fabs(p.y), az = fabs(p.z);
  return ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
}
S2B_HD double comp(const P3& p, int k) { return k == 0 ? p.x : (k == 1 ? p.y : p.z); }
// lexicographic operator< of Vector3 (used by ExactSign's sort).
S2B_HD bool lex_less(const P3& a, const P3& b) {
  if (a.x < b.x) return true; if (b.x < a.x) return false;
  if (a.y < b.y) return true; if (b.y < a.y) return false;
  return a.z < b.z;
}

// ---------------------------------------------------------------------------------------------------
// Cube-face projection: S2::XYZtoFaceUV (s2coords.h:389-419), quadratic UVtoST (:329-332), STtoIJ (:345-356).
S2B_HD int xyz_to_face_uv(const P3& p, double* u, double* v) {
  // GetFace (s2coords.h:409-413) + ValidFaceXYZtoUV (:389-403), written with selects instead of a six-way switch:
  //   face:  0: ( y/x,  z/x)  1: (-x/y,  z/y)  2: (-x/z, -y/z)  3: ( z/x,  y/x)  4: ( z/y, -x/y)  5: (-y/z, -x/z)
  // i.e. the major axis m, and numerators chosen per axis, swapped and sign-adjusted for the negative faces.
  // The unary minus is applied to the numerator, as in the reference; exactly two IEEE divisions are issued.
  const double ax = fabs(p.x), ay = fabs(p.y), az = fabs(p.z);
  const int axis = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  const double m = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
  const bool neg = m < 0;
  // positive faces: (a, b) = axis 0: (y, z), 1: (-x, z), 2: (-x, -y)
  const double a = axis == 0 ? p.y : -p.x;
  const double b = axis == 2 ? -p.y : p.z;
  // negative faces: axis 0: (z, y) = (b, a); axis 1: (z, -x) = (b, a); axis 2: (-y, -x) = (b, a)
  *u = (neg ? b : a) / m;
  *v = (neg ? a : b) / m;
  return axis + (neg ? 3 : 0);
}
S2B_HD double uv_to_st(double u) {
  // u >= 0 ? 0.5 * sqrt(1 + 3u) : 1 - 0.5 * sqrt(1 - 3u)   (s2coords.h:329-332). 1 - 3u for u < 0 equals 1 + |3u|
  // exactly (negation is exact), so one square-root site serves both branches.
  double r = 0.5 * sqrt(1 + fabs(3 * u));
  return u >= 0 ? r : 1 - r;
}
S2B_HD double st_to_uv(double s) {  // s2coords.h:324-327
  if (s >= 0.5) return (1 / 3.) * (4 * s * s - 1);
  return (1 / 3.) * (1 - 4 * (1 - s) * (1 - s));
}
S2B_HD uint32_t st_to_ij(double s) {
  if (!(s > 0)) return 0;  // also NaN
  int v = (int)(1073741824.0 * s);  // truncation, s <= 1
  return (uint32_t)(v < 1073741823 ? v : 1073741823);
}

// S2CellId::FromFaceIJ (s2cell_id.cc:267-307). `pos` is the 1024-entry iiiijjjjoo -> ppppppppoo table.
template <class TableT>
S2B_HD uint64_t from_face_ij(int face, uint32_t i, uint32_t j, const TableT* pos) {
  uint64_t n = (uint64_t)face << 60;
  uint32_t bits = (uint32_t)face & 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 7; k >= 0; --k) {
    bits += ((i >> (4 * k)) & 15u) << 6;
    bits += ((j >> (4 * k)) & 15u) << 2;
    bits = pos[bits];
    n |= (uint64_t)(bits >> 2) << (8 * k);
    bits &= 3u;
  }
  return n * 2 + 1;
}

// FromFaceIJ in two parts: steps kHi..kLo of the loop above on a carried state, so a caller that only needs the leading
// digits of the Hilbert position (a bucket of a coarse table) can stop early and finish the id only when it has to.
struct HilbertState { uint64_t n; uint32_t bits; };
S2B_HD HilbertState hilbert_begin(int face) { return HilbertState{(uint64_t)face << 60, (uint32_t)face & 1u}; }
template <int kHi, int kLo, class TableT>
S2B_HD HilbertState hilbert_steps(HilbertState st, uint32_t i, uint32_t j, const TableT* pos) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = kHi; k >= kLo; --k) {
    st.bits += ((i >> (4 * k)) & 15u) << 6;
    st.bits += ((j >> (4 * k)) & 15u) << 2;
    st.bits = pos[st.bits];
    st.n |= (uint64_t)(st.bits >> 2) << (8 * k);
    st.bits &= 3u;
  }
  return st;
}

// The projection half of S2CellId(S2Point): face and the UNTRUNCATED leaf coordinates fi = 2^30 * s, fj = 2^30 * t.
S2B_HD int point_to_face_fij(const P3& p, double* fi, double* fj) {
  double u, v;
  int face = xyz_to_face_uv(p, &u, &v);
  *fi = 1073741824.0 * uv_to_st(u);
  *fj = 1073741824.0 * uv_to_st(v);
  return face;
}
// STtoIJ applied to fi = 2^30 * s (s2coords.h:345-356): s > 0 <=> fi > 0; NaN -> 0.
S2B_HD uint32_t fij_to_ij(double f) {
#if defined(__CUDA_ARCH__)
  // cvt.rzi.s32.f64 saturates and maps NaN to 0, so "!(f > 0) -> 0" is max(., 0)
  int v = __double2int_rz(f);
  return (uint32_t)max(min(v, 1073741823), 0);
#else
  if (!(f > 0)) return 0;
  int v = (int)f;
  return (uint32_t)(v < 1073741823 ? v : 1073741823);
#endif
}

// FromFaceIJ with the 6-bit table (16384 entries: key iiiiiijjjjjjoo -> pppppppppppp oo): 5 dependent lookups instead
// of 8. The table is derived from the 4-bit one at start-up (s2b_hilbert.h MakeWideHilbert).
template <class TableT>
S2B_HD uint64_t from_face_ij_wide(int face, uint32_t i, uint32_t j, const TableT* pos6) {
  uint64_t n = (uint64_t)face << 60;
  uint32_t bits = (uint32_t)face & 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 4; k >= 0; --k) {
    bits += ((i >> (6 * k)) & 63u) << 8;
    bits += ((j >> (6 * k)) & 63u) << 2;
    bits = pos6[bits];
    n |= (uint64_t)(bits >> 2) << (12 * k);
    bits &= 3u;
  }
  return n * 2 + 1;
}

// The same encoding with the table laid out for the dependent chain of the encode kernel: entry = (pos12 << 3) |
// (orientation << 1), i.e. the orientation bits sit where the BYTE offset of the next lookup needs them, so one step is
// "offset = precomputed (i6, j6) offset | (entry & 6); entry = table[offset]" — one logic op + one load on the critical
// path instead of four. `tab` = the 16384 x 16-bit table as bytes (staged in shared memory by the kernel).
S2B_HD uint32_t hilbert_chain_entry(const unsigned char* tab, uint32_t byte_offset) {
  return *(const uint16_t*)(tab + byte_offset);
}
S2B_HD uint16_t hilbert_chain_repack(uint16_t e) { return (uint16_t)(((e >> 2) << 3) | ((e & 3u) << 1)); }   // from the pos6 layout
S2B_HD uint64_t from_face_ij_chain(int face, uint32_t i, uint32_t j, const unsigned char* tab) {
  // byte offsets of the five (i6, j6) pairs: (i6 << 9) | (j6 << 3), most significant first
  const uint32_t o4 = ((i >> 15) & 0x7E00u) | ((j >> 21) & 0x1F8u);
  const uint32_t o3 = ((i >> 9) & 0x7E00u) | ((j >> 15) & 0x1F8u);
  const uint32_t o2 = ((i >> 3) & 0x7E00u) | ((j >> 9) & 0x1F8u);
  const uint32_t o1 = ((i << 3) & 0x7E00u) | ((j >> 3) & 0x1F8u);
  const uint32_t o0 = ((i << 9) & 0x7E00u) | ((j << 3) & 0x1F8u);
  const uint32_t e4 = hilbert_chain_entry(tab, o4 | (((uint32_t)face & 1u) << 1));
  const uint32_t e3 = hilbert_chain_entry(tab, o3 | (e4 & 6u));
  const uint32_t e2 = hilbert_chain_entry(tab, o2 | (e3 & 6u));
  const uint32_t e1 = hilbert_chain_entry(tab, o1 | (e2 & 6u));
  const uint32_t e0 = hilbert_chain_entry(tab, o0 | (e1 & 6u));
  // id = face << 61 | pos4 << 49 | pos3 << 37 | pos2 << 25 | pos1 << 13 | pos0 << 1 | 1, with pos_k = e_k >> 3
  const uint32_t p2 = e2 >> 3;
  const uint32_t lo = 1u | ((e0 >> 3) << 1) | ((e1 >> 3) << 13) | (p2 << 25);
  const uint32_t hi = (p2 >> 7) | ((e3 >> 3) << 5) | ((e4 >> 3) << 17) | ((uint32_t)face << 29);
  return ((uint64_t)hi << 32) | lo;
}

// S2CellId(const S2Point&) (s2cell_id.cc:309-315).
template <class TableT>
S2B_HD uint64_t cellid_from_point(const P3& p, const TableT* pos) {
  double u, v;
  int face = xyz_to_face_uv(p, &u, &v);
  uint32_t i = st_to_ij(uv_to_st(u));
  uint32_t j = st_to_ij(uv_to_st(v));
  return from_face_ij(face, i, j, pos);
}

// S2CellId::ToFaceIJOrientation without the orientation (s2cell_id.cc:319-355). `ij` is the inverse table.
template <class TableT>
S2B_HD int cellid_to_face_ij(uint64_t id, const TableT* ij, int* pi, int* pj) {
  int i = 0, j = 0;
  int face = (int)(id >> 61);
  int bits = face & 1;
  for (int k = 7; k >= 0; --k) {
    const int nbits = (k == 7) ? 2 : 4;
    bits += ((int)(id >> (k * 8 + 1)) & ((1 << (2 * nbits)) - 1)) << 2;
    bits = ij[bits];
    i += (bits >> 6) << (k * 4);
    j += ((bits >> 2) & 15) << (k * 4);
    bits &= 3;
  }
  *pi = i; *pj = j;
  return face;
}

// S2::FaceUVtoXYZ (s2coords.h:368-382).
S2B_HD P3 face_uv_to_xyz(int face, double u, double v) {
  switch (face) {
    case 0: return P3{1, u, v};
    case 1: return P3{-u, 1, v};
    case 2: return P3{-u, -v, 1};
    case 3: return P3{-1, -v, -u};
    case 4: return P3{v, -1, -u};
    default: return P3{v, u, -1};
  }
}

// S2CellId::ToPoint: GetCenterSiTi (s2cell_id.h:555-581) -> FaceSiTitoXYZ (s2coords.cc:68-72) -> Normalize.
template <class TableT>
S2B_HD P3 cellid_to_point_raw(uint64_t id, const TableT* ij) {   // S2CellId::ToPointRaw: not normalised
  int i, j;
  int face = cellid_to_face_ij(id, ij, &i, &j);
  bool is_leaf = (id & 1) != 0;
  int delta = is_leaf ? 1 : (((i ^ ((int)(uint32_t)id >> 2)) & 1) ? 2 : 0);
  unsigned si = (unsigned)(2 * i + delta), ti = (unsigned)(2 * j + delta);
  double u = st_to_uv((1.0 / 2147483648.0) * si);
  double v = st_to_uv((1.0 / 2147483648.0) * ti);
  return face_uv_to_xyz(face, u, v);
}
template <class TableT>
S2B_HD P3 cellid_to_point(uint64_t id, const TableT* ij) { return normalize(cellid_to_point_raw(id, ij)); }

// ---------------------------------------------------------------------------------------------------
// Cell id algebra (s2cell_id.h:583-668).
S2B_HD uint64_t lsb_for_level(int level) { return 1ull << (2 * (30 - level)); }
S2B_HD uint64_t cellid_lsb(uint64_t id) { return id & (~id + 1); }
S2B_HD uint64_t cellid_parent(uint64_t id, int level) {
  uint64_t lsb = lsb_for_level(level);
  return (id & (~lsb + 1)) | lsb;
}
S2B_HD uint64_t cellid_range_min(uint64_t id) { return id - (cellid_lsb(id) - 1); }
S2B_HD uint64_t cellid_range_max(uint64_t id) { return id + (cellid_lsb(id) - 1); }
S2B_HD bool cellid_contains(uint64_t a, uint64_t b) { return b >= cellid_range_min(a) && b <= cellid_range_max(a); }

// S2CellId::ToToken (s2cell_id.cc:217-233): the 16-digit lowercase hex of the id without its trailing '0'
// digits ("X" for id 0), written as a fixed 16-byte slot (four little-endian 32-bit words, NUL padded). Returns strlen.
// Works on 16-bit chunks in 32-bit registers (64-bit SWAR costs twice the instructions on the GPU).
S2B_HD uint32_t ascii_hex4(uint32_t v) {  // 4 bytes holding 0..15 each -> their lowercase hex digits
  uint32_t ge10 = ((v + 0x06060606u) >> 4) & 0x01010101u;
  return v + 0x30303030u + ge10 * 39u;                 // '0'.. or 'a'.. ('a' - '0' - 10 = 39)
}
S2B_HD uint32_t hex4(uint32_t x) {  // 4 nibbles (16 bits) -> 4 ASCII bytes, most significant nibble in the lowest byte
  uint32_t v = (x | (x << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;                    // byte b = nibble b
  v = (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);   // reverse the bytes
  return ascii_hex4(v);
}
// The 8 hex digits of a 32-bit word as two ASCII words (digits 7..4 and 3..0, most significant digit in the lowest
// byte). On the device one byte permute places each source byte twice, in output order, and two logic ops split it into
// nibbles: 9 instructions per word.
S2B_HD void hex8(uint32_t x, uint32_t* hi4, uint32_t* lo4) {
#if defined(__CUDA_ARCH__)
  const uint32_t th = __byte_perm(x, 0, 0x2233), tl = __byte_perm(x, 0, 0x0011);   // [B3 B3 B2 B2] / [B1 B1 B0 B0], low byte first
  *hi4 = ascii_hex4(((th >> 4) & 0x000F000Fu) | (th & 0x0F000F00u));
  *lo4 = ascii_hex4(((tl >> 4) & 0x000F000Fu) | (tl & 0x0F000F00u));
#else
  *hi4 = hex4(x >> 16);
  *lo4 = hex4(x & 0xFFFFu);
#endif
}
S2B_HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
S2B_HD int cellid_to_token(uint64_t id, uint32_t out[4]) {
  if (id == 0) { out[0] = 0x58; out[1] = out[2] = out[3] = 0; return 1; }  // "X"
  const int len = 16 - (ctz64(id) >> 2);
  const uint32_t hi = (uint32_t)(id >> 32), lo = (uint32_t)id;
  uint32_t w[4];
  hex8(hi, &w[0], &w[1]);
  hex8(lo, &w[2], &w[3]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    int keep = len - 4 * k;                            // characters of word k that belong to the token
    out[k] = keep >= 4 ? w[k] : (keep <= 0 ? 0u : (w[k] & ((1u << (8 * keep)) - 1u)));
  }
  return len;
}

// S2CellId::level() for a valid id (s2cell_id.h:595-603).
S2B_HD int cellid_level(uint64_t id) { return 30 - (ctz64(id) >> 1); }
S2B_HD int bit_width64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return 64 - __clzll((long long)v);
#else
  return v ? 64 - __builtin_clzll(v) : 0;
#endif
}

// ---- traversal helpers of S2CellId (pure integer) -------------------------------------------------------------------
S2B_HD uint64_t cellid_child(uint64_t id, int position) {            // s2cell_id.h:670-677
  const uint64_t new_lsb = cellid_lsb(id) >> 2;
  return id + (uint64_t)(2 * position + 1 - 4) * new_lsb;
}
constexpr uint64_t kCellIdWrapOffset = 6ull << 61;                     // S2CellId::kWrapOffset (s2cell_id.h:525)
// S2CellId::advance (s2cell_id.cc:119-138; clamped to [Begin(level), End(level)]) and ::advance_wrap (:145-168).
S2B_HD uint64_t cellid_advance(uint64_t id, long long steps, bool wrap) {
  if (steps == 0) return id;
  const int step_shift = 2 * (30 - cellid_level(id)) + 1;
  if (steps < 0) {
    const long long min_steps = -(long long)(id >> step_shift);
    if (steps < min_steps) {
      if (!wrap) {
        steps = min_steps;
      } else {
        const long long step_wrap = (long long)(kCellIdWrapOffset >> step_shift);
        steps %= step_wrap;
        if (steps < min_steps) steps += step_wrap;
      }
    }
  } else {
    const long long max_steps = (long long)((wrap ? kCellIdWrapOffset - id : kCellIdWrapOffset + cellid_lsb(id) - id) >> step_shift);
    if (steps > max_steps) {
      if (!wrap) {
        steps = max_steps;
      } else {
        const long long step_wrap = (long long)(kCellIdWrapOffset >> step_shift);
        steps %= step_wrap;
        if (steps > max_steps) steps -= step_wrap;
      }
    }
  }
  return id + ((uint64_t)steps << step_shift);
}
// S2CellId::GetCommonAncestorLevel (s2cell_id.cc:193-207): -1 when the cells are on different faces.
S2B_HD int cellid_common_ancestor_level(uint64_t a, uint64_t b) {
  uint64_t bits = a ^ b;
  const uint64_t la = cellid_lsb(a), lb = cellid_lsb(b);
  if (la > bits) bits = la;
  if (lb > bits) bits = lb;
  const int width = bit_width64(bits);                                 // absl::bit_width: position of the top bit + 1
  const int v = 61 - width;
  return (v > -1 ? v : -1) >> 1;
}

// S2CellId::FromFaceIJWrap (s2cell_id.cc:458-489): (i,j) just outside a face -> the leaf cell on the adjacent face.
template <class TableT>
S2B_HD uint64_t from_face_ij_wrap(int face, int i, int j, const TableT* pos) {
  const int kMaxSize = 1 << 30;
  i = i < -1 ? -1 : (i > kMaxSize ? kMaxSize : i);
  j = j < -1 ? -1 : (j > kMaxSize ? kMaxSize : j);
  const double kScale = 1.0 / kMaxSize;
  const double kLimit = 1.0 + 2.2204460492503131e-16;
  double u = kScale * (2 * (i - kMaxSize / 2) + 1);
  double v = kScale * (2 * (j - kMaxSize / 2) + 1);
  u = u < -kLimit ? -kLimit : (u > kLimit ? kLimit : u);
  v = v < -kLimit ? -kLimit : (v > kLimit ? kLimit : v);
  const P3 p = face_uv_to_xyz(face, u, v);
  const int f2 = xyz_to_face_uv(p, &u, &v);
  return from_face_ij(f2, st_to_ij(0.5 * (u + 1)), st_to_ij(0.5 * (v + 1)), pos);
}

// S2CellId::GetEdgeNeighbors (s2cell_id.cc:499-512): down, right, up, left neighbours at the cell's own level.
template <class TableT>
S2B_HD void cellid_edge_neighbors(uint64_t id, const TableT* pos, const TableT* ij, uint64_t out[4]) {
  const int kMaxSize = 1 << 30;
  int i, j;
  const int level = cellid_level(id);
  const int size = 1 << (30 - level);
  const int face = cellid_to_face_ij(id, ij, &i, &j);
  const int ni[4] = {i, i + size, i, i - size};
  const int nj[4] = {j - size, j, j + size, j};
  const bool same[4] = {j - size >= 0, i + size < kMaxSize, j + size < kMaxSize, i - size >= 0};
  for (int k = 0; k < 4; ++k) {
    uint64_t leaf = same[k] ? from_face_ij(face, (uint32_t)ni[k], (uint32_t)nj[k], pos) : from_face_ij_wrap(face, ni[k], nj[k], pos);
    out[k] = cellid_parent(leaf, level);
  }
}

// S2CellId::FromToken (s2cell_id.cc:235-254) from a 16-byte slot + length: hex digits (either case) placed from the
// top nibble down; any other character, or a length above 16, gives S2CellId::None() (0). "X" therefore decodes to 0.
S2B_HD uint64_t cellid_from_token(const unsigned char* tok, int len) {
  if (len > 16 || len < 0) return 0;
  uint64_t id = 0;
  for (int i = 0; i < len; ++i) {
    const unsigned c = tok[i];
    uint64_t d;
    if (c - '0' <= 9u) d = c - '0';
    else if (c - 'a' <= 5u) d = c - 'a' + 10;
    else if (c - 'A' <= 5u) d = c - 'A' + 10;
    else return 0;
    id |= d << (60 - 4 * i);
  }
  return id;
}

}  // namespace s2b
