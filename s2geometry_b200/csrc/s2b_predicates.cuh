// s2b_predicates.cuh — robust orientation and edge-crossing predicates for device code (also host-compilable for
// tests/hostsim). Everything here must return exactly what the reference returns for the same bits:
//
//   TriageSign           s2predicates.h:342-403      filtered (a x b) . c
//   StableSign           s2predicates.cc:64-104      filtered determinant from the two shorter edge vectors
//   ExactSign            s2predicates.cc:226-262     exact determinant sign of the lexicographically sorted points
//   SymbolicallyPerturbedSign  s2predicates.cc:131-222
//   ExpensiveSign / Sign s2predicates.cc:275-297, s2predicates.h:330-335
//   OrderedCCW           s2predicates.cc:299-312
//   S2::Ortho / RefDir   s2pointutil.cc:48-59, s2pointutil.h:119-121
//   S2::VertexCrossing   s2edge_crossings.cc:371-391
//   CrossingSign         s2edge_crosser.h:365-388 + s2edge_crosser.cc:40-96 as a pure function
//
// The reference evaluates the exact stage with ExactFloat (arbitrary-precision binary floating point). Here the
// exact stage is an integer superaccumulator: a sum of signed products of up to three doubles is accumulated
// exactly in a 6.7-kbit two's-complement fixed-point register (every product of three doubles, from 2^-3222 to
// 2^3072, lands inside it), and only its sign is read. No expansion arithmetic, so no underflow hazard, no CPU
// fallback, and the reference's ">2000-bit" test cases (s2edge_crosser_test.cc:226-248) are ordinary inputs.
#pragma once
#include "s2b_core.cuh"

namespace s2b {

// Statistic only: how many predicate evaluations reached the exact stage (s2b_exact_stage_count in the C ABI).
#if defined(__CUDACC__)
// One copy per translation unit that includes this header (s2b_index.cu: point queries; s2b_region.cu: region cell
// tests); s2b_exact_stage_count sums them.
static __device__ unsigned long long g_exact_count = 0;
#endif
#if defined(__CUDA_ARCH__)
#define S2B_NOTE_EXACT() atomicAdd(&::s2b::g_exact_count, 1ull)
#else
inline unsigned long long& host_exact_count() { static unsigned long long c = 0; return c; }
#define S2B_NOTE_EXACT() (++::s2b::host_exact_count())
#endif

#ifndef DBL_EPSILON
#define S2B_DBL_EPSILON 2.2204460492503131e-16
#else
#define S2B_DBL_EPSILON DBL_EPSILON
#endif

// ---------------------------------------------------------------------------------------------------
// Filters (literal restatements: same operations, same order, same constants).

S2B_HD int triage_sign(const P3& a_cross_b, const P3& c) {
  const double kMaxDetError = 3.6548 * S2B_DBL_EPSILON;
  double det = dot(a_cross_b, c);
  if (det > kMaxDetError) return 1;
  if (det < -kMaxDetError) return -1;
  return 0;
}

S2B_HD int stable_sign(const P3& a, const P3& b, const P3& c) {
  P3 ab = sub(b, a), bc = sub(c, b), ca = sub(a, c);
  double ab2 = norm2(ab), bc2 = norm2(bc), ca2 = norm2(ca);
  const double kDetErrorMultiplier = 3.2321 * S2B_DBL_EPSILON;
  double det, max_error;
  if (ab2 >= bc2 && ab2 >= ca2) {
    det = -dot(cross(ca, bc), c);
    max_error = kDetErrorMultiplier * sqrt(ca2 * bc2);
  } else if (bc2 >= ca2) {
    det = -dot(cross(ab, ca), a);
    max_error = kDetErrorMultiplier * sqrt(ab2 * ca2);
  } else {
    det = -dot(cross(bc, ab), b);
    max_error = kDetErrorMultiplier * sqrt(bc2 * ab2);
  }
  const double kMinNoUnderflowError = kDetErrorMultiplier * 1.4916681462400413e-154;  // sqrt(DBL_MIN)
  if (max_error < kMinNoUnderflowError) return 0;
  return (fabs(det) <= max_error) ? 0 : (det > 0) ? 1 : -1;
}

// ---------------------------------------------------------------------------------------------------
// Exact stage: superaccumulator.

// value = (-1)^neg * m * 2^e with integer m < 2^53 (m == 0 for zeros; NaN/inf never reach the exact stage with a
// meaningful answer in the reference either — ExactFloat would hold NaN — and are mapped to 0 here).
struct Dec { uint64_t m; int e; bool neg; };

S2B_HD Dec decompose(double d) {
  union { double f; uint64_t u; } x;
  x.f = d;
  Dec r;
  r.neg = (x.u >> 63) != 0;
  int ef = (int)((x.u >> 52) & 0x7FF);
  uint64_t frac = x.u & 0x000FFFFFFFFFFFFFull;
  if (ef == 0) { r.m = frac; r.e = -1074; }
  else if (ef == 0x7FF) { r.m = 0; r.e = 0; }
  else { r.m = frac | 0x0010000000000000ull; r.e = ef - 1075; }
  return r;
}

S2B_HD void mul64wide(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}

// 104 limbs x 64 bits: bit 0 has weight 2^-3222 (three subnormal factors); the largest term, < 2^159 * 2^2913, ends
// below bit 6294; the top limbs absorb carries and hold the sign.
constexpr int kAccLimbs = 104;
constexpr int kAccBias = 3222;

struct ExactAcc {
  uint64_t w[kAccLimbs];
  S2B_HD void clear() { for (int i = 0; i < kAccLimbs; ++i) w[i] = 0; }
  // acc += sign * x*y*z (sign = +1/-1), exactly.
  S2B_HD void add_product(double x, double y, double z, int sign) {
    Dec a = decompose(x), b = decompose(y), c = decompose(z);
    if (a.m == 0 || b.m == 0 || c.m == 0) return;
    bool neg = (a.neg != b.neg) != c.neg;
    if (sign < 0) neg = !neg;
    uint64_t p_hi, p_lo;
    mul64wide(a.m, b.m, p_hi, p_lo);           // < 2^106
    uint64_t t0, t1, t2, c1, c2;
    mul64wide(p_lo, c.m, c1, t0);              // low limb and carry
    mul64wide(p_hi, c.m, c2, t1);              // p_hi < 2^42, c.m < 2^53: c2 < 2^31
    t1 += c1;
    t2 = c2 + (t1 < c1 ? 1 : 0);               // magnitude = t2:t1:t0 < 2^159
    int off = a.e + b.e + c.e + kAccBias;      // >= 0
    int limb = off >> 6, sh = off & 63;
    uint64_t v[4];
    if (sh == 0) { v[0] = t0; v[1] = t1; v[2] = t2; v[3] = 0; }
    else {
      v[0] = t0 << sh;
      v[1] = (t1 << sh) | (t0 >> (64 - sh));
      v[2] = (t2 << sh) | (t1 >> (64 - sh));
      v[3] = t2 >> (64 - sh);
    }
    if (!neg) {
      uint64_t carry = 0;
      for (int i = limb; i < kAccLimbs; ++i) {
        uint64_t add = (i - limb < 4) ? v[i - limb] : 0;
        if (i - limb >= 4 && carry == 0) break;
        uint64_t s = w[i] + add;
        uint64_t c_out = s < add ? 1 : 0;
        uint64_t s2 = s + carry;
        c_out |= (s2 < s) ? 1 : 0;
        w[i] = s2;
        carry = c_out;
      }
    } else {
      uint64_t borrow = 0;
      for (int i = limb; i < kAccLimbs; ++i) {
        uint64_t subv = (i - limb < 4) ? v[i - limb] : 0;
        if (i - limb >= 4 && borrow == 0) break;
        uint64_t d = w[i] - subv;
        uint64_t b_out = w[i] < subv ? 1 : 0;
        uint64_t d2 = d - borrow;
        b_out |= (d < borrow) ? 1 : 0;
        w[i] = d2;
        borrow = b_out;
      }
    }
  }
  S2B_HD int sign() const {
    if (w[kAccLimbs - 1] >> 63) return -1;
    for (int i = kAccLimbs - 1; i >= 0; --i) if (w[i] != 0) return 1;
    return 0;
  }
  // The exact value as ExactFloat::ToDouble would report it (util/math/exactfloat/exactfloat.cc:131-161): the magnitude
  // rounded to 53 bits, ties to even, as mant * 2^exp2 (mant < 2^53 after renormalisation), plus ExactFloat::exp() of the
  // UNROUNDED value (|v| in [2^(top-1), 2^top)). Destroys the register when the value is negative (it is negated in place).
  struct Read { int sign; uint64_t mant; int exp2; int top; };
  S2B_HD Read read() {
    Read r{0, 0, 0, 0};
    r.sign = sign();
    if (r.sign == 0) return r;
    if (r.sign < 0) {   // two's complement negate
      uint64_t carry = 1;
      for (int i = 0; i < kAccLimbs; ++i) { const uint64_t v = ~w[i] + carry; carry = (carry && v == 0) ? 1 : 0; w[i] = v; }
    }
    int hi = kAccLimbs - 1;
    while (w[hi] == 0) --hi;
    int lz = 0;
    for (uint64_t t = w[hi]; !(t >> 63); t <<= 1) ++lz;
    const int top_bit = hi * 64 + 63 - lz;                 // index of the leading one; weight 2^(top_bit - kAccBias)
    r.top = top_bit - kAccBias + 1;
    // the 64 bits starting at the leading one, and whether anything non-zero lies below them
    auto bit_window = [&](int from_bit) -> uint64_t {      // bits [from_bit, from_bit + 64) of the register (may start below 0)
      if (from_bit <= -64) return 0;
      if (from_bit < 0) return w[0] << (-from_bit);
      const int l = from_bit >> 6, sft = from_bit & 63;
      uint64_t v = w[l] >> sft;
      if (sft && l + 1 < kAccLimbs) v |= w[l + 1] << (64 - sft);
      return v;
    };
    const int lsb = top_bit - 63;                          // register bit that lands in bit 0 of `top64`
    const uint64_t top64 = bit_window(lsb);
    bool sticky = false;
    if (lsb > 0) {
      const int l = lsb >> 6, sft = lsb & 63;
      if (sft && (w[l] << (64 - sft)) != 0) sticky = true;
      for (int i = l - 1; i >= 0 && !sticky; --i) sticky = w[i] != 0;
    }
    uint64_t mant = top64 >> 11;
    const uint64_t rest = top64 & 0x7FF;
    int exp2 = (top_bit - 52) - kAccBias;
    if (rest > 0x400 || (rest == 0x400 && (sticky || (mant & 1)))) ++mant;
    if (mant >> 53) { mant >>= 1; ++exp2; }
    r.mant = mant; r.exp2 = exp2;
    return r;
  }
};

// sign(p*q - r*s), exact.
S2B_HD_NOINLINE int exact_minor_sign(double p, double q, double r, double s) {
  ExactAcc acc;
  acc.clear();
  acc.add_product(p, q, 1.0, 1);
  acc.add_product(r, s, 1.0, -1);
  return acc.sign();
}

S2B_HD int dsgn(double v) { return v > 0 ? 1 : (v < 0 ? -1 : 0); }

// sign of det [a; b; c] = a . (b x c), exact.
S2B_HD_NOINLINE int exact_det_sign(const P3& a, const P3& b, const P3& c) {
  ExactAcc acc;
  acc.clear();
  acc.add_product(a.x, b.y, c.z, 1);
  acc.add_product(a.x, b.z, c.y, -1);
  acc.add_product(a.y, b.z, c.x, 1);
  acc.add_product(a.y, b.x, c.z, -1);
  acc.add_product(a.z, b.x, c.y, 1);
  acc.add_product(a.z, b.y, c.x, -1);
  return acc.sign();
}

// SymbolicallyPerturbedSign (s2predicates.cc:131-222) for a < b < c with det == 0: the first non-zero of thirteen
// signs of minors/entries (the order is the reference's).
S2B_HD int symbolic_sign(const P3& a, const P3& b, const P3& c) {
  int s;
  s = exact_minor_sign(b.x, c.y, b.y, c.x); if (s) return s;   // (b x c)[2]
  s = exact_minor_sign(b.z, c.x, b.x, c.z); if (s) return s;   // (b x c)[1]
  s = exact_minor_sign(b.y, c.z, b.z, c.y); if (s) return s;   // (b x c)[0]
  s = exact_minor_sign(c.x, a.y, c.y, a.x); if (s) return s;   // c0*a1 - c1*a0
  s = dsgn(c.x); if (s) return s;
  s = -dsgn(c.y); if (s) return s;
  s = exact_minor_sign(c.z, a.x, c.x, a.z); if (s) return s;   // c2*a0 - c0*a2
  s = dsgn(c.z); if (s) return s;
  s = exact_minor_sign(a.x, b.y, a.y, b.x); if (s) return s;   // a0*b1 - a1*b0
  s = -dsgn(b.x); if (s) return s;
  s = dsgn(b.y); if (s) return s;
  s = dsgn(a.x); if (s) return s;
  return 1;
}

// ExactSign (s2predicates.cc:226-262).
S2B_HD_NOINLINE int exact_sign(const P3& a, const P3& b, const P3& c, bool perturb) {
  int perm_sign = 1;
  P3 pa = a, pb = b, pc = c, t;
  if (lex_less(pb, pa)) { t = pa; pa = pb; pb = t; perm_sign = -perm_sign; }
  if (lex_less(pc, pb)) { t = pb; pb = pc; pc = t; perm_sign = -perm_sign; }
  if (lex_less(pb, pa)) { t = pa; pa = pb; pb = t; perm_sign = -perm_sign; }
  int det_sign = exact_det_sign(pa, pb, pc);
  if (det_sign == 0 && perturb) det_sign = symbolic_sign(pa, pb, pc);
  return perm_sign * det_sign;
}

// ExpensiveSign (s2predicates.cc:275-297).
S2B_HD int expensive_sign(const P3& a, const P3& b, const P3& c) {
  if (eq(a, b) || eq(b, c) || eq(c, a)) return 0;
  int s = stable_sign(a, b, c);
  if (s != 0) return s;
  S2B_NOTE_EXACT();
  return exact_sign(a, b, c, true);
}

// Sign(a, b, c, a x b) (s2predicates.h:330-335) and Sign(a, b, c) (s2predicates.cc:43-48).
S2B_HD int sign_with_cross(const P3& a, const P3& b, const P3& c, const P3& a_cross_b) {
  int s = triage_sign(a_cross_b, c);
  if (s == 0) s = expensive_sign(a, b, c);
  return s;
}
S2B_HD int sign3(const P3& a, const P3& b, const P3& c) {
  return sign_with_cross(a, b, c, cross(a, b));
}

// OrderedCCW (s2predicates.cc:299-312).
S2B_HD bool ordered_ccw(const P3& a, const P3& b, const P3& c, const P3& o) {
  int sum = 0;
  if (sign3(b, o, a) >= 0) ++sum;
  if (sign3(c, o, b) >= 0) ++sum;
  if (sign3(a, o, c) > 0) ++sum;
  return sum >= 2;
}

// S2::Ortho (s2pointutil.cc:48-59) = S2::RefDir.
S2B_HD P3 ortho(const P3& a) {
  int k = largest_abs_component(a) - 1;
  if (k < 0) k = 2;
  P3 t{0.012, 0.0053, 0.00457};
  if (k == 0) t.x = 1; else if (k == 1) t.y = 1; else t.z = 1;
  return normalize(cross(a, t));
}

// S2::VertexCrossing (s2edge_crossings.cc:371-391).
S2B_HD bool vertex_crossing(const P3& a, const P3& b, const P3& c, const P3& d) {
  if (eq(a, b) || eq(c, d)) return false;
  if (eq(a, c)) return eq(b, d) || ordered_ccw(ortho(a), d, b, a);
  if (eq(b, d)) return ordered_ccw(ortho(b), c, a, b);
  if (eq(a, d)) return eq(b, c) || ordered_ccw(ortho(a), c, b, a);
  if (eq(b, c)) return ordered_ccw(ortho(b), d, a, b);
  return false;
}

// The slow half of S2EdgeCrosser::CrossingSign (s2edge_crosser.cc:40-96), entered when the triage signs of C and D
// against AB do not prove "same side". The tangent-plane early-out (:53-71) is omitted: it can only return -1 in
// configurations where the exact tests below also return -1.
S2B_HD int crossing_sign_slow(const P3& a, const P3& b, const P3& c, const P3& d, int acb, int bda) {
  if (eq(a, c) || eq(a, d) || eq(b, c) || eq(b, d)) return 0;
  if (eq(a, b) || eq(c, d)) return -1;
  if (acb == 0) acb = -expensive_sign(a, b, c);
  if (bda == 0) bda = expensive_sign(a, b, d);
  if (bda != acb) return -1;
  P3 c_cross_d = cross(c, d);
  int cbd = -sign_with_cross(c, d, b, c_cross_d);
  if (cbd != acb) return -1;
  int dac = sign_with_cross(c, d, a, c_cross_d);
  return (dac != acb) ? -1 : 1;
}

// S2::CrossingSign(a, b, c, d) (s2edge_crossings.cc:365-369) as a pure function.
S2B_HD int crossing_sign(const P3& a, const P3& b, const P3& c, const P3& d) {
  P3 axb = cross(a, b);
  int acb = -triage_sign(axb, c);
  int bda = triage_sign(axb, d);
  if (acb == -bda && bda != 0) return -1;
  return crossing_sign_slow(a, b, c, d, acb, bda);
}

}  // namespace s2b
