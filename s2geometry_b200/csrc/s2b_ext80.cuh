// s2b_ext80.cuh — x87 extended precision (64-bit significand, round to nearest even) in integer arithmetic, for the ONE
// place the hot path's callers depend on it: the `long double` stage of S2::RobustCrossProd (s2edge_crossings.cc:167-173,
// GetStableCrossProd<long double>), which the reference runs on x86-64 for edges whose endpoints are closer than about
// 9 * DBL_EPSILON (or nearly antipodal). To return the bits the reference returns there, (a - b) x (a + b), its squared
// norm and the final conversion to double are evaluated here operation by operation with the x87's rounding. Exponent
// range is not modelled (|values| between 2^-1100 and 4 cannot leave the 15-bit exponent range); zero is m == 0.
#pragma once
#include "s2b_core.cuh"

namespace s2b {

struct X80 { uint64_t m; int e; bool neg; };   // value = (-1)^neg * m * 2^(e - 63), m has bit 63 set (or is 0)

S2B_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}
S2B_HD void x80_mul64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
  *lo = a * b;
  *hi = __umul64hi(a, b);
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  *lo = (uint64_t)p;
  *hi = (uint64_t)(p >> 64);
#endif
}

S2B_HD X80 x80_from_double(double d) {   // exact
  uint64_t bits;
#if defined(__CUDA_ARCH__)
  bits = (uint64_t)__double_as_longlong(d);
#else
  __builtin_memcpy(&bits, &d, 8);
#endif
  X80 r;
  r.neg = (bits >> 63) != 0;
  const int ex = (int)((bits >> 52) & 0x7FF);
  uint64_t frac = bits & 0xFFFFFFFFFFFFFull;
  if (ex == 0) {
    if (frac == 0) { r.m = 0; r.e = 0; return r; }
    const int lz = clz64(frac);           // subnormal: normalise
    r.m = frac << lz;
    r.e = -1022 - 52 + (63 - lz);
    return r;
  }
  r.m = (frac | (1ull << 52)) << 11;
  r.e = ex - 1023;
  return r;
}

// Rounds the 128-bit magnitude hi:lo (hi has bit 63 set), whose bit 127 has weight 2^e, to 64 bits, ties to even.
S2B_HD X80 x80_round(uint64_t hi, uint64_t lo, int e, bool neg) {
  const bool round = (lo >> 63) != 0, sticky = (lo << 1) != 0;
  if (round && (sticky || (hi & 1))) {
    if (++hi == 0) { hi = 1ull << 63; ++e; }
  }
  return X80{hi, e, neg};
}

S2B_HD X80 x80_mul(const X80& a, const X80& b) {
  const bool neg = a.neg != b.neg;
  if (a.m == 0 || b.m == 0) return X80{0, 0, neg};
  uint64_t hi, lo;
  x80_mul64(a.m, b.m, &hi, &lo);            // in [2^126, 2^128)
  int e = a.e + b.e + 1;                    // weight of bit 127
  if (!(hi >> 63)) { hi = (hi << 1) | (lo >> 63); lo <<= 1; --e; }
  return x80_round(hi, lo, e, neg);
}

// a + b (b's sign flipped when `subtract`), rounded to 64 bits.
S2B_HD X80 x80_add(const X80& a, X80 b, bool subtract = false) {
  if (subtract) b.neg = !b.neg;
  if (b.m == 0) return a.m == 0 ? X80{0, 0, a.neg && b.neg} : a;
  if (a.m == 0) return b;
  // x = the operand of larger magnitude
  const bool a_big = a.e > b.e || (a.e == b.e && a.m >= b.m);
  const X80& x = a_big ? a : b;
  const X80& y = a_big ? b : a;
  const int d = x.e - y.e;
  // 128-bit frame with x.m in bits 126..63 (one bit of headroom for the carry); y shifted right by d, sticky jammed in
  uint64_t xh = x.m >> 1, xl = x.m << 63;
  uint64_t yh, yl;
  if (d >= 127) { yh = 0; yl = 1; }
  else {
    const int s = d + 1;
    if (s < 64) { yh = y.m >> s; yl = y.m << (64 - s); }
    else { yh = 0; yl = y.m >> (s - 64); if (s > 64 && (y.m << (128 - s)) != 0) yl |= 1; }
  }
  uint64_t rh, rl;
  if (x.neg == y.neg) {
    rl = xl + yl;
    rh = xh + yh + (rl < xl ? 1 : 0);
  } else {
    rl = xl - yl;
    rh = xh - yh - (xl < yl ? 1 : 0);
    if ((rh | rl) == 0) return X80{0, 0, false};   // exact cancellation: +0 under round to nearest
  }
  // normalise so that bit 127 is set; weight of bit 127 before normalisation is 2^(x.e + 1)
  int e = x.e + 1;
  int lz = rh ? clz64(rh) : 64 + clz64(rl);
  if (lz >= 64) { rh = rl << (lz - 64); rl = 0; }
  else if (lz > 0) { rh = (rh << lz) | (rl >> (64 - lz)); rl <<= lz; }
  e -= lz;
  return x80_round(rh, rl, e, x.neg);
}

S2B_HD bool x80_ge(const X80& a, const X80& b) {   // a >= b for non-negative a, b
  if (b.m == 0) return true;
  if (a.m == 0) return false;
  return a.e > b.e || (a.e == b.e && a.m >= b.m);
}

S2B_HD double x80_to_double(const X80& a) {   // static_cast<double>(long double): round to nearest even
  if (a.m == 0) return a.neg ? -0.0 : 0.0;
  uint64_t m = a.m >> 11;
  const uint64_t rest = a.m & 0x7FF;
  int e = a.e - 52;                              // value = m * 2^e (before rounding)
  if (rest > 0x400 || (rest == 0x400 && (m & 1))) ++m;   // m == 2^53 is still exactly representable
  const double v = ldexp((double)m, e);
  return a.neg ? -v : v;
}

}  // namespace s2b
