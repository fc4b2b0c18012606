// s2b_synth.cuh — counter-based synthetic inputs: every value is a pure function of (seed, global
// index), built from integer hashing and + - * / sqrt in IEEE double only (no libm, no FMA), so the device kernel below,
// its host compilation (tests/hostsim) and the NumPy restatement in s2geometry_b200/synth.py produce the same bits.
// That is what lets bench.py give every rank its slice of ONE global point stream: the union over ranks, and therefore
// every histogram and hit count, is identical at 1, 2, 4 and 8 GPUs and can be checked against the reference.
#pragma once
#include "s2b_core.cuh"

namespace s2b {

S2B_HD uint64_t synth_mix(uint64_t z) {   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// The k-th pair of 64-bit draws of point `index`.
S2B_HD uint64_t synth_state(uint64_t seed, uint64_t index, uint32_t attempt) {
  return synth_mix(seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull) ^ (index * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)attempt * 0xD1B54A32D192ED03ull);
}
S2B_HD double synth_unit(uint64_t r) {     // (0, 1), 52 random bits, exact
  return ((double)(r >> 12) + 0.5) * 0x1p-52;
}
S2B_HD double synth_signed(uint64_t r) {   // (-1, 1), exact: (2k + 1) * 2^-52 - 1
  return ((double)(r >> 12) + 0.5) * 0x1p-51 - 1.0;
}

// Uniform on the sphere (Marsaglia 1972): (x1, x2) uniform in the unit disk by rejection, s = x1^2 + x2^2,
// p = (2 x1 sqrt(1 - s), 2 x2 sqrt(1 - s), 1 - 2 s).
S2B_HD P3 synth_sphere_point(uint64_t seed, uint64_t index) {
  for (uint32_t k = 0;; ++k) {
    const uint64_t st = synth_state(seed, index, k);
    const double x1 = synth_signed(synth_mix(st)), x2 = synth_signed(synth_mix(st + 0x9E3779B97F4A7C15ull));
    const double s = x1 * x1 + x2 * x2;
    if (s < 1.0 || k == 63) {
      const double ss = s < 1.0 ? s : 0.5;
      const double r = sqrt(1.0 - ss);
      return P3{(2.0 * x1) * r, (2.0 * x2) * r, 1.0 - 2.0 * ss};
    }
  }
}

// asin restated with + - * / sqrt only (rational approximation of (asin(x) - x) / x^3 on [0, 0.5], the classic
// coefficients; reflection asin(x) = pi/2 - 2 asin(sqrt((1 - x) / 2)) above): within 3e-16 of asin, which is all a
// synthetic latitude needs — the point is that host and device agree on the bits.
S2B_HD double synth_asin(double z) {
  const double a = fabs(z);
  const bool small = a < 0.5;
  const double w = small ? z * z : (1.0 - a) * 0.5;
  const double p = w * (1.66666666666666657415e-01 + w * (-3.25565818622400915405e-01 + w * (2.01212532134862925881e-01 +
                   w * (-4.00555345006794114027e-02 + w * (7.91534994289814532176e-04 + w * 3.47933107596021167570e-05)))));
  const double q = 1.0 + w * (-2.40339491173441421878e+00 + w * (2.02094576023350569471e+00 + w * (-6.88283971605453293030e-01 +
                   w * 7.70381505559019352791e-02)));
  const double r = p / q;
  if (small) return z + z * r;
  const double s = sqrt(w);
  const double big = 1.57079632679489655800e+00 - 2.0 * (s + s * r);
  return z < 0 ? -big : big;
}

// (lat, lng) radians uniform on the sphere: lat = asin(U(-1, 1)), lng = U(-1, 1) * pi.
S2B_HD void synth_sphere_latlng(uint64_t seed, uint64_t index, double* lat, double* lng) {
  const uint64_t st = synth_state(seed, index, 0);
  *lat = synth_asin(synth_signed(synth_mix(st)));
  *lng = synth_signed(synth_mix(st + 0x9E3779B97F4A7C15ull)) * 3.14159265358979323846;
}

// Uniform in the spherical cap of height `height` (= 1 - cos(radius)) around frame[2]; frame = orthonormal (x, y, z) rows.
// Direction in the tangent plane from a rejection-sampled point of the unit disk (no sin / cos).
S2B_HD P3 synth_cap_point(uint64_t seed, uint64_t index, const double* frame, double height) {
  const uint64_t s0 = synth_state(seed, index, 0);
  const double cz = 1.0 - synth_unit(synth_mix(s0)) * height;
  const double t = 1.0 - cz * cz;
  const double sz = sqrt(t > 0 ? t : 0.0);
  double a = 1.0, b = 0.0;
  for (uint32_t k = 1; k < 64; ++k) {
    const uint64_t st = synth_state(seed, index, k);
    const double x1 = synth_signed(synth_mix(st)), x2 = synth_signed(synth_mix(st + 0x9E3779B97F4A7C15ull));
    const double s = x1 * x1 + x2 * x2;
    if (s < 1.0 && s > 0x1p-40) {
      const double inv = 1.0 / sqrt(s);
      a = x1 * inv; b = x2 * inv;
      break;
    }
  }
  P3 p;
  p.x = cz * frame[6] + sz * (a * frame[0] + b * frame[3]);
  p.y = cz * frame[7] + sz * (a * frame[1] + b * frame[4]);
  p.z = cz * frame[8] + sz * (a * frame[2] + b * frame[5]);
  const double inv = 1.0 / sqrt((p.x * p.x + p.y * p.y) + p.z * p.z);
  return P3{p.x * inv, p.y * inv, p.z * inv};
}

}  // namespace s2b
