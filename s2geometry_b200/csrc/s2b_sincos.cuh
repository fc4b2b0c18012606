// s2b_sincos.cuh — correctly rounded sin/cos for the lat/lng front end of the encoder
// (S2LatLng::ToPoint, s2latlng.cc:68-77, calls libm sin and cos separately: s1angle.h:343-357).
//
// Why not the CUDA math library: its sin/cos are 1-2 ulp, and a 1-ulp slip in x/y/z moves a level-30 cell
// with p ≈ 2.4e-7 per coordinate. The reference's own results depend on the platform libm (glibc is < 1 ulp but
// NOT correctly rounded for ~0.13% of arguments), so the only platform-independent definition is the correctly
// rounded one; that is what this file computes. Ziv's two-step strategy:
//   1. Cody-Waite reduction by pi/2 in four parts -> r as a double-double, |r| <= pi/4 (+ slack);
//   2. fast path: table of sin/cos(i/128) as double-double + short Taylor polynomials in delta = |r| - i/128,
//      assembled as hi + lo with relative error <= kFastErr; accept RN(hi+lo) when the rounding cannot be affected;
//   3. otherwise (≈1e-5 of arguments) redo the evaluation entirely in double-double (error ~2^-100).
// |x| >= 2^20 (outside S2LatLng::is_valid's domain by a factor 3e5) and non-finite x use the CUDA / libm function.
#pragma once
#include "s2b_core.cuh"
#include "s2b_sincos_tables.h"
#include "s2b_sincos_tab512.h"

namespace s2b {

struct SinCos { double s, c; };

S2B_HD void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
S2B_HD void fast_two_sum(double a, double b, double& s, double& e) {  // requires |a| >= |b| or a == 0
  s = a + b;
  e = b - (s - a);
}
S2B_HD void two_prod(double a, double b, double& p, double& e) {
  p = a * b;
  e = fma(a, b, -p);
}

// double-double helpers for the slow path
struct DD { double h, l; };
S2B_HD DD dd_norm(double h, double l) { DD r; two_sum(h, l, r.h, r.l); return r; }
S2B_HD DD dd_add(DD a, DD b) {
  double s, e, t, f;
  two_sum(a.h, b.h, s, e);
  two_sum(a.l, b.l, t, f);
  e += t;
  double s2, e2;
  fast_two_sum(s, e, s2, e2);
  e2 += f;
  DD r;
  fast_two_sum(s2, e2, r.h, r.l);
  return r;
}
S2B_HD DD dd_mul(DD a, DD b) {
  double p, e;
  two_prod(a.h, b.h, p, e);
  e += a.h * b.l + a.l * b.h;
  DD r;
  fast_two_sum(p, e, r.h, r.l);
  return r;
}
S2B_HD DD dd_neg(DD a) { return DD{-a.h, -a.l}; }

// Relative error bound of the fast path's hi+lo (measured max 2^-69.3 against mpmath over 4e6 arguments incl.
// the neighbourhoods of multiples of pi/2 and of the table nodes; see tests/test_sincos.py). 2^-67 leaves ~7x margin.
static constexpr double kFastErr = 0x1p-67;

// Reduction: x = k*(pi/2) + (rh + rl), |x| < 2^20. Returns k.
S2B_HD int reduce_pio2(double x, double& rh, double& rl) {
  using namespace sc_tab;
  double kd = rint(x * kTwoOverPi);
  double t1 = x - kd * kPio2_1;  // product exact (33+20 bits), difference exact (Sterbenz)
  double w2 = kd * kPio2_2;      // exact
  double a, e1;
  two_sum(t1, -w2, a, e1);
  double wh, wl;
  two_prod(kd, kPio2_3, wh, wl);
  double b, e2;
  two_sum(a, -wh, b, e2);
  double low = ((e1 + e2) - wl) - kd * kPio2_4;
  two_sum(b, low, rh, rl);
  return (int)kd;
}

// sin and cos of a = ah + al, 0 <= a <= ~0.81, as hi+lo pairs. tab = sc_tab::kSinCosTable (shared memory on device).
template <bool kSlow>
S2B_HD void sincos_core(double ah, double al, const double (*tab)[4], double& sh, double& sl, double& ch, double& cl) {
  int i = (int)(ah * 128.0 + 0.5);
  double d0 = ah - (double)i * 0.0078125;  // exact
  double dh, dl;
  two_sum(d0, al, dh, dl);
  const double Sh = tab[i][0], Sl = tab[i][1], Ch = tab[i][2], Cl = tab[i][3];
  if (!kSlow) {
    double qh, ql;
    two_prod(dh, dh, qh, ql);
    double z = qh;
    double sp = dh * (z * (-1.0 / 6 + z * (1.0 / 120 + z * (-1.0 / 5040))));         // sin(d) - d
    double cp = (z * z) * (1.0 / 24 + z * (-1.0 / 720 + z * (1.0 / 40320)));         // cos(d) - 1 + d^2/2
    double cm1h = -0.5 * qh;                                                          // cos(d) - 1 = cm1h + cm1l
    double cm1l = (-0.5 * ql - dh * dl) + cp;
    double sdl = dl + sp;                                                             // sin(d) = dh + sdl
    double p1, e1, p2, e2, s1, t1, s2, t2;
    // sin(a) = S*cos(d) + C*sin(d)
    two_prod(Ch, dh, p1, e1);
    two_prod(Sh, cm1h, p2, e2);
    two_sum(Sh, p1, s1, t1);
    two_sum(s1, p2, s2, t2);
    double low = (((Sl * cm1h + Cl * dh) + (Sh * cm1l + Ch * sdl)) + (e1 + e2)) + ((t1 + t2) + Sl);
    fast_two_sum(s2, low, sh, sl);
    // cos(a) = C*cos(d) - S*sin(d)
    two_prod(-Sh, dh, p1, e1);
    two_prod(Ch, cm1h, p2, e2);
    two_sum(Ch, p1, s1, t1);
    two_sum(s1, p2, s2, t2);
    low = (((Cl * cm1h - Sl * dh) + (Ch * cm1l - Sh * sdl)) + (e1 + e2)) + ((t1 + t2) + Cl);
    fast_two_sum(s2, low, ch, cl);
  } else {
    using sc_tab::kInvFact;
    DD d{dh, dl};
    DD z = dd_mul(d, d);
    // sin(d)/d - 1 = z*(-1/3! + z*(1/5! + z*(-1/7! + z*(1/9! + z*(-1/11! + z/13!)))))
    DD ps{kInvFact[13][0], kInvFact[13][1]};
    ps = dd_add(dd_mul(ps, z), DD{-kInvFact[11][0], -kInvFact[11][1]});
    ps = dd_add(dd_mul(ps, z), DD{kInvFact[9][0], kInvFact[9][1]});
    ps = dd_add(dd_mul(ps, z), DD{-kInvFact[7][0], -kInvFact[7][1]});
    ps = dd_add(dd_mul(ps, z), DD{kInvFact[5][0], kInvFact[5][1]});
    ps = dd_add(dd_mul(ps, z), DD{-kInvFact[3][0], -kInvFact[3][1]});
    ps = dd_mul(ps, z);
    DD sind = dd_add(d, dd_mul(d, ps));
    // cos(d) - 1 = z*(-1/2! + z*(1/4! + z*(-1/6! + z*(1/8! + z*(-1/10! + z/12!)))))
    DD pc{kInvFact[12][0], kInvFact[12][1]};
    pc = dd_add(dd_mul(pc, z), DD{-kInvFact[10][0], -kInvFact[10][1]});
    pc = dd_add(dd_mul(pc, z), DD{kInvFact[8][0], kInvFact[8][1]});
    pc = dd_add(dd_mul(pc, z), DD{-kInvFact[6][0], -kInvFact[6][1]});
    pc = dd_add(dd_mul(pc, z), DD{kInvFact[4][0], kInvFact[4][1]});
    pc = dd_add(dd_mul(pc, z), DD{-0.5, 0.0});
    DD cm1 = dd_mul(pc, z);
    DD S{Sh, Sl}, C{Ch, Cl};
    // sin(a) = S + (S*cm1 + C*sind), cos(a) = C + (C*cm1 - S*sind)
    DD rs = dd_add(S, dd_add(dd_mul(S, cm1), dd_mul(C, sind)));
    DD rc = dd_add(C, dd_add(dd_mul(C, cm1), dd_neg(dd_mul(S, sind))));
    sh = rs.h; sl = rs.l; ch = rc.h; cl = rc.l;
  }
}

// true when RN(hi + lo + e) is the same double for every |e| <= kFastErr*|hi|.
S2B_HD bool ziv_ok(double hi, double lo) {
  double E = fabs(hi) * kFastErr;
  return hi + (lo + E) == hi + (lo - E);
}

// Correctly rounded sin(x), cos(x). n_slow (nullable) counts arguments that needed the double-double stage.
template <bool kForceSlow = false>
S2B_HD SinCos sincos_cr(double x, const double (*tab)[4], int* n_slow = nullptr) {
  double ax = fabs(x);
  SinCos out;
  if (!(ax < 1048576.0)) {  // huge, inf or NaN: outside the supported domain, defer to the platform function
    out.s = sin(x);
    out.c = cos(x);
    return out;
  }
  if (ax < 0x1p-27) { out.s = x; out.c = 1.0; return out; }  // |x^2/6| < 2^-56: RN(sin x) = x, RN(cos x) = 1
  // Always reduce: for |x| <= pi/4 the multiple is 0 and every step below is exact (rh = x, rl = 0), so the common
  // path has no data-dependent branch (lat is mostly below pi/4, lng mostly above: a branch here splits every warp).
  double rh, rl;
  int k = reduce_pio2(x, rh, rl);
  bool negr = rh < 0;
  double ah = negr ? -rh : rh, al = negr ? -rl : rl;
  double sh, sl, ch, cl;
  bool ok = false;
  if (!kForceSlow) {
    sincos_core<false>(ah, al, tab, sh, sl, ch, cl);
    ok = ziv_ok(sh, sl) && ziv_ok(ch, cl);
  }
  if (!ok) {
    sincos_core<true>(ah, al, tab, sh, sl, ch, cl);
    if (n_slow) ++*n_slow;
  }
  double sr = negr ? -sh : sh;  // sin(r), cos(r); sh = RN(sh + sl) by construction
  double cr = ch;
  switch (k & 3) {
    case 0: out.s = sr; out.c = cr; break;
    case 1: out.s = cr; out.c = -sr; break;
    case 2: out.s = -sr; out.c = -cr; break;
    default: out.s = -cr; out.c = sr; break;
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// Fast sin/cos with a PROVEN-SMALL absolute error, used by the encoder as a filter in front of sincos_cr: the cell id
// computed from these values is identical to the one computed from correctly rounded values unless the point is
// within kIjTolerance of a leaf-cell boundary (or two |components| nearly tie for the face), and exactly those points
// are re-evaluated with sincos_cr (s2b_encode.cu). So the ids are those of the correctly rounded definition, at a
// third of the FP64 work.
//   reduction: x - k*pi/2 with three Cody-Waite parts in plain double (error <= 0.5 ulp(r) + k*2^-120)
//   sin(r), cos(r): degree-13 / degree-14 minimax polynomials on |r| <= pi/4 (the classic fdlibm kernel coefficients,
//   approximation error < 2^-58), Horner in r^2
// Absolute error bound assumed by the encoder: kFastSinCosAbsErr = 4.5e-16 (about twice the measured maximum,
// tests/test_encode_cpu.py::test_fast_sincos_error_bound).
static constexpr double kFastSinCosAbsErr = 4.5e-16;

// Coefficients live in constant memory on the device so that DFMA reads them as c[bank][offset] operands; as
// immediates each would cost two IMAD.MOV per use (measured: 175 IMAD/point before, see profiles/).
#define S2B_FAST_COEFS { \
  -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06, \
  -2.50507602534068634195e-08, 1.58969099521155010221e-10, \
  4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07, \
  2.08757232129817482790e-09, -1.13596475577881948265e-11, \
  sc_tab::kTwoOverPi, -sc_tab::kPio2_1, -sc_tab::kPio2_2, -sc_tab::kPio2_3 }
// sincos_tab: 256/pi, -pi/256 in three parts, then the sin / cos series coefficients
#define S2B_TAB_COEFS { sc_tab::k256OverPi, -sc_tab::kPi256_1, -sc_tab::kPi256_2, -sc_tab::kPi256_3, \
  1.0 / 120, -1.0 / 6, -1.0 / 720, 1.0 / 24 }
#if defined(__CUDACC__)
__constant__ double c_fast_coef[16] = S2B_FAST_COEFS;
__constant__ double c_tab_coef[8] = S2B_TAB_COEFS;
#endif
static constexpr double kTabCoefHost[8] = S2B_TAB_COEFS;
#if defined(__CUDA_ARCH__)
#define S2B_TC(i) c_tab_coef[i]
#else
#define S2B_TC(i) kTabCoefHost[i]
#endif
static constexpr double kFastCoefHost[16] = S2B_FAST_COEFS;
#if defined(__CUDA_ARCH__)
#define S2B_FC(i) c_fast_coef[i]
#else
#define S2B_FC(i) kFastCoefHost[i]
#endif

// -x or x: an integer XOR on the sign bit (the FP64 negate-and-select sequence costs a DADD and two FSELs).
S2B_HD double flip_sign_if(double x, bool flip) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ (flip ? (int)0x80000000 : 0), __double2loint(x));
#else
  return flip ? -x : x;
#endif
}

// This routine is a filter, not part of the bit-exact path, so fused multiply-adds are used freely.
S2B_HD SinCos sincos_fast(double x) {  // requires |x| < 2^20
  double kd = rint(x * S2B_FC(12));
  double r = fma(kd, S2B_FC(15), fma(kd, S2B_FC(14), fma(kd, S2B_FC(13), x)));
  double z = r * r;
  double ps = S2B_FC(5);
  ps = fma(ps, z, S2B_FC(4));
  ps = fma(ps, z, S2B_FC(3));
  ps = fma(ps, z, S2B_FC(2));
  ps = fma(ps, z, S2B_FC(1));
  ps = fma(ps, z, S2B_FC(0));
  double sr = fma(r * z, ps, r);
  double pc = S2B_FC(11);
  pc = fma(pc, z, S2B_FC(10));
  pc = fma(pc, z, S2B_FC(9));
  pc = fma(pc, z, S2B_FC(8));
  pc = fma(pc, z, S2B_FC(7));
  pc = fma(pc, z, S2B_FC(6));
  double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  int q = (int)kd;
  // quadrant: swap for odd k, negate sin for k = 2,3 (mod 4), cos for k = 1,2
  double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
  SinCos out;
  out.s = flip_sign_if(s0, (q & 2) != 0);
  out.c = flip_sign_if(c0, ((q + 1) & 2) != 0);
  return out;
}

// The filter the encoder actually uses: the same contract as sincos_fast (|error| <= kFastSinCosAbsErr, requires
// |x| < 2^20) at about half the instructions. x = k * (pi/256) + d with |d| <= pi/512 (three-part Cody-Waite reduction, the
// first step exact); sin and cos of k * pi/256 come from a table of the WHOLE circle (512 x 16 bytes, shared memory on the
// device), so there is no quadrant logic at all; sin(d) and cos(d) - 1 need only d^5 and d^6 terms (remainders < 7e-20):
//   sin x = S + (S * (cos d - 1) + C * sin d),   cos x = C + (C * (cos d - 1) - S * sin d).
// Error: table rounding <= 5.6e-17, final addition <= 5.6e-17, everything else < 3e-18: <= 1.2e-16 in total (measured
// 1.1e-16 against mpmath, tests/test_encode_cpu.py::test_fast_sincos_error_bound).
S2B_HD SinCos sincos_tab(double x, const double (*tab512)[2]) {
  // constants come from the constant bank (as immediates each costs two uniform moves per use)
  const double kd = rint(x * S2B_TC(0));
  const double d = fma(kd, S2B_TC(3), fma(kd, S2B_TC(2), fma(kd, S2B_TC(1), x)));
  const int k = (int)kd & 511;
#if defined(__CUDA_ARCH__)
  const double2 sc = *(const double2*)tab512[k];
  const double S = sc.x, C = sc.y;
#else
  const double S = tab512[k][0], C = tab512[k][1];
#endif
  const double d2 = d * d;
  const double sd = fma(d * d2, fma(d2, S2B_TC(4), S2B_TC(5)), d);                // sin d
  const double cm = d2 * fma(d2, fma(d2, S2B_TC(6), S2B_TC(7)), -0.5);            // cos d - 1
  SinCos out;
  out.s = S + fma(S, cm, C * sd);
  out.c = C + fma(C, cm, -(S * sd));
  return out;
}

S2B_HD P3 latlng_to_point_fast(double lat, double lng, const double (*tab512)[2]) {
  SinCos phi = sincos_tab(lat, tab512);
  SinCos theta = sincos_tab(lng, tab512);
  return P3{theta.c * phi.c, theta.s * phi.c, phi.s};
}

// Reciprocal and reciprocal square root for the filter path: hardware seed (MUFU.RCP64H / RSQ64H, ~2^-20 relative) and
// one third-order step, relative error < 2^-57 before the final rounding. No special-case handling: callers pass normal,
// finite arguments of moderate magnitude (the major axis of a unit vector; 1 + 3|u| in [1, 4]).
S2B_HD double rcp_filter(double m) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(m));
  const double e = fma(-m, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
  return 1.0 / m;
#endif
}
S2B_HD double rsqrt_filter(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
#else
  return 1.0 / sqrt(x);
#endif
}

// point_to_face_fij for the filter path: same face selection, but u, v, s, t are computed with the approximations above
// and fused multiply-adds (one shared reciprocal, no IEEE division / square root sequences): 2^30 * s is off by at most
// ~5e-7 leaf cells more than with the IEEE operations (u: <= 2 ulp instead of 0.5; sqrt: <= 2 ulp), see the budget below.
S2B_HD int point_to_face_fij_filter(const P3& p, double* fi, double* fj) {
  const double ax = fabs(p.x), ay = fabs(p.y), az = fabs(p.z);
  const int axis = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  const double m = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
  const bool neg = m < 0;
  const double a = axis == 0 ? p.y : -p.x;
  const double b = axis == 2 ? -p.y : p.z;
  const double r = rcp_filter(m);
  const double u = (neg ? b : a) * r, v = (neg ? a : b) * r;
  // 2^30 * s = u >= 0 ? 2^29 * sqrt(w) : 2^30 - 2^29 * sqrt(w),  w = 1 + 3|u|,  sqrt(w) = w * rsqrt(w)
  const double wu = fma(3.0, fabs(u), 1.0), wv = fma(3.0, fabs(v), 1.0);
  const double su = wu * rsqrt_filter(wu), sv = wv * rsqrt_filter(wv);
  *fi = u >= 0 ? 536870912.0 * su : fma(-536870912.0, su, 1073741824.0);
  *fj = v >= 0 ? 536870912.0 * sv : fma(-536870912.0, sv, 1073741824.0);
  return axis + (neg ? 3 : 0);
}

// S2LatLng::ToPoint (s2latlng.cc:68-77): (cos(lng)*cos(lat), sin(lng)*cos(lat), sin(lat)).
S2B_HD P3 latlng_to_point(double lat, double lng, const double (*tab)[4], int* n_slow = nullptr) {
  SinCos phi = sincos_cr(lat, tab, n_slow);
  SinCos theta = sincos_cr(lng, tab, n_slow);
  return P3{theta.c * phi.c, theta.s * phi.c, phi.s};
}

// Sensitivity of the leaf cell to small errors in p (s2b_capi.h boundary_flag_out, and the filter in front of the
// correctly rounded sin/cos). With |error| <= e in each of the four sin/cos values: x,y are off by <= 2e + 1.2e-16,
// z by <= e; u,v by <= (3e + 2.4e-16)/0.577 + 1.2e-16 (the major axis is >= 1/sqrt(3)); s,t by <= 0.75 of that + 2e-16;
// i.e. in units of leaf cells (x 2^30):  e = 4.5e-16 (fast sin/cos) -> 3.1e-6 ;  e = 1.1e-16 (1 ulp) -> 1.1e-6.
// The filter path's approximate reciprocal / square root (point_to_face_fij_filter) add <= 3.3e-16 to u,v and
// <= 2.2e-16 to s,t: + 5e-7 leaf cells, 3.6e-6 in total (measured maximum on the GPU: tests/test_encode_gpu.py).
S2B_HD bool near_cell_boundary(double fi, double fj, double tol_ij) {
  // A face flip needs two |components| to (nearly) tie, i.e. |u| or |v| within ~1e-14 of 1 on the chosen face, i.e.
  // fi or fj within ~4e-6 of 0 or 2^30: already covered by the distance-to-integer test below (tolerances >= 4e-6).
  const double di = fabs(fi - rint(fi)), dj = fabs(fj - rint(fj));
  return !(di > tol_ij && dj > tol_ij);   // NaN -> true
}
static constexpr double kFilterTolIj = 1.6e-5;   // fast-path filter: 4.4x the 3.6e-6 bound
static constexpr double kFlagTolIj = 4e-6;      // user-visible flag: 1-ulp sin/cos changes, 3.5x margin

// lat/lng -> (face, fi, fj) with the filter described above. *sensitive reports whether the correctly rounded path ran.
S2B_HD int latlng_to_face_fij(double lat, double lng, const double (*tab)[4], const double (*tab512)[2], bool force_cr, P3* p, double* fi, double* fj, bool* sensitive) {
  int face = 0;
  // out-of-domain angles (|x| >= 2^20, inf, NaN) always take the correctly-rounded path, which handles them
  bool sens = !(fabs(lat) < 1048576.0 && fabs(lng) < 1048576.0) || force_cr;
  if (!sens) {
    *p = latlng_to_point_fast(lat, lng, tab512);
    face = point_to_face_fij_filter(*p, fi, fj);
    sens = near_cell_boundary(*fi, *fj, kFilterTolIj);
  }
  if (sens) {
    *p = latlng_to_point(lat, lng, tab);
    face = point_to_face_fij(*p, fi, fj);
  }
  *sensitive = sens;
  return face;
}

// S2Point -> (face, fi, fj) with the same idea applied to the IEEE division / square root sequences of
// point_to_face_fij: the approximate version is exact about the face (same comparisons) and within a few ulps of
// 2^30 (7e-7 leaf cells; measured maximum in tests/test_encode_gpu.py) in fi and fj, so unless fi or fj is within
// kXyzFilterTolIj of an integer the leaf cell — floor(fi), floor(fj) — is the one the IEEE sequence gives. Points that
// are closer (about 3e-5 of all points), non-finite input and major axes outside [2^-500, 2^500] take the IEEE path.
static constexpr double kXyzFilterTolIj = 8e-6;   // > kFlagTolIj, so boundary flags are always computed from IEEE values
S2B_HD int point_to_face_fij_checked(const P3& p, double* fi, double* fj) {
  int face = point_to_face_fij_filter(p, fi, fj);
  const double m = fmax(fmax(fabs(p.x), fabs(p.y)), fabs(p.z));
  const bool trust = m > 3.0549363634996047e-151 && m < 3.2733906078961419e+150 && !near_cell_boundary(*fi, *fj, kXyzFilterTolIj);
  if (!trust) face = point_to_face_fij(p, fi, fj);
  return face;
}

}  // namespace s2b
