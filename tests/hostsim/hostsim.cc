// tests/hostsim/hostsim.cc — TEST-ONLY host compilation of the device headers (s2b_core.cuh etc.).
// There is no GPU in the development container, so the scalar logic that the CUDA kernels inline is compiled
// here with g++ -ffp-contract=off and checked against the oracle on the CPU. Nothing in the product loads this.
#include <cstddef>
#include <cstdint>
#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_sincos.cuh"

using namespace s2b;

extern "C" {

void hs_encode_points(const double* xyz, size_t n, int level, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) {
    uint64_t id = cellid_from_point(P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, kHilbert.pos);
    out[i] = level < 30 ? cellid_parent(id, level) : id;
  }
}

// The kernel's lat/lng path: fast bounded-error sin/cos as a filter, correctly rounded sin/cos where it matters.
// force_cr = 1: every point takes the correctly rounded path. n_sensitive counts points that took it.
void hs_encode_latlng_filtered(const double* ll, size_t n, int level, int force_cr, uint64_t* out, uint64_t* n_sensitive) {
  uint64_t ns = 0;
  for (size_t i = 0; i < n; ++i) {
    P3 p; double fi, fj; bool sens;
    int face = latlng_to_face_fij(ll[2 * i], ll[2 * i + 1], sc_tab::kSinCosTable, force_cr != 0, &p, &fi, &fj, &sens);
    ns += sens;
    uint64_t id = from_face_ij(face, fij_to_ij(fi), fij_to_ij(fj), kHilbert.pos);
    out[i] = level < 30 ? cellid_parent(id, level) : id;
  }
  if (n_sensitive) *n_sensitive = ns;
}

void hs_sincos_fast(const double* x, size_t n, double* s, double* c) {
  for (size_t i = 0; i < n; ++i) { SinCos r = sincos_fast(x[i]); s[i] = r.s; c[i] = r.c; }
}

void hs_encode_latlng(const double* ll, size_t n, int level, uint64_t* out, uint64_t* n_slow) {
  int slow = 0;
  for (size_t i = 0; i < n; ++i) {
    P3 p = latlng_to_point(ll[2 * i], ll[2 * i + 1], sc_tab::kSinCosTable, &slow);
    uint64_t id = cellid_from_point(p, kHilbert.pos);
    out[i] = level < 30 ? cellid_parent(id, level) : id;
  }
  if (n_slow) *n_slow = (uint64_t)slow;
}

// mode 0: production (Ziv), 1: forced slow path. slow_out[i] = 1 if the slow path ran.
void hs_sincos(const double* x, size_t n, int mode, double* s, double* c, uint8_t* slow_out) {
  for (size_t i = 0; i < n; ++i) {
    int ns = 0;
    SinCos r = mode ? sincos_cr<true>(x[i], sc_tab::kSinCosTable, &ns) : sincos_cr<false>(x[i], sc_tab::kSinCosTable, &ns);
    s[i] = r.s; c[i] = r.c;
    if (slow_out) slow_out[i] = (uint8_t)ns;
  }
}

// Raw hi/lo pairs of the reduced-argument kernel, to measure the fast path's error. out = n x 4 (sh, sl, ch, cl);
// red = n x 3 (k, rh, rl).
void hs_sincos_hilo(const double* x, size_t n, int slow, double* out, double* red) {
  for (size_t i = 0; i < n; ++i) {
    double rh = x[i], rl = 0;
    int k = 0;
    if (fabs(x[i]) > sc_tab::kPio4Hi) k = reduce_pio2(x[i], rh, rl);
    bool neg = rh < 0;
    double ah = neg ? -rh : rh, al = neg ? -rl : rl;
    double* o = out + 4 * i;
    if (slow) sincos_core<true>(ah, al, sc_tab::kSinCosTable, o[0], o[1], o[2], o[3]);
    else sincos_core<false>(ah, al, sc_tab::kSinCosTable, o[0], o[1], o[2], o[3]);
    red[3 * i] = k; red[3 * i + 1] = rh; red[3 * i + 2] = rl;
  }
}

void hs_cellid_to_token(const uint64_t* ids, size_t n, char* tok16, uint8_t* len) {
  for (size_t i = 0; i < n; ++i) { uint32_t w[4]; len[i] = (uint8_t)cellid_to_token(ids[i], w); __builtin_memcpy(tok16 + 16 * i, w, 16); }
}

}  // extern "C"

// from_face_ij (4-bit table, the reference's layout) vs from_face_ij_wide (6-bit table): returns the number of mismatches.
extern "C" uint64_t hs_check_wide_hilbert(const uint32_t* ij, size_t n) {
  static uint16_t wide[16384];
  static bool init = false;
  if (!init) { MakeWideHilbert(wide); init = true; }
  uint64_t bad = 0;
  for (size_t k = 0; k < n; ++k)
    for (int face = 0; face < 6; ++face)
      bad += from_face_ij(face, ij[2 * k], ij[2 * k + 1], kHilbert.pos) != from_face_ij_wide(face, ij[2 * k], ij[2 * k + 1], wide);
  return bad;
}

extern "C" void hs_cellid_from_token(const char* tok16, const uint8_t* len, size_t n, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = cellid_from_token((const unsigned char*)tok16 + 16 * i, len[i]);
}

extern "C" void hs_edge_neighbors(const uint64_t* ids, size_t n, uint64_t* out4) {
  for (size_t i = 0; i < n; ++i) cellid_edge_neighbors(ids[i], kHilbert.pos, kHilbert.ij, out4 + 4 * i);
}
