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
    int face = latlng_to_face_fij(ll[2 * i], ll[2 * i + 1], sc_tab::kSinCosTable, sc_tab::kSinCos512, force_cr != 0, &p, &fi, &fj, &sens);
    ns += sens;
    uint64_t id = from_face_ij(face, fij_to_ij(fi), fij_to_ij(fj), kHilbert.pos);
    out[i] = level < 30 ? cellid_parent(id, level) : id;
  }
  if (n_sensitive) *n_sensitive = ns;
}

void hs_sincos_fast(const double* x, size_t n, double* s, double* c) {
  for (size_t i = 0; i < n; ++i) { SinCos r = sincos_fast(x[i]); s[i] = r.s; c[i] = r.c; }
}
void hs_sincos_tab(const double* x, size_t n, double* s, double* c) {
  for (size_t i = 0; i < n; ++i) { SinCos r = sincos_tab(x[i], sc_tab::kSinCos512); s[i] = r.s; c[i] = r.c; }
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
  static uint16_t chain[16384];   // the encode kernel's repacked copy (from_face_ij_chain)
  if (!init) { MakeWideHilbert(wide); for (int k = 0; k < 16384; ++k) chain[k] = hilbert_chain_repack(wide[k]); init = true; }
  uint64_t bad = 0;
  for (size_t k = 0; k < n; ++k)
    for (int face = 0; face < 6; ++face) {
      bad += from_face_ij(face, ij[2 * k], ij[2 * k + 1], kHilbert.pos) != from_face_ij_wide(face, ij[2 * k], ij[2 * k + 1], wide);
      bad += from_face_ij(face, ij[2 * k], ij[2 * k + 1], kHilbert.pos) != from_face_ij_chain(face, ij[2 * k], ij[2 * k + 1], (const unsigned char*)chain);
    }
  return bad;
}

extern "C" void hs_cellid_from_token(const char* tok16, const uint8_t* len, size_t n, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = cellid_from_token((const unsigned char*)tok16 + 16 * i, len[i]);
}

extern "C" void hs_edge_neighbors(const uint64_t* ids, size_t n, uint64_t* out4) {
  for (size_t i = 0; i < n; ++i) cellid_edge_neighbors(ids[i], kHilbert.pos, kHilbert.ij, out4 + 4 * i);
}

extern "C" void hs_cellid_traverse(const uint64_t* ids, const uint64_t* others, const int64_t* steps, size_t n, uint64_t* children4, uint64_t* adv,
                                   uint64_t* adv_wrap, int8_t* ancestor) {
  for (size_t i = 0; i < n; ++i) {
    for (int k = 0; k < 4; ++k) children4[4 * i + k] = (ids[i] & 1) ? 0 : cellid_child(ids[i], k);
    adv[i] = cellid_advance(ids[i], steps[i], false);
    adv_wrap[i] = cellid_advance(ids[i], steps[i], true);
    ancestor[i] = (int8_t)cellid_common_ancestor_level(ids[i], others[i]);
  }
}

// ---- counter-based synthetic streams (s2b_synth.cuh) and the strict host-libm lat/lng mode, host compilation ----
#include <math.h>
#include "s2b_synth.cuh"

extern "C" {

void hs_synth_sphere_points(uint64_t seed, uint64_t first, size_t n, double* out) {
  for (size_t i = 0; i < n; ++i) { P3 p = synth_sphere_point(seed, first + i); out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z; }
}
void hs_synth_sphere_latlng(uint64_t seed, uint64_t first, size_t n, double* out) {
  for (size_t i = 0; i < n; ++i) synth_sphere_latlng(seed, first + i, &out[2 * i], &out[2 * i + 1]);
}
void hs_synth_cap_points(uint64_t seed, uint64_t first, size_t n, const double* frame9, double height, double* out) {
  for (size_t i = 0; i < n; ++i) { P3 p = synth_cap_point(seed, first + i, frame9, height); out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z; }
}

// What the encode kernel + the strict completion do per point (s2b_encode.cu / s2b_capi.cu): the device id, and for the
// points the kernel would hand back (sensitive && near a boundary at kFlagTolIj, or out-of-domain angles) the id from this
// host's sin/cos. *n_reeval counts the latter.
void hs_encode_latlng_strict(const double* ll, size_t n, uint64_t* out, uint64_t* n_reeval) {
  uint64_t nr = 0;
  for (size_t i = 0; i < n; ++i) {
    const double lat = ll[2 * i], lng = ll[2 * i + 1];
    P3 p; double fi, fj; bool sens;
    int face = latlng_to_face_fij(lat, lng, sc_tab::kSinCosTable, sc_tab::kSinCos512, false, &p, &fi, &fj, &sens);
    uint64_t id = from_face_ij(face, fij_to_ij(fi), fij_to_ij(fj), kHilbert.pos);
    if (sens && (near_cell_boundary(fi, fj, kFlagTolIj) || !(fabs(lat) < 1048576.0 && fabs(lng) < 1048576.0))) {
      ++nr;
      const double sp = sin(lat), cp = cos(lat), st = sin(lng), ct = cos(lng);
      id = cellid_from_point(P3{ct * cp, st * cp, sp}, kHilbert.pos);
    }
    out[i] = id;
  }
  if (n_reeval) *n_reeval = nr;
}

}  // extern "C"
