// tests/hostsim/hostsim_pred.cc — TEST-ONLY host compilation of s2b_predicates.cuh.
#include <cstddef>
#include <cstdint>
#include "s2b_predicates.cuh"

using namespace s2b;

static inline P3 P(const double* p) { return P3{p[0], p[1], p[2]}; }

extern "C" {

// kind: 0 Sign, 1 TriageSign, 2 StableSign, 3 ExactSign(perturb), 4 ExpensiveSign, 5 ExactSign(no perturb)
void hs_sign(const double* abc, size_t n, int kind, int8_t* out, uint64_t* n_exact_out) {
  unsigned long long before = host_exact_count();
  for (size_t i = 0; i < n; ++i) {
    P3 a = P(abc + 9 * i), b = P(abc + 9 * i + 3), c = P(abc + 9 * i + 6);
    int s;
    switch (kind) {
      case 1: s = triage_sign(cross(a, b), c); break;
      case 2: s = stable_sign(a, b, c); break;
      case 3: s = exact_sign(a, b, c, true); break;
      case 4: s = expensive_sign(a, b, c); break;
      case 5: s = exact_sign(a, b, c, false); break;
      default: s = sign3(a, b, c);
    }
    out[i] = (int8_t)s;
  }
  if (n_exact_out) *n_exact_out = (uint64_t)(host_exact_count() - before);
}

// kind: 0 CrossingSign, 2 VertexCrossing, 3 EdgeOrVertexCrossing
void hs_crossing(const double* abcd, size_t n, int kind, int8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    P3 a = P(abcd + 12 * i), b = P(abcd + 12 * i + 3), c = P(abcd + 12 * i + 6), d = P(abcd + 12 * i + 9);
    int s;
    if (kind == 2) s = vertex_crossing(a, b, c, d);
    else if (kind == 3) { int cs = crossing_sign(a, b, c, d); s = cs < 0 ? 0 : (cs > 0 ? 1 : (int)vertex_crossing(a, b, c, d)); }
    else s = crossing_sign(a, b, c, d);
    out[i] = (int8_t)s;
  }
}

void hs_ortho(const double* a, size_t n, double* out) {
  for (size_t i = 0; i < n; ++i) { P3 o = ortho(P(a + 3 * i)); out[3 * i] = o.x; out[3 * i + 1] = o.y; out[3 * i + 2] = o.z; }
}

}  // extern "C"
