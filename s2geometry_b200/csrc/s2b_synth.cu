// s2b_synth.cu — device generators of the counter-based synthetic inputs (s2b_synth.cuh): fill HBM with the same
// points any host (or any other rank) computes from (seed, global index). Benchmark / test support, not on the query path.
#include "../../include/s2b_capi.h"
#include "s2b_internal.h"
#include "s2b_synth.cuh"

namespace s2b {

struct CapFrame { double f[9]; double height; };

template <int kKind>
__global__ void __launch_bounds__(256) synth_kernel(uint64_t seed, uint64_t first, size_t n, double* __restrict__ out, const CapFrame cap) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    if (kKind == 1) {
      double lat, lng;
      synth_sphere_latlng(seed, first + i, &lat, &lng);
      __stcs((double2*)out + i, make_double2(lat, lng));
    } else {
      const P3 p = kKind == 0 ? synth_sphere_point(seed, first + i) : synth_cap_point(seed, first + i, cap.f, cap.height);
      double* q = out + 3 * i;
      __stcs(q, p.x); __stcs(q + 1, p.y); __stcs(q + 2, p.z);
    }
  }
}

int set_error(int code, const char* fmt, ...);

template <int kKind>
static int launch_synth(uint64_t seed, uint64_t first, size_t n, double* out, const CapFrame& cap, void* stream) {
  if (n == 0) return S2B_OK;
  if (!out) return set_error(S2B_EINVAL, "null output pointer");
  if (kKind == 1 && ((uintptr_t)out % 16)) return set_error(S2B_EINVAL, "latlng_out_dev must be 16-byte aligned");
  size_t want = (n + 255) / 256, cap_blocks = (size_t)sm_count() * 8;
  synth_kernel<kKind><<<(int)(want < cap_blocks ? want : cap_blocks), 256, 0, (cudaStream_t)stream>>>(seed, first, n, out, cap);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? S2B_OK : set_error(S2B_ECUDA, "synth launch: %s", cudaGetErrorString(e));
}

}  // namespace s2b

using namespace s2b;

extern "C" {
int s2b_synth_sphere_points_dev(uint64_t seed, uint64_t first_index, size_t n, double* xyz_out_dev, void* stream) {
  return launch_synth<0>(seed, first_index, n, xyz_out_dev, CapFrame{}, stream);
}
int s2b_synth_sphere_latlng_dev(uint64_t seed, uint64_t first_index, size_t n, double* latlng_out_dev, void* stream) {
  return launch_synth<1>(seed, first_index, n, latlng_out_dev, CapFrame{}, stream);
}
int s2b_synth_cap_points_dev(uint64_t seed, uint64_t first_index, size_t n, const double* frame9, double height, double* xyz_out_dev, void* stream) {
  if (!frame9) return set_error(S2B_EINVAL, "null frame");
  CapFrame c;
  for (int k = 0; k < 9; ++k) c.f[k] = frame9[k];
  c.height = height;
  return launch_synth<2>(seed, first_index, n, xyz_out_dev, c, stream);
}
}
