// s2b_diag.cu — diagnostics: a DFMA-chain microbenchmark that measures this GPU's FP64 (CUDA-core) peak, the second
// roofline denominator of the encode kernels (MEASURED_PEAKS.json records only HBM and bf16 tensor peaks).
#include "../../include/s2b_capi.h"
#include "s2b_internal.h"

namespace s2b {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chain alive
}

}  // namespace s2b

extern "C" int s2b_diag_fp64_peak(double* dfma_per_second_out) {
  using namespace s2b;
  if (!dfma_per_second_out) return S2B_EINVAL;
  double* buf = nullptr;
  int blocks = sm_count() * 8, iters = 1 << 14;
  if (cudaMalloc(&buf, sizeof(double) * blocks * 256) != cudaSuccess) return S2B_ENOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, 0);
    dfma_kernel<<<blocks, 256>>>(buf, iters, 0.9999999, 1e-9);
    count_launch();
    cudaEventRecord(e1, 0);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(buf); return S2B_ECUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  *dfma_per_second_out = (double)blocks * 256 * 8.0 * iters / (best * 1e-3);
  return S2B_OK;
}
