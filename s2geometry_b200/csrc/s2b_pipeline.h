// s2b_pipeline.h — the chunked host<->device pipeline behind every host-pointer entry point of the C ABI: the batch is
// streamed through the device in chunks on three CUDA streams so that the H2D copy of chunk k+1, the kernel of chunk k
// and the D2H copy of chunk k-1 overlap (PCIe is full duplex). One workspace per host thread (and device): the
// multi-GPU driver (s2b_multi.cu) runs one such pipeline per GPU, each on its own worker thread.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/s2b_capi.h"

namespace s2b {

int set_error(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define S2B_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// Chunked host<->device pipeline. One workspace per (thread, device): three slots of device buffers + streams.
struct Slot {
  cudaStream_t stream = nullptr;
  void* in = nullptr;
  void* out = nullptr;
  size_t in_cap = 0, out_cap = 0;
};

struct Workspace {
  int device = -1;
  Slot slot[3];
  int ensure(int dev, size_t in_bytes, size_t out_bytes) {
    if (device != dev) {
      release();
      device = dev;
    }
    for (Slot& s : slot) {
      if (!s.stream) S2B_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      if (s.in_cap < in_bytes) {
        if (s.in) cudaFree(s.in);
        s.in = nullptr; s.in_cap = 0;
        S2B_CUDA(cudaMalloc(&s.in, in_bytes));
        s.in_cap = in_bytes;
      }
      if (s.out_cap < out_bytes) {
        if (s.out) cudaFree(s.out);
        s.out = nullptr; s.out_cap = 0;
        S2B_CUDA(cudaMalloc(&s.out, out_bytes));
        s.out_cap = out_bytes;
      }
    }
    return S2B_OK;
  }
  void release() {
    for (Slot& s : slot) {
      if (s.in) cudaFree(s.in);
      if (s.out) cudaFree(s.out);
      if (s.stream) cudaStreamDestroy(s.stream);
      s = Slot();
    }
  }
};

inline thread_local Workspace t_ws;

constexpr size_t kChunk = size_t(1) << 22;  // points per chunk (4 Mi): 64-100 MB per slot, >> launch overheads

// One output stream of a pipelined call: host base pointer, bytes per point, and (for multi-array outputs) the
// number of arrays laid out n apart on the host but chunk-size apart in the device slot.
struct OutSpec { void* host; size_t bytes_per_point; int arrays; size_t host_stride = 0; /* points between the arrays on the host; 0 = n */ };

struct NoChunkHook {
  int operator()(int, size_t, size_t, const void*, void**, cudaStream_t) const { return S2B_OK; }
};

// launch(in_dev, count, out_dev[], stream): out_dev[j] points at the slot region for OutSpec j.
// done(slot, off, count, in_dev, out_dev[], stream) runs on the host once chunk [off, off + count) has fully landed in
// the caller's buffers (its slot stream is idle and its device buffers are still intact), in chunk order.
template <class Launch, class Done = NoChunkHook>
int run_pipelined(const void* in_host, size_t in_bpp, size_t n, const OutSpec* outs, int n_outs, Launch launch, Done done = Done()) {
  int dev = 0;
  S2B_CUDA(cudaGetDevice(&dev));
  size_t chunk = n < kChunk ? n : kChunk;
  size_t out_bytes = 0;
  auto region = [&](int j) { return outs[j].host ? ((outs[j].bytes_per_point * outs[j].arrays * chunk + 255) / 256) * 256 : size_t(0); };
  for (int j = 0; j < n_outs; ++j) out_bytes += region(j);
  int rc = t_ws.ensure(dev, in_bpp * chunk, out_bytes ? out_bytes : 256);
  if (rc) return rc;
  // Once copies are in flight they reference the caller's buffers: on any failure drain the slot streams before
  // returning, so nothing touches those buffers after the call has reported an error.
  auto drain_and_return = [&](int code) {
    for (Slot& s : t_ws.slot) cudaStreamSynchronize(s.stream);
    cudaGetLastError();
    return code;
  };
#define S2B_CUDA_DRAIN(call)                                                    \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) return drain_and_return(cuda_fail(e__, #call));     \
  } while (0)
  struct Pending { bool live = false; size_t off = 0, cnt = 0; } pending[3];
  auto slot_regions = [&](Slot& s, void** out_dev) {
    char* cur = (char*)s.out;
    for (int j = 0; j < n_outs; ++j) {
      out_dev[j] = cur;
      cur += region(j);
    }
  };
  auto complete = [&](int si) -> int {   // slot stream already synchronised
    if (!pending[si].live) return S2B_OK;
    pending[si].live = false;
    void* out_dev[4] = {nullptr, nullptr, nullptr, nullptr};
    slot_regions(t_ws.slot[si], out_dev);
    return done(si, pending[si].off, pending[si].cnt, t_ws.slot[si].in, out_dev, t_ws.slot[si].stream);
  };
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, ++k) {
    const int si = (int)(k % 3);
    Slot& s = t_ws.slot[si];
    size_t cnt = n - off < chunk ? n - off : chunk;
    // the slot's previous D2H must have drained before its buffers are reused
    if (k >= 3) {
      S2B_CUDA_DRAIN(cudaStreamSynchronize(s.stream));
      int drc = complete(si);
      if (drc != S2B_OK) return drain_and_return(drc);
    }
    S2B_CUDA_DRAIN(cudaMemcpyAsync(s.in, (const char*)in_host + off * in_bpp, cnt * in_bpp, cudaMemcpyHostToDevice, s.stream));
    void* out_dev[4] = {nullptr, nullptr, nullptr, nullptr};
    slot_regions(s, out_dev);
    pending[si].live = true; pending[si].off = off; pending[si].cnt = cnt;
    int lrc = launch(si, s.in, cnt, out_dev, s.stream);
    if (lrc != S2B_OK) return drain_and_return(lrc);
    for (int j = 0; j < n_outs; ++j) {
      if (!outs[j].host) continue;
      for (int a = 0; a < outs[j].arrays; ++a) {
        // device arrays are cnt apart (the kernels index array a at base + a*cnt), host arrays n apart
        S2B_CUDA_DRAIN(cudaMemcpyAsync((char*)outs[j].host + ((size_t)a * (outs[j].host_stride ? outs[j].host_stride : n) + off) * outs[j].bytes_per_point,
                                       (char*)out_dev[j] + (size_t)a * cnt * outs[j].bytes_per_point,
                                       cnt * outs[j].bytes_per_point, cudaMemcpyDeviceToHost, s.stream));
      }
    }
  }
  // drain in chunk order: the oldest pending chunk sits in slot k % 3
  for (size_t d = 0; d < 3; ++d) {
    const int si = (int)((k + d) % 3);
    S2B_CUDA_DRAIN(cudaStreamSynchronize(t_ws.slot[si].stream));
    int drc = complete(si);
    if (drc != S2B_OK) return drain_and_return(drc);
  }
#undef S2B_CUDA_DRAIN
  return S2B_OK;
}

}  // namespace s2b
