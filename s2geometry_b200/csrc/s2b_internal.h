// s2b_internal.h — declarations shared by the .cu translation units of libs2b.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace s2b {

struct LevelSpec { int n; int level[4]; };

// One boundary-sensitive lat/lng point: its index in the caller's batch, the angles in radians as the device evaluated
// them (for E7 input: (M_PI / 180) * (1e-7 * e7), s1angle.h:363-377) and the leaf id the device computed.
struct FixRecord { uint64_t index; double lat, lng; uint64_t id; };
struct FixPatch { uint64_t index, id; };

// Everything the encode kernel can emit for a point. ids: `levels.n` arrays of n ids (array k at ids + k*n).
// tokens (optional): ToToken of the id at levels.level[token_level_index], 16-byte slots + strlen bytes.
struct EncodeOut {
  LevelSpec levels;
  uint64_t* ids;
  uint8_t* flags;       // nullable: boundary-sensitivity flags (lat/lng input)
  int token_level_index;  // -1: no tokens
  char* tokens16;       // nullable unless token_level_index >= 0
  uint8_t* token_len;   // nullable
  int force_cr = 0;     // diagnostics: 1 = always use the correctly rounded sin/cos (skip the fast filter)
  // Strict host-libm mode (lat/lng input): points whose leaf cell could change under a <= 1 ulp change of a sin/cos value
  // (and angles outside the device routines' domain) are appended to fix_list for re-evaluation with the host's libm.
  FixRecord* fix_list = nullptr;   // nullable
  uint32_t* fix_count = nullptr;          // device counter, incremented past fix_cap on overflow
  uint32_t fix_cap = 0;
  uint32_t fix_base = 0;                  // value of *fix_count when the launch starts (the counter is never reset for the live list)
  uint32_t* fix_overflow = nullptr;       // nullable: set to 1 (mapped host memory) by an append that found the list full
  uint64_t index_base = 0;                // index of point 0 of this launch in the caller's batch
  size_t ids_stride = 0;                  // distance between the per-level id arrays; 0 = n
};

enum InputKind { kInXYZ = 0, kInLatLngRad = 1, kInLatLngE7 = 2 };

// Kernel launchers (device pointers, enqueue only). Return cudaGetLastError() of the launch.
cudaError_t launch_encode(InputKind kind, const void* in_dev, size_t n, const EncodeOut& out, cudaStream_t stream);
// Overwrites the outputs of the points in `patches` (leaf ids under the host libm) at every level + token of `out`.
cudaError_t launch_encode_patch(const FixPatch* patches_dev, uint32_t m, const EncodeOut& out, size_t n, cudaStream_t stream);
cudaError_t launch_parent(const uint64_t* ids_dev, size_t n, int level, uint64_t* out_dev, cudaStream_t stream);
cudaError_t launch_filter_deviation(InputKind kind, const double* in_dev, size_t n, unsigned long long* out2_dev, cudaStream_t stream);
cudaError_t launch_edge_neighbors(const uint64_t* ids_dev, size_t n, uint64_t* out4_dev, cudaStream_t stream);
cudaError_t launch_to_point(const uint64_t* ids_dev, size_t n, double* xyz_dev, cudaStream_t stream);
cudaError_t launch_to_latlng(const uint64_t* ids_dev, size_t n, double* latlng_dev, cudaStream_t stream);
cudaError_t launch_from_token(const char* tokens16_dev, const uint8_t* len_dev, size_t n, uint64_t* ids_dev, cudaStream_t stream);
cudaError_t launch_to_token(const uint64_t* ids_dev, size_t n, char* tokens16_dev, uint8_t* len_dev, cudaStream_t stream);

// The host-pointer encode pipeline (s2b_capi.cu), shared with the multi-GPU driver. tok_level_index < 0: no tokens.
int encode_host(InputKind kind, const void* in, size_t in_bpp, size_t n, const int* levels, int num_levels, uint64_t* ids_out,
                uint8_t* flags_out, int tok_level_index = -1, char* tok_out = nullptr, uint8_t* tok_len_out = nullptr, size_t ids_stride = 0);

void count_launch();
// Host copy of the 16384-entry 6-bit Hilbert table (built once); each .cu uploads it to its own __device__ array.
const uint16_t* wide_hilbert_host();
int sm_count();  // SMs of the current device (cached)

}  // namespace s2b
