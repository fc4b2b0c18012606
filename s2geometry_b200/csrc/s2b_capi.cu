// s2b_capi.cu — the C ABI of libs2b.so (include/s2b_capi.h): argument checking, error reporting, and the
// host-pointer entry points, which stream the batch through the device in chunks on three CUDA streams so that
// the H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap (PCIe is full duplex).
// There is no CPU implementation behind these calls: without a usable CUDA device they return S2B_ECUDA.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_core.cuh"
#include "s2b_hilbert.h"
#include "s2b_internal.h"
#include "s2b_pipeline.h"

namespace s2b {

static thread_local char t_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_force_cr{0};
static std::atomic<int> g_latlng_mode{S2B_LATLNG_HOST_LIBM};
static std::atomic<uint64_t> g_strict_reevaluated{0}, g_strict_changed{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

const uint16_t* wide_hilbert_host() {
  static const std::vector<uint16_t> table = [] { std::vector<uint16_t> t(16384); MakeWideHilbert(t.data()); return t; }();
  return table.data();
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return set_error(e == cudaErrorMemoryAllocation ? S2B_ENOMEM : S2B_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

static int check_levels(const int* levels, int num_levels, LevelSpec* out) {
  if (!levels || num_levels < 1 || num_levels > 4) return set_error(S2B_EINVAL, "num_levels must be in [1,4]");
  out->n = num_levels;
  for (int k = 0; k < 4; ++k) out->level[k] = 30;
  for (int k = 0; k < num_levels; ++k) {
    if (levels[k] < 0 || levels[k] > 30) return set_error(S2B_EINVAL, "level %d out of range [0,30]", levels[k]);
    out->level[k] = levels[k];
  }
  return S2B_OK;
}

// ---------------------------------------------------------------------------------------------------
// Strict host-libm mode for lat/lng input (S2B_LATLNG_HOST_LIBM, the default).
//
// S2CellId(S2LatLng) is defined through the platform libm: S2LatLng::ToPoint calls sin and cos (s2latlng.cc:68-77,
// s1angle.h:343-357). glibc's are accurate to < 1 ulp but not correctly rounded, so for a few points in 10^5 the leaf
// cell depends on WHICH libm ran. The device cannot know that; it evaluates the correctly rounded values and reports every
// point whose leaf cell could move under a <= 1 ulp change of one of its four sin/cos values (kFlagTolIj: 3.5x margin
// over the propagated bound, s2b_sincos.cuh) plus every angle outside the device routines' domain. Those points — and
// only those — are re-evaluated here with the calling process's sin/cos and the same IEEE projection, which by
// construction is the value the reference computes in this process. Everything else is provably libm-independent.
// One list of boundary-sensitive points: device buffer = 32-byte header (the counter) + cap records, and a pinned host
// mirror, so the counter and the first records come back in ONE copy.
constexpr size_t kFixHeader = 32;
struct FixBuf {
  char* d = nullptr;
  char* h = nullptr;             // pinned
  FixPatch* d_patch = nullptr;
  cudaEvent_t ev = nullptr;
  uint32_t cap = 0, patch_cap = 0;
  int device = -1;
  uint32_t* d_count() const { return (uint32_t*)d; }
  FixRecord* d_list() const { return (FixRecord*)(d + kFixHeader); }
  uint32_t h_count() const { return *(const volatile uint32_t*)h; }
  FixRecord* h_list() const { return (FixRecord*)(h + kFixHeader); }
  void release() {
    if (d) cudaFree(d);
    if (d_patch) cudaFree(d_patch);
    if (h) cudaFreeHost(h);
    if (ev) cudaEventDestroy(ev);
    *this = FixBuf();
  }
  int ensure(int dev, uint32_t want) {
    if (device == dev && cap >= want) return S2B_OK;
    release();
    device = dev;
    const size_t bytes = kFixHeader + sizeof(FixRecord) * (size_t)want;
    S2B_CUDA(cudaMalloc((void**)&d, bytes));
    S2B_CUDA(cudaMallocHost((void**)&h, bytes));
    S2B_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cap = want;
    return S2B_OK;
  }
  int ensure_patches(uint32_t want) {
    if (patch_cap >= want) return S2B_OK;
    if (d_patch) cudaFree(d_patch);
    d_patch = nullptr; patch_cap = 0;
    S2B_CUDA(cudaMalloc(&d_patch, sizeof(FixPatch) * (size_t)want));
    patch_cap = want;
    return S2B_OK;
  }
  void attach(EncodeOut* o) const { o->fix_list = d_list(); o->fix_count = d_count(); o->fix_cap = cap; }
  // after the kernel: counter + the first `eager` records to the host, in one copy
  cudaError_t fetch(uint32_t eager, cudaStream_t st) const {
    return cudaMemcpyAsync(h, d, kFixHeader + sizeof(FixRecord) * (size_t)(eager < cap ? eager : cap), cudaMemcpyDeviceToHost, st);
  }
};
static thread_local FixBuf t_fix_slot[3];   // host-pointer pipeline: one per slot
static thread_local FixBuf t_fix_dev;       // s2b_encode_latlng_strict_dev: the piecewise fallback

// The list of s2b_encode_latlng_strict_dev: records go straight into MAPPED pinned host memory and the host re-evaluates
// them while the kernel is still running (the marker of a record is its id field, written last and never 0), so the
// host's libm work overlaps the launch and only the last few records are left when it ends.
struct LiveList {
  FixRecord* h = nullptr;        // cudaHostAlloc(mapped); device alias below
  FixRecord* d_alias = nullptr;
  uint32_t* d_count = nullptr;   // device counter
  uint32_t* h_count = nullptr;   // mapped host word: [0] = set to 1 by the kernel when the list overflowed
  uint32_t* d_overflow = nullptr; // its device alias
  cudaEvent_t ev = nullptr;
  uint32_t cap = 0;
  uint32_t base = 0;             // the device counter is never reset: records of a launch start at slot (count - base)
  int device = -1;
  bool dirty = false;
  std::vector<FixRecord> seen;
  std::vector<uint64_t> host_id;
  void release() {
    if (h) cudaFreeHost(h);
    if (d_count) cudaFree(d_count);
    if (h_count) cudaFreeHost(h_count);
    if (ev) cudaEventDestroy(ev);
    h = nullptr; d_alias = nullptr; d_count = nullptr; h_count = nullptr; d_overflow = nullptr; ev = nullptr; cap = 0; device = -1; dirty = false;
  }
  int ensure(int dev, uint32_t want) {
    if (device == dev && cap >= want) return S2B_OK;
    release();
    device = dev;
    S2B_CUDA(cudaHostAlloc((void**)&h, sizeof(FixRecord) * (size_t)want, cudaHostAllocMapped));
    memset(h, 0, sizeof(FixRecord) * (size_t)want);
    S2B_CUDA(cudaHostGetDevicePointer((void**)&d_alias, h, 0));
    S2B_CUDA(cudaMalloc((void**)&d_count, sizeof(uint32_t)));
    S2B_CUDA(cudaMemset(d_count, 0, sizeof(uint32_t)));
    S2B_CUDA(cudaHostAlloc((void**)&h_count, sizeof(uint32_t), cudaHostAllocMapped));
    *h_count = 0;
    S2B_CUDA(cudaHostGetDevicePointer((void**)&d_overflow, h_count, 0));
    S2B_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cap = want;
    base = 0;
    return S2B_OK;
  }
};
static thread_local LiveList t_live;
constexpr uint32_t kLiveCap = 1u << 16;   // 2 MB of pinned memory; 40x the uniform rate at 100M points

// S2CellId(S2LatLng::FromRadians(lat, lng)) with THIS process's libm: S2LatLng::ToPoint (s2latlng.cc:68-77; SinCos() =
// {sin(x), cos(x)}, s1angle.h:343-357) then S2CellId(S2Point) (s2cell_id.cc:309-315). Only ever applied to the
// boundary-sensitive points the device hands back.
static uint64_t leaf_id_host_libm(double lat, double lng) {
  const double sin_phi = sin(lat), cos_phi = cos(lat);
  const double sin_theta = sin(lng), cos_theta = cos(lng);
  const P3 p{cos_theta * cos_phi, sin_theta * cos_phi, sin_phi};
  return cellid_from_point(p, kHilbert.pos);
}

constexpr uint32_t kFixEagerPiece = 1u << 10;

// Re-evaluates `c` records and appends the ones whose leaf id changes.
static void reevaluate(const FixRecord* recs, size_t c, std::vector<FixPatch>* changed) {
  g_strict_reevaluated.fetch_add(c, std::memory_order_relaxed);
  const size_t before = changed->size();
  unsigned hw = std::thread::hardware_concurrency();
  const size_t threads = c < (1u << 15) ? 1 : (hw ? (hw > 32 ? 32 : hw) : 1);
  if (threads <= 1) {
    for (size_t i = 0; i < c; ++i) {
      const uint64_t id = leaf_id_host_libm(recs[i].lat, recs[i].lng);
      if (id != recs[i].id) changed->push_back(FixPatch{recs[i].index, id});
    }
  } else {   // degenerate input (e.g. every point on a cell boundary): spread the libm calls over the host cores
    std::vector<std::vector<FixPatch>> part(threads);
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; ++t)
      pool.emplace_back([&, t] {
        for (size_t i = c * t / threads; i < c * (t + 1) / threads; ++i) {
          const uint64_t id = leaf_id_host_libm(recs[i].lat, recs[i].lng);
          if (id != recs[i].id) part[t].push_back(FixPatch{recs[i].index, id});
        }
      });
    for (auto& th : pool) th.join();
    for (auto& v : part) changed->insert(changed->end(), v.begin(), v.end());
  }
  g_strict_changed.fetch_add(changed->size() - before, std::memory_order_relaxed);
}

// Turns the list of one launch over points [0, n) of `in_dev` into patches. On entry the launch has completed and
// fb.h holds its counter and the first `eager` records (FixBuf::fetch). `o` describes where that launch wrote (device
// pointers). May enqueue further copies / launches on `st` and synchronise it.
static int collect_patches(InputKind kind, const void* in_dev, size_t in_bpp, size_t n, EncodeOut o, cudaStream_t st, FixBuf& fb,
                           uint32_t eager, std::vector<FixPatch>* changed) {
  const uint32_t c = fb.h_count();
  if (c == 0) return S2B_OK;
  if (c <= fb.cap) {
    if (c > eager) {
      S2B_CUDA(cudaMemcpyAsync(fb.h_list() + eager, fb.d_list() + eager, sizeof(FixRecord) * (size_t)(c - eager), cudaMemcpyDeviceToHost, st));
      S2B_CUDA(cudaStreamSynchronize(st));
    }
    reevaluate(fb.h_list(), c, changed);
    return S2B_OK;
  }
  // More sensitive points than the list holds (input concentrated on cell boundaries): redo the range in pieces of
  // fb.cap points, each of which fits by construction. The kernel rewrites identical outputs.
  const size_t stride = o.ids_stride ? o.ids_stride : n;
  const uint64_t base = o.index_base;
  o.ids_stride = stride;
  fb.attach(&o);
  for (size_t off = 0; off < n; off += fb.cap) {
    const size_t cnt = n - off < fb.cap ? n - off : fb.cap;
    EncodeOut q = o;
    q.ids = o.ids + off;
    q.flags = o.flags ? o.flags + off : nullptr;
    q.tokens16 = o.tokens16 ? o.tokens16 + 16 * off : nullptr;
    q.token_len = o.token_len ? o.token_len + off : nullptr;
    q.index_base = base + off;
    S2B_CUDA(cudaMemsetAsync(fb.d_count(), 0, sizeof(uint32_t), st));
    cudaError_t e = launch_encode(kind, (const char*)in_dev + off * in_bpp, cnt, q, st);
    if (e != cudaSuccess) return cuda_fail(e, "encode launch");
    S2B_CUDA(fb.fetch(kFixEagerPiece, st));
    S2B_CUDA(cudaStreamSynchronize(st));
    const uint32_t cc = fb.h_count();
    if (cc > kFixEagerPiece) {
      S2B_CUDA(cudaMemcpyAsync(fb.h_list() + kFixEagerPiece, fb.d_list() + kFixEagerPiece, sizeof(FixRecord) * (size_t)(cc - kFixEagerPiece), cudaMemcpyDeviceToHost, st));
      S2B_CUDA(cudaStreamSynchronize(st));
    }
    reevaluate(fb.h_list(), cc, changed);
  }
  return S2B_OK;
}

static bool strict_latlng(InputKind kind) { return kind != kInXYZ && g_latlng_mode.load() == S2B_LATLNG_HOST_LIBM; }

constexpr uint32_t kFixCapChunk = 1u << 16;   // list capacity per 4Mi-point chunk of the host pipeline (1000x the uniform rate)
constexpr uint32_t kFixEager = 1u << 12;      // records copied back with the chunk without waiting for the count

// tok_level_index < 0: no tokens. ids_stride: points between the per-level id arrays in ids_out (0 = n; the multi-GPU
// driver passes the whole batch's size while each GPU fills its own range).
int encode_host(InputKind kind, const void* in, size_t in_bpp, size_t n, const int* levels, int num_levels,
                uint64_t* ids_out, uint8_t* flags_out, int tok_level_index, char* tok_out, uint8_t* tok_len_out, size_t ids_stride) {
  if (ids_stride == 0) ids_stride = n;
  LevelSpec lv;
  int rc = check_levels(levels, num_levels, &lv);
  if (rc) return rc;
  if (tok_level_index >= num_levels) return set_error(S2B_EINVAL, "token_level_index out of range");
  if (n == 0) return S2B_OK;
  if (!in || !ids_out || (tok_level_index >= 0 && !tok_out)) return set_error(S2B_EINVAL, "null input or output pointer");
  const bool strict = strict_latlng(kind);
  if (strict) {
    int dev = 0;
    S2B_CUDA(cudaGetDevice(&dev));
    for (FixBuf& fb : t_fix_slot) {
      rc = fb.ensure(dev, kFixCapChunk);
      if (rc) return rc;
    }
  }
  auto make_out = [&](void** out_dev) {
    EncodeOut o;
    o.levels = lv;
    o.ids = (uint64_t*)out_dev[0];
    o.flags = flags_out ? (uint8_t*)out_dev[1] : nullptr;
    o.token_level_index = tok_level_index;
    o.tokens16 = tok_level_index >= 0 ? (char*)out_dev[2] : nullptr;
    o.token_len = tok_level_index >= 0 && tok_len_out ? (uint8_t*)out_dev[3] : nullptr;
    o.force_cr = g_force_cr.load();
    return o;
  };
  std::vector<FixPatch> changed;
  OutSpec outs[4] = {{ids_out, 8, num_levels, ids_stride}, {flags_out, 1, 1}, {tok_out, 16, 1}, {tok_len_out, 1, 1}};
  return run_pipelined(
      in, in_bpp, n, outs, 4,
      [&](int si, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
        EncodeOut o = make_out(out_dev);
        if (strict) {
          FixBuf& fb = t_fix_slot[si];
          fb.attach(&o);
          S2B_CUDA(cudaMemsetAsync(fb.d_count(), 0, sizeof(uint32_t), st));
        }
        cudaError_t e = launch_encode(kind, in_dev, cnt, o, st);
        if (e != cudaSuccess) return cuda_fail(e, "encode launch");
        if (strict) S2B_CUDA(t_fix_slot[si].fetch(kFixEager, st));
        return S2B_OK;
      },
      [&](int si, size_t off, size_t cnt, const void* in_dev, void** out_dev, cudaStream_t st) {
        if (!strict) return S2B_OK;
        FixBuf& fb = t_fix_slot[si];
        if (fb.h_count() == 0) return S2B_OK;
        changed.clear();
        int crc = collect_patches(kind, in_dev, in_bpp, cnt, make_out(out_dev), st, fb, kFixEager, &changed);
        if (crc) return crc;
        // records carry chunk-local indices (index_base 0); the chunk's results are already in the caller's arrays
        for (const FixPatch& p : changed) {
          const size_t i = off + (size_t)p.index;
          for (int k = 0; k < num_levels; ++k) ids_out[(size_t)k * ids_stride + i] = cellid_parent(p.id, lv.level[k]);
          if (tok_level_index >= 0) {
            uint32_t w[4];
            const int len = cellid_to_token(cellid_parent(p.id, lv.level[tok_level_index]), w);
            memcpy(tok_out + 16 * i, w, 16);
            if (tok_len_out) tok_len_out[i] = (uint8_t)len;
          }
        }
        return S2B_OK;
      });
}

static int encode_dev(InputKind kind, const void* in, size_t n, const int* levels, int num_levels, uint64_t* ids_out,
                      uint8_t* flags_out, void* stream, int tok_level_index = -1, char* tok_out = nullptr,
                      uint8_t* tok_len_out = nullptr, bool strict = false) {
  LevelSpec lv;
  int rc = check_levels(levels, num_levels, &lv);
  if (rc) return rc;
  if (tok_level_index >= num_levels) return set_error(S2B_EINVAL, "token_level_index out of range");
  if (n == 0) return S2B_OK;
  if (!in || !ids_out || (tok_level_index >= 0 && !tok_out)) return set_error(S2B_EINVAL, "null input or output pointer");
  size_t align = kind == kInLatLngRad ? 16 : 8;
  if (((uintptr_t)in % align) || ((uintptr_t)ids_out % 8) || ((uintptr_t)tok_out % 16))
    return set_error(S2B_EINVAL, "device pointers must be %zu-byte (input), 8-byte (ids) and 16-byte (tokens) aligned", align);
  EncodeOut o;
  o.levels = lv;
  o.ids = ids_out;
  o.flags = flags_out;
  o.token_level_index = tok_level_index;
  o.tokens16 = tok_out;
  o.token_len = tok_len_out;
  o.force_cr = g_force_cr.load();
  cudaStream_t st = (cudaStream_t)stream;
  if (!(strict && strict_latlng(kind))) {
    cudaError_t e = launch_encode(kind, in, n, o, st);
    if (e != cudaSuccess) return cuda_fail(e, "encode launch");
    return S2B_OK;
  }
  // Strict: ONE launch whose list of boundary-sensitive points lands in mapped host memory; the host re-evaluates the
  // records with its libm as they arrive, so when the kernel ends only the stragglers are left; ids that differ are then
  // patched on the device.
  int dev = 0;
  S2B_CUDA(cudaGetDevice(&dev));
  const size_t in_bpp = kind == kInLatLngRad ? 16 : 8;
  LiveList& live = t_live;
  rc = live.ensure(dev, kLiveCap);
  if (rc) return rc;
  if (live.dirty) {   // a previous call failed half way: stale markers may be left
    S2B_CUDA(cudaStreamSynchronize(st));
    memset(live.h, 0, sizeof(FixRecord) * (size_t)live.cap);
    *live.h_count = 0;
    S2B_CUDA(cudaMemcpy(&live.base, live.d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    live.dirty = false;
  }
  o.ids_stride = n;
  o.index_base = 0;
  o.fix_list = live.d_alias; o.fix_count = live.d_count; o.fix_cap = live.cap; o.fix_base = live.base; o.fix_overflow = live.d_overflow;
  cudaError_t e = launch_encode(kind, in, n, o, st);
  if (e != cudaSuccess) return cuda_fail(e, "encode launch");
  S2B_CUDA(cudaEventRecord(live.ev, st));
  std::vector<FixPatch> changed;
  std::vector<FixRecord>& seen = live.seen;   // the records as consumed (id = the device's id) ...
  std::vector<uint64_t>& host_id = live.host_id;   // ... and the id this host's libm gives
  seen.clear(); host_id.clear();
  uint32_t k = 0;
  auto consume_ready = [&](uint32_t limit) {   // records [k, limit) whose marker has arrived, in slot order
    while (k < limit) {
      volatile FixRecord* r = live.h + k;
      const uint64_t id = r->id;
      if (id == 0) break;
      __atomic_thread_fence(__ATOMIC_ACQUIRE);
      FixRecord rec;
      rec.index = r->index; rec.lat = r->lat; rec.lng = r->lng; rec.id = id;
      r->id = 0;                               // the slot is clean for the next call
      seen.push_back(rec);
      host_id.push_back(leaf_id_host_libm(rec.lat, rec.lng));
      ++k;
    }
  };
  live.dirty = true;
  for (;;) {
    consume_ready(live.cap);
    cudaError_t q = cudaEventQuery(live.ev);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) return cuda_fail(q, "encode kernel");
  }
  const uint32_t early = k;
  // The kernel has completed, so every record it wrote is in host memory: the list ends at the first empty slot. No copy
  // of the device counter is needed unless an append found the list full (the kernel then set the overflow word).
  const bool overflowed = *(volatile uint32_t*)live.h_count != 0;
  if (!overflowed) {
    consume_ready(live.cap);
    const uint32_t c = k;
    live.base += c;
    // Records consumed while the kernel was running were read behind a marker; now that the launch has completed their
    // payload is visible by definition — make sure it is what was used (it always is when stores land in order).
    for (uint32_t s = 0; s < early; ++s) {
      const FixRecord& now = live.h[s];
      if (now.index != seen[s].index || now.lat != seen[s].lat || now.lng != seen[s].lng) {
        seen[s].index = now.index; seen[s].lat = now.lat; seen[s].lng = now.lng;
        host_id[s] = leaf_id_host_libm(now.lat, now.lng);
      }
    }
    for (uint32_t s = 0; s < c; ++s)
      if (host_id[s] != seen[s].id) changed.push_back(FixPatch{seen[s].index, host_id[s]});
    g_strict_reevaluated.fetch_add(c, std::memory_order_relaxed);
    g_strict_changed.fetch_add(changed.size(), std::memory_order_relaxed);
    live.dirty = false;
  } else {
    // More sensitive points than the live list holds (input concentrated on cell boundaries): wipe it, resynchronise the
    // slot base with the device counter and redo the batch piecewise through a device-side list (collect_patches'
    // fallback), which counts and patches everything.
    for (uint32_t s = 0; s < live.cap; ++s) live.h[s].id = 0;
    *live.h_count = 0;
    S2B_CUDA(cudaMemcpy(&live.base, live.d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    live.dirty = false;
    changed.clear();
    FixBuf& fb = t_fix_dev;
    const size_t want = n / 128 < (1u << 16) ? (1u << 16) : (n / 128 > (1u << 24) ? (1u << 24) : n / 128);
    rc = fb.ensure(dev, (uint32_t)want);
    if (rc) return rc;
    *(volatile uint32_t*)fb.h = 0xFFFFFFFFu;   // "overflowed": collect_patches goes piecewise
    EncodeOut q = o;
    q.fix_list = nullptr; q.fix_count = nullptr; q.fix_cap = 0; q.fix_base = 0;
    rc = collect_patches(kind, in, in_bpp, n, q, st, fb, 0, &changed);
    if (rc) return rc;
  }
  FixBuf& pb = t_fix_dev;
  for (size_t off = 0; off < changed.size(); off += (1u << 20)) {
    const uint32_t m = (uint32_t)(changed.size() - off < (1u << 20) ? changed.size() - off : (1u << 20));
    pb.device = dev;
    rc = pb.ensure_patches(m);
    if (rc) return rc;
    S2B_CUDA(cudaMemcpyAsync(pb.d_patch, changed.data() + off, sizeof(FixPatch) * m, cudaMemcpyHostToDevice, st));
    EncodeOut whole = o;
    whole.index_base = 0;
    e = launch_encode_patch(pb.d_patch, m, whole, n, st);
    if (e != cudaSuccess) return cuda_fail(e, "patch launch");
    S2B_CUDA(cudaStreamSynchronize(st));   // `changed` is pageable host memory
  }
  return S2B_OK;
}

}  // namespace s2b

using namespace s2b;

extern "C" {

const char* s2b_last_error(void) { return t_err; }
const char* s2b_version(void) { return "s2b 0.1 (sm_100a)"; }
int s2b_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
uint64_t s2b_launch_count(void) { return g_launches.load(); }
int s2b_diag_force_correctly_rounded(int on) { return g_force_cr.exchange(on ? 1 : 0); }
int s2b_set_latlng_mode(int mode) {
  if (mode != S2B_LATLNG_HOST_LIBM && mode != S2B_LATLNG_CORRECTLY_ROUNDED) return set_error(S2B_EINVAL, "unknown lat/lng mode %d", mode);
  return g_latlng_mode.exchange(mode);
}
int s2b_latlng_strict_stats(uint64_t* reevaluated_out, uint64_t* changed_out, int reset) {
  if (reevaluated_out) *reevaluated_out = reset ? g_strict_reevaluated.exchange(0) : g_strict_reevaluated.load();
  if (changed_out) *changed_out = reset ? g_strict_changed.exchange(0) : g_strict_changed.load();
  if (reset && !reevaluated_out) g_strict_reevaluated.store(0);
  if (reset && !changed_out) g_strict_changed.store(0);
  return S2B_OK;
}
int s2b_encode_latlng_strict_dev(int input_kind, const void* latlng_dev, size_t n, const int* levels, int num_levels,
                                 uint64_t* ids_out_dev, uint8_t* boundary_flag_out_dev, int token_level_index,
                                 char* tokens16_out_dev, uint8_t* len_out_dev, void* stream) {
  if (input_kind != 1 && input_kind != 2) return set_error(S2B_EINVAL, "input_kind must be 1 (lat/lng radians) or 2 (lat/lng E7)");
  return encode_dev(input_kind == 1 ? kInLatLngRad : kInLatLngE7, latlng_dev, n, levels, num_levels, ids_out_dev,
                    boundary_flag_out_dev, stream, token_level_index < 0 ? -1 : token_level_index, tokens16_out_dev, len_out_dev, true);
}

int s2b_diag_filter_deviation_dev(int input_kind, const double* in_dev, size_t n, double* max_deviation_out, uint64_t* unflagged_face_flips_out, void* stream) {
  if (!max_deviation_out || !unflagged_face_flips_out || (n && !in_dev)) return set_error(S2B_EINVAL, "null argument");
  if (input_kind != 0 && input_kind != 1) return set_error(S2B_EINVAL, "input_kind must be 0 (xyz) or 1 (lat/lng radians)");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* d = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&d, 16, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  e = cudaMemsetAsync(d, 0, 16, st);
  if (e == cudaSuccess) e = launch_filter_deviation(input_kind == 0 ? kInXYZ : kInLatLngRad, in_dev, n, d, st);
  unsigned long long h[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFreeAsync(d, st);
  if (e != cudaSuccess) return cuda_fail(e, "filter deviation");
  memcpy(max_deviation_out, &h[0], 8);
  *unflagged_face_flips_out = h[1];
  return S2B_OK;
}

int s2b_encode_points(const double* xyz, size_t n, int level, uint64_t* ids_out) {
  return encode_host(kInXYZ, xyz, 24, n, &level, 1, ids_out, nullptr);
}
int s2b_encode_points_dev(const double* xyz_dev, size_t n, int level, uint64_t* ids_out_dev, void* stream) {
  return encode_dev(kInXYZ, xyz_dev, n, &level, 1, ids_out_dev, nullptr, stream);
}
int s2b_encode_latlng(const double* ll, size_t n, int level, uint64_t* ids_out, uint8_t* flag_out) {
  return encode_host(kInLatLngRad, ll, 16, n, &level, 1, ids_out, flag_out);
}
int s2b_encode_latlng_dev(const double* ll_dev, size_t n, int level, uint64_t* ids_out_dev, uint8_t* flag_out_dev, void* stream) {
  return encode_dev(kInLatLngRad, ll_dev, n, &level, 1, ids_out_dev, flag_out_dev, stream);
}
int s2b_encode_latlng_e7(const int32_t* e7, size_t n, int level, uint64_t* ids_out, uint8_t* flag_out) {
  return encode_host(kInLatLngE7, e7, 8, n, &level, 1, ids_out, flag_out);
}
int s2b_encode_latlng_e7_dev(const int32_t* e7_dev, size_t n, int level, uint64_t* ids_out_dev, uint8_t* flag_out_dev, void* stream) {
  return encode_dev(kInLatLngE7, e7_dev, n, &level, 1, ids_out_dev, flag_out_dev, stream);
}
int s2b_encode_latlng_levels(const double* ll, size_t n, const int* levels, int num_levels, uint64_t* ids_out) {
  return encode_host(kInLatLngRad, ll, 16, n, levels, num_levels, ids_out, nullptr);
}
int s2b_encode_latlng_levels_dev(const double* ll_dev, size_t n, const int* levels, int num_levels, uint64_t* ids_out_dev, void* stream) {
  return encode_dev(kInLatLngRad, ll_dev, n, levels, num_levels, ids_out_dev, nullptr, stream);
}
int s2b_encode_points_levels_dev(const double* xyz_dev, size_t n, const int* levels, int num_levels, uint64_t* ids_out_dev, void* stream) {
  return encode_dev(kInXYZ, xyz_dev, n, levels, num_levels, ids_out_dev, nullptr, stream);
}
int s2b_encode_latlng_levels_tokens(const double* ll, size_t n, const int* levels, int num_levels, uint64_t* ids_out,
                                    int token_level_index, char* tokens16_out, uint8_t* len_out) {
  if (token_level_index < 0) return set_error(S2B_EINVAL, "token_level_index must be >= 0");
  return encode_host(kInLatLngRad, ll, 16, n, levels, num_levels, ids_out, nullptr, token_level_index, tokens16_out, len_out);
}
int s2b_encode_latlng_levels_tokens_dev(const double* ll_dev, size_t n, const int* levels, int num_levels, uint64_t* ids_out_dev,
                                        int token_level_index, char* tokens16_out_dev, uint8_t* len_out_dev, void* stream) {
  if (token_level_index < 0) return set_error(S2B_EINVAL, "token_level_index must be >= 0");
  return encode_dev(kInLatLngRad, ll_dev, n, levels, num_levels, ids_out_dev, nullptr, stream, token_level_index,
                    tokens16_out_dev, len_out_dev);
}

int s2b_cellid_parent(const uint64_t* ids, size_t n, int level, uint64_t* out) {
  if (level < 0 || level > 30) return set_error(S2B_EINVAL, "level %d out of range [0,30]", level);
  if (n == 0) return S2B_OK;
  if (!ids || !out) return set_error(S2B_EINVAL, "null pointer");
  OutSpec outs[1] = {{out, 8, 1}};
  return run_pipelined(ids, 8, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    cudaError_t e = launch_parent((const uint64_t*)in_dev, cnt, level, (uint64_t*)out_dev[0], st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "parent launch");
  });
}
int s2b_cellid_parent_dev(const uint64_t* ids_dev, size_t n, int level, uint64_t* out_dev, void* stream) {
  if (level < 0 || level > 30) return set_error(S2B_EINVAL, "level %d out of range [0,30]", level);
  if (n == 0) return S2B_OK;
  if (!ids_dev || !out_dev) return set_error(S2B_EINVAL, "null pointer");
  cudaError_t e = launch_parent(ids_dev, n, level, out_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "parent launch");
}

int s2b_cellid_to_token(const uint64_t* ids, size_t n, char* tokens16_out, uint8_t* len_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !tokens16_out) return set_error(S2B_EINVAL, "null pointer");
  OutSpec outs[2] = {{tokens16_out, 16, 1}, {len_out, 1, 1}};
  return run_pipelined(ids, 8, n, outs, 2, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    cudaError_t e = launch_to_token((const uint64_t*)in_dev, cnt, (char*)out_dev[0], len_out ? (uint8_t*)out_dev[1] : nullptr, st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "token launch");
  });
}
int s2b_cellid_to_token_dev(const uint64_t* ids_dev, size_t n, char* tok_dev, uint8_t* len_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev || !tok_dev) return set_error(S2B_EINVAL, "null pointer");
  if ((uintptr_t)tok_dev % 16) return set_error(S2B_EINVAL, "tokens16_out_dev must be 16-byte aligned");
  cudaError_t e = launch_to_token(ids_dev, n, tok_dev, len_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "token launch");
}

int s2b_cellid_edge_neighbors(const uint64_t* ids, size_t n, uint64_t* neighbors4_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !neighbors4_out) return set_error(S2B_EINVAL, "null pointer");
  OutSpec outs[1] = {{neighbors4_out, 32, 1}};
  return run_pipelined(ids, 8, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    cudaError_t e = launch_edge_neighbors((const uint64_t*)in_dev, cnt, (uint64_t*)out_dev[0], st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "edge_neighbors launch");
  });
}
int s2b_cellid_edge_neighbors_dev(const uint64_t* ids_dev, size_t n, uint64_t* neighbors4_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev || !neighbors4_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if ((uintptr_t)neighbors4_out_dev % 16) return set_error(S2B_EINVAL, "neighbors4_out_dev must be 16-byte aligned");
  cudaError_t e = launch_edge_neighbors(ids_dev, n, neighbors4_out_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "edge_neighbors launch");
}
int s2b_cellid_to_point(const uint64_t* ids, size_t n, double* xyz_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !xyz_out) return set_error(S2B_EINVAL, "null pointer");
  OutSpec outs[1] = {{xyz_out, 24, 1}};
  return run_pipelined(ids, 8, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    cudaError_t e = launch_to_point((const uint64_t*)in_dev, cnt, (double*)out_dev[0], st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "to_point launch");
  });
}
int s2b_cellid_to_point_dev(const uint64_t* ids_dev, size_t n, double* xyz_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev || !xyz_out_dev) return set_error(S2B_EINVAL, "null pointer");
  cudaError_t e = launch_to_point(ids_dev, n, xyz_out_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "to_point launch");
}
int s2b_cellid_to_latlng(const uint64_t* ids, size_t n, double* latlng_rad_out) {
  if (n == 0) return S2B_OK;
  if (!ids || !latlng_rad_out) return set_error(S2B_EINVAL, "null pointer");
  OutSpec outs[1] = {{latlng_rad_out, 16, 1}};
  return run_pipelined(ids, 8, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    cudaError_t e = launch_to_latlng((const uint64_t*)in_dev, cnt, (double*)out_dev[0], st);
    return e == cudaSuccess ? S2B_OK : cuda_fail(e, "to_latlng launch");
  });
}
int s2b_cellid_to_latlng_dev(const uint64_t* ids_dev, size_t n, double* latlng_rad_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!ids_dev || !latlng_rad_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if ((uintptr_t)latlng_rad_out_dev % 16) return set_error(S2B_EINVAL, "latlng_rad_out_dev must be 16-byte aligned");
  cudaError_t e = launch_to_latlng(ids_dev, n, latlng_rad_out_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "to_latlng launch");
}
int s2b_cellid_from_token_dev(const char* tokens16_dev, const uint8_t* len_dev, size_t n, uint64_t* ids_out_dev, void* stream) {
  if (n == 0) return S2B_OK;
  if (!tokens16_dev || !len_dev || !ids_out_dev) return set_error(S2B_EINVAL, "null pointer");
  if ((uintptr_t)tokens16_dev % 16) return set_error(S2B_EINVAL, "tokens16_dev must be 16-byte aligned");
  cudaError_t e = launch_from_token(tokens16_dev, len_dev, n, ids_out_dev, (cudaStream_t)stream);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "from_token launch");
}
int s2b_cellid_from_token(const char* tokens16, const uint8_t* len, size_t n, uint64_t* ids_out) {
  if (n == 0) return S2B_OK;
  if (!tokens16 || !len || !ids_out) return set_error(S2B_EINVAL, "null pointer");
  // two input arrays: stage the lengths next to the tokens (17 bytes per token) through one pipelined pass each
  char* d_tok = nullptr; uint8_t* d_len = nullptr; uint64_t* d_ids = nullptr;
  int rc = S2B_OK;
  cudaError_t e = cudaMalloc(&d_tok, 16 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_len, n);
  if (e == cudaSuccess) e = cudaMalloc(&d_ids, 8 * n);
  if (e == cudaSuccess) e = cudaMemcpy(d_tok, tokens16, 16 * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_len, len, n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_from_token(d_tok, d_len, n, d_ids, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(ids_out, d_ids, 8 * n, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) rc = cuda_fail(e, "from_token");
  cudaFree(d_tok); cudaFree(d_len); cudaFree(d_ids);
  return rc;
}

// ---------------------------------------------------------------------------------------------------
// Host-pointer forms of the index queries: the points stream through the same three-slot pipeline.

int s2b_contains(const S2bIndex* ix, const double* xyz, size_t n, int model, uint8_t* out) {
  if (n && (!xyz || !out)) return set_error(S2B_EINVAL, "null pointer");
  if (n == 0) return s2b_contains_dev(ix, nullptr, 0, model, nullptr, nullptr);
  OutSpec outs[1] = {{out, 1, 1}};
  return run_pipelined(xyz, 24, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    return s2b_contains_dev(ix, (const double*)in_dev, cnt, model, (uint8_t*)out_dev[0], st);
  });
}

int s2b_index_locate_points(const S2bIndex* ix, const double* xyz, size_t n, int32_t* cell_out) {
  if (n && (!xyz || !cell_out)) return set_error(S2B_EINVAL, "null pointer");
  if (n == 0) return s2b_index_locate_points_dev(ix, nullptr, 0, nullptr, nullptr);
  OutSpec outs[1] = {{cell_out, 4, 1}};
  return run_pipelined(xyz, 24, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    return s2b_index_locate_points_dev(ix, (const double*)in_dev, cnt, (int32_t*)out_dev[0], st);
  });
}

int s2b_shape_contains(const S2bIndex* ix, int shape_id, const double* xyz, size_t n, int model, uint8_t* out) {
  if (n && (!xyz || !out)) return set_error(S2B_EINVAL, "null pointer");
  if (n == 0) return s2b_shape_contains_dev(ix, shape_id, nullptr, 0, model, nullptr, nullptr);
  OutSpec outs[1] = {{out, 1, 1}};
  return run_pipelined(xyz, 24, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    return s2b_shape_contains_dev(ix, shape_id, (const double*)in_dev, cnt, model, (uint8_t*)out_dev[0], st);
  });
}

int s2b_shape_histogram(const S2bIndex* ix, const double* xyz, size_t n, int model, uint64_t* counts) {
  if (!ix || !counts) return set_error(S2B_EINVAL, "null argument");
  int32_t S = 0;
  s2b_index_info(ix, nullptr, nullptr, nullptr, &S, nullptr);
  if (n && !xyz) return set_error(S2B_EINVAL, "null pointer");
  uint64_t* d_counts = nullptr;
  size_t bytes = sizeof(uint64_t) * (size_t)(S > 0 ? S : 1);
  S2B_CUDA(cudaMalloc(&d_counts, bytes));
  cudaError_t e = cudaMemset(d_counts, 0, bytes);  // default stream: ordered before the non-blocking slot streams by the sync below
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int rc = e == cudaSuccess ? S2B_OK : cuda_fail(e, "cudaMemset");
  if (rc == S2B_OK && n)
    rc = run_pipelined(xyz, 24, n, nullptr, 0, [&](int, const void* in_dev, size_t cnt, void**, cudaStream_t st) {
      return s2b_shape_histogram_dev(ix, (const double*)in_dev, cnt, model, d_counts, st);
    });
  if (rc == S2B_OK && S > 0) {
    e = cudaMemcpy(counts, d_counts, sizeof(uint64_t) * (size_t)S, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(counts)");
  }
  cudaFree(d_counts);
  return rc;
}

int s2b_containing_shapes(const S2bIndex* ix, const double* xyz, size_t n, int model, uint32_t* offsets_out,
                          int32_t* shape_ids_out, size_t capacity, size_t* total_out) {
  if (!ix || !offsets_out) return set_error(S2B_EINVAL, "null argument");
  if (n && !xyz) return set_error(S2B_EINVAL, "null pointer");
  if (total_out) *total_out = 0;
  offsets_out[0] = 0;
  if (n == 0) return S2B_OK;
  // pass 1: per-point counts land in offsets_out[1..n], then an exclusive scan on the host
  OutSpec outs[1] = {{offsets_out + 1, 4, 1}};
  int rc = run_pipelined(xyz, 24, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
    return s2b_containing_shapes_count_dev(ix, (const double*)in_dev, cnt, model, (uint32_t*)out_dev[0], st);
  });
  if (rc) return rc;
  uint64_t total = 0;
  for (size_t i = 1; i <= n; ++i) {
    total += offsets_out[i];
    if (total > 0xFFFFFFFFull) return set_error(S2B_ECAPACITY, "more than 2^32 (point, shape) pairs: split the batch");
    offsets_out[i] = (uint32_t)total;
  }
  if (total_out) *total_out = (size_t)total;
  if (total > capacity) return set_error(S2B_ECAPACITY, "shape_ids_out capacity %zu < %llu pairs", capacity, (unsigned long long)total);
  if (total == 0) return S2B_OK;
  if (!shape_ids_out) return set_error(S2B_EINVAL, "null shape_ids_out");
  // pass 2: chunk by chunk, offsets rebased to the chunk, ids written to a device buffer and copied to their place
  size_t chunk = n < kChunk ? n : kChunk;
  uint32_t* d_off = nullptr;
  int32_t* d_ids = nullptr;
  size_t ids_cap = 0;
  S2B_CUDA(cudaMalloc(&d_off, sizeof(uint32_t) * chunk));
  std::vector<uint32_t> rebased(chunk);
  rc = t_ws.ensure(/*dev*/ t_ws.device, 24 * chunk, 256);
  cudaStream_t st = t_ws.slot[0].stream;
  for (size_t off = 0; off < n && rc == S2B_OK; off += chunk) {
    size_t cnt = n - off < chunk ? n - off : chunk;
    uint32_t base = offsets_out[off];
    size_t pairs = offsets_out[off + cnt] - base;
    if (pairs == 0) continue;
    for (size_t i = 0; i < cnt; ++i) rebased[i] = offsets_out[off + i] - base;
    if (pairs > ids_cap) {
      if (d_ids) cudaFree(d_ids);
      d_ids = nullptr;
      cudaError_t e = cudaMalloc(&d_ids, sizeof(int32_t) * pairs);
      if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc(ids)"); break; }
      ids_cap = pairs;
    }
    cudaError_t e = cudaMemcpyAsync(t_ws.slot[0].in, xyz + 3 * off, 24 * cnt, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, rebased.data(), sizeof(uint32_t) * cnt, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync"); break; }
    rc = s2b_containing_shapes_fill_dev(ix, (const double*)t_ws.slot[0].in, cnt, model, d_off, d_ids, st);
    if (rc) break;
    e = cudaMemcpyAsync(shape_ids_out + base, d_ids, sizeof(int32_t) * pairs, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "fill copy"); break; }
  }
  cudaFree(d_off);
  if (d_ids) cudaFree(d_ids);
  return rc;
}

int s2b_cellunion_contains(const uint64_t* sorted_ids, size_t m, const double* xyz, size_t n, uint8_t* out) {
  if (n == 0) return S2B_OK;
  if ((m && !sorted_ids) || !xyz || !out) return set_error(S2B_EINVAL, "null pointer");
  for (size_t i = 1; i < m; ++i)
    if (sorted_ids[i - 1] >= sorted_ids[i]) return set_error(S2B_EINVAL, "sorted_ids must be strictly ascending (S2CellUnion::cell_ids())");
  uint64_t* d_ids = nullptr;
  S2B_CUDA(cudaMalloc(&d_ids, sizeof(uint64_t) * (m ? m : 1)));
  cudaError_t e = m ? cudaMemcpy(d_ids, sorted_ids, sizeof(uint64_t) * m, cudaMemcpyHostToDevice) : cudaSuccess;
  int rc = e == cudaSuccess ? S2B_OK : cuda_fail(e, "cudaMemcpy(ids)");
  if (rc == S2B_OK) {
    OutSpec outs[1] = {{out, 1, 1}};
    rc = run_pipelined(xyz, 24, n, outs, 1, [&](int, const void* in_dev, size_t cnt, void** out_dev, cudaStream_t st) {
      return s2b_cellunion_contains_dev(d_ids, m, (const double*)in_dev, cnt, (uint8_t*)out_dev[0], st);
    });
  }
  cudaFree(d_ids);
  return rc;
}

}  // extern "C"
