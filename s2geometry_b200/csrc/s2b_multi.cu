// s2b_multi.cu — the multi-GPU driver behind the C ABI (north star (4)): ONE process drives every GPU of
// the box. Point batches are sharded by contiguous range, the index is replicated on each GPU, and the only exchange is
// the all-reduce of the per-shape hit histogram (uint64[num_shape_ids], 80 KB for 10k polygons):
//   * NCCL (ncclCommInitAll over the replica devices, ncclAllReduce(uint64, sum) on each replica's stream), or
//   * the library's own reduction over NVLink peer memory: every GPU sums the peers' partial histograms with direct
//     peer loads in one small kernel, ordered behind the peers' query kernels by cross-device events — no second
//     library, no extra launch latency of a collective (S2B_REDUCE_PEER).
// Each GPU is driven by its own persistent worker thread (its thread-local pipeline workspace, streams and scratch stay
// warm between calls). Encoding shards the same way and needs no exchange at all.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstdio>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/s2b_capi.h"
#include "s2b_index_handle.h"
#include "s2b_internal.h"
#include "s2b_pipeline.h"

namespace s2b {

// ---------------------------------------------------------------------------------------------------
// One persistent worker thread per device ordinal.
class DeviceWorker {
 public:
  explicit DeviceWorker(int device) : device_(device), thread_([this] { loop(); }) {}
  // Runs job() on the worker (its current device = this worker's device) and returns immediately; wait() collects rc.
  void submit(std::function<int()> job) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = std::move(job);
    has_job_ = true; done_ = false;
    cv_.notify_all();
  }
  int wait(char* err, size_t err_len) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return done_; });
    if (rc_ != S2B_OK && err) snprintf(err, err_len, "%s", err_);
    return rc_;
  }

 private:
  void loop() {
    cudaSetDevice(device_);
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return has_job_; });
        job = std::move(job_);
        has_job_ = false;
      }
      int rc = job();
      {
        std::unique_lock<std::mutex> lk(mu_);
        rc_ = rc;
        if (rc != S2B_OK) snprintf(err_, sizeof err_, "device %d: %s", device_, s2b_last_error());
        done_ = true;
        cv_.notify_all();
      }
    }
  }
  int device_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool has_job_ = false, done_ = true;
  int rc_ = S2B_OK;
  char err_[512] = "";
  std::thread thread_;   // detached lifetime: lives until process exit
};

static DeviceWorker* worker_for(int device) {
  static std::mutex mu;
  static DeviceWorker* workers[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64) return nullptr;
  if (!workers[device]) workers[device] = new DeviceWorker(device);   // intentionally never destroyed (no join at exit)
  return workers[device];
}

// Runs fn(k) for k in [0, ndev) on the workers of devices[k], waits for all, returns the first error.
static int on_all_devices(const int* devices, int ndev, const std::function<int(int)>& fn) {
  static std::mutex one_call_at_a_time;   // the workers hold one job each
  std::lock_guard<std::mutex> lk(one_call_at_a_time);
  std::vector<DeviceWorker*> w(ndev);
  for (int k = 0; k < ndev; ++k) {
    w[k] = worker_for(devices[k]);
    if (!w[k]) return set_error(S2B_EINVAL, "device ordinal %d out of range", devices[k]);
  }
  for (int k = 0; k < ndev; ++k) w[k]->submit([&fn, k] { return fn(k); });
  int rc = S2B_OK;
  char err[512] = "";
  for (int k = 0; k < ndev; ++k) {
    char e[512] = "";
    int r = w[k]->wait(e, sizeof e);
    if (r != S2B_OK && rc == S2B_OK) { rc = r; snprintf(err, sizeof err, "%s", e); }
  }
  return rc == S2B_OK ? S2B_OK : set_error(rc, "%s", err);
}

static int check_devices(const int* devices, int ndev) {
  if (!devices || ndev < 1 || ndev > 64) return set_error(S2B_EINVAL, "devices must list between 1 and 64 CUDA ordinals");
  const int visible = s2b_device_count();
  for (int k = 0; k < ndev; ++k) {
    if (devices[k] < 0 || devices[k] >= visible) return set_error(visible ? S2B_EINVAL : S2B_ECUDA, "device %d is not one of the %d visible CUDA devices", devices[k], visible);
    for (int j = 0; j < k; ++j)
      if (devices[j] == devices[k]) return set_error(S2B_EINVAL, "device %d is listed twice", devices[k]);
  }
  return S2B_OK;
}

// Contiguous range of n units owned by shard k of ndev: sizes differ by at most one (the same rule as dist.shard_range).
static void shard_range(size_t n, int k, int ndev, size_t* lo, size_t* hi) {
  const size_t base = n / ndev, rem = n % ndev;
  *lo = k * base + ((size_t)k < rem ? (size_t)k : rem);
  *hi = *lo + base + ((size_t)k < rem ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------------
// NCCL, bound at run time (libs2b.so has no link-time dependency on it; a process that already loaded NCCL — e.g. through
// torch — shares that copy).
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static const Nccl& nccl() {
  static const Nccl api = [] {
    Nccl a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return a;
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.handle, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.handle, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.handle, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

// Own reduction over peer memory: out[i] = sum over replicas r of parts[r][i]; parts[r] live on the peers' GPUs and are
// read with ordinary loads over NVLink (peer access enabled at replication time).
struct PeerParts { const unsigned long long* p[16]; int n; };
__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerParts parts, unsigned long long* __restrict__ out, int count) {
  for (int i = blockIdx.x * 256 + threadIdx.x; i < count; i += gridDim.x * 256) {
    unsigned long long s = 0;
#pragma unroll 4
    for (int r = 0; r < parts.n; ++r) s += parts.p[r][i];
    out[i] = s;
  }
}

}  // namespace s2b

using namespace s2b;

// The replicas of one index: rep[0] is the index itself, the others are owned copies on the other devices.
struct S2bReplicaSet {
  std::vector<S2bIndex*> rep;
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;          // empty when NCCL is unavailable
  std::vector<cudaStream_t> streams;      // one per replica, on its device
  std::vector<cudaEvent_t> done;          // "partial histogram ready" per replica
  std::vector<cudaEvent_t> read_done;     // "this replica has finished reading its peers' partial histograms" (peer reduction)
  std::vector<unsigned long long*> part, total;   // per replica: partial and reduced histogram (num_shape_ids each)
  bool peer_access = false;               // every pair of replica devices can load from each other
  int reduce_mode = S2B_REDUCE_NCCL;
};

void s2b_replica_set_destroy(S2bReplicaSet* g) {
  if (!g) return;
  int cur = 0;
  cudaGetDevice(&cur);
  for (size_t k = 0; k < g->rep.size(); ++k) {
    cudaSetDevice(g->devices[k]);
    if (k < g->comms.size() && g->comms[k]) nccl().CommDestroy(g->comms[k]);
    if (k < g->streams.size() && g->streams[k]) cudaStreamDestroy(g->streams[k]);
    if (k < g->done.size() && g->done[k]) cudaEventDestroy(g->done[k]);
    if (k < g->read_done.size() && g->read_done[k]) cudaEventDestroy(g->read_done[k]);
    if (k < g->part.size() && g->part[k]) cudaFree(g->part[k]);
    if (k < g->total.size() && g->total[k]) cudaFree(g->total[k]);
    if (k > 0 && g->rep[k]) {
      if (g->rep[k]->blob) cudaFree(g->rep[k]->blob);
      if (g->rep[k]->edge_ids) cudaFree(g->rep[k]->edge_ids);
      delete g->rep[k];
    }
  }
  cudaSetDevice(cur);
  delete g;
}

extern "C" {

int s2b_index_replicate(S2bIndex* ix, const int* devices, int ndev) {
  if (!ix) return set_error(S2B_EINVAL, "null index");
  int rc = check_devices(devices, ndev);
  if (rc) return rc;
  if (ix->replicas) { s2b_replica_set_destroy(ix->replicas); ix->replicas = nullptr; }
  S2bReplicaSet* g = new (std::nothrow) S2bReplicaSet();
  if (!g) return set_error(S2B_ENOMEM, "host allocation failed");
  // replica 0 is the index itself; its device leads the list whether or not the caller listed it
  g->rep.push_back(ix);
  g->devices.push_back(ix->device);
  for (int k = 0; k < ndev; ++k)
    if (devices[k] != ix->device) { g->devices.push_back(devices[k]); g->rep.push_back(nullptr); }
  const int R = (int)g->devices.size();
  if (R > 16) { s2b_replica_set_destroy(g); return set_error(S2B_EINVAL, "at most 16 replicas"); }
  int cur = 0;
  cudaGetDevice(&cur);
  auto fail = [&](int code) { cudaSetDevice(cur); s2b_replica_set_destroy(g); return code; };
  const size_t S = (size_t)(ix->view.num_shape_ids > 0 ? ix->view.num_shape_ids : 1);
  g->streams.assign(R, nullptr); g->done.assign(R, nullptr); g->read_done.assign(R, nullptr);
  g->part.assign(R, nullptr); g->total.assign(R, nullptr);
  for (int k = 0; k < R; ++k) {
    cudaError_t e = cudaSetDevice(g->devices[k]);
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaSetDevice"));
    if (k > 0) {
      // copy the device blob GPU to GPU and rebase the view's pointers: the arrays sit at the same offsets
      S2bIndex* r = new (std::nothrow) S2bIndex(*ix);
      if (!r) return fail(set_error(S2B_ENOMEM, "host allocation failed"));
      r->replicas = nullptr; r->blob = nullptr; r->edge_ids = nullptr; r->device = g->devices[k];
      g->rep[k] = r;
      e = cudaMalloc(&r->blob, ix->blob_bytes);
      if (e == cudaSuccess) e = cudaMemcpyPeer(r->blob, r->device, ix->blob, ix->device, ix->blob_bytes);
      if (e != cudaSuccess) return fail(cuda_fail(e, "replicating the index"));
      const ptrdiff_t shift = (char*)r->blob - (char*)ix->blob;
      auto rebase = [shift](auto*& p) { p = (std::remove_reference_t<decltype(p)>)((char*)p + shift); };
      rebase(r->view.cell_ids); rebase(r->view.cells); rebase(r->view.clips); rebase(r->view.edges);
      rebase(r->view.lut); rebase(r->view.mixed); rebase(r->view.interior_info); rebase(r->view.id_first); rebase(r->view_dev);
      e = cudaMemcpy((void*)r->view_dev, &r->view, sizeof r->view, cudaMemcpyHostToDevice);
      if (e == cudaSuccess && ix->edge_ids) {
        e = cudaMalloc((void**)&r->edge_ids, sizeof(int32_t) * (ix->num_edges ? ix->num_edges : 1));
        if (e == cudaSuccess) e = cudaMemcpyPeer(r->edge_ids, r->device, ix->edge_ids, ix->device, sizeof(int32_t) * ix->num_edges);
      }
      if (e != cudaSuccess) return fail(cuda_fail(e, "replicating the index"));
    }
    e = cudaStreamCreateWithFlags(&g->streams[k], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->read_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->part[k], 8 * S);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->total[k], 8 * S);
    if (e != cudaSuccess) return fail(cuda_fail(e, "replica scratch"));
  }
  // peer access between every pair (for the library's own reduction); absence is not an error
  g->peer_access = R > 1;
  for (int a = 0; a < R && g->peer_access; ++a)
    for (int b = 0; b < R; ++b) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, g->devices[a], g->devices[b]);
      if (!can) { g->peer_access = false; break; }
      cudaSetDevice(g->devices[a]);
      cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[b], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      if (e != cudaSuccess) { cudaGetLastError(); g->peer_access = false; break; }
    }
  if (R > 1 && nccl().ok) {
    g->comms.assign(R, nullptr);
    ncclResult_t nr = nccl().CommInitAll(g->comms.data(), R, g->devices.data());
    if (nr != ncclSuccess) {
      g->comms.clear();
      if (!g->peer_access) return fail(set_error(S2B_ECUDA, "ncclCommInitAll: %s (and no peer access for the built-in reduction)", nccl().GetErrorString(nr)));
    }
  } else if (R > 1 && !g->peer_access) {
    return fail(set_error(S2B_ECUDA, "neither NCCL (libnccl.so.2) nor peer access between the devices is available for the histogram all-reduce"));
  }
  g->reduce_mode = g->comms.empty() ? S2B_REDUCE_PEER : S2B_REDUCE_NCCL;
  cudaSetDevice(cur);
  ix->replicas = g;
  return S2B_OK;
}

int s2b_index_num_replicas(const S2bIndex* ix) { return !ix ? 0 : (ix->replicas ? (int)ix->replicas->rep.size() : 1); }

int s2b_index_replica_devices(const S2bIndex* ix, int* devices_out, int capacity) {
  if (!ix || !devices_out) return set_error(S2B_EINVAL, "null argument");
  const int R = s2b_index_num_replicas(ix);
  if (capacity < R) return set_error(S2B_ECAPACITY, "capacity %d < %d replicas", capacity, R);
  for (int k = 0; k < R; ++k) devices_out[k] = ix->replicas ? ix->replicas->devices[k] : ix->device;
  return S2B_OK;
}

int s2b_index_set_reduce_mode(S2bIndex* ix, int mode) {
  if (!ix || !ix->replicas) return set_error(S2B_EINVAL, "the index has no replicas (s2b_index_replicate)");
  S2bReplicaSet* g = ix->replicas;
  if (mode == S2B_REDUCE_NCCL && g->comms.empty()) return set_error(S2B_EUNSUPPORTED, "NCCL is not available in this process");
  if (mode == S2B_REDUCE_PEER && !g->peer_access) return set_error(S2B_EUNSUPPORTED, "the replica devices cannot access each other's memory");
  if (mode != S2B_REDUCE_NCCL && mode != S2B_REDUCE_PEER) return set_error(S2B_EINVAL, "unknown reduce mode %d", mode);
  const int old = g->reduce_mode;
  g->reduce_mode = mode;
  return old;
}

// Enqueues the all-reduce of the replicas' partial histograms (part[k], complete once streams[k] reaches this point)
// into total[k] on every replica. Called from ONE host thread.
static int enqueue_allreduce(S2bReplicaSet* g) {
  const int R = (int)g->rep.size();
  const int S = g->rep[0]->view.num_shape_ids > 0 ? g->rep[0]->view.num_shape_ids : 1;
  int cur = 0;
  cudaGetDevice(&cur);
  int rc = S2B_OK;
  if (g->reduce_mode == S2B_REDUCE_NCCL) {
    const Nccl& n = nccl();
    n.GroupStart();
    for (int k = 0; k < R; ++k) {
      ncclResult_t nr = n.AllReduce(g->part[k], g->total[k], (size_t)S, ncclUint64, ncclSum, g->comms[k], g->streams[k]);
      if (nr != ncclSuccess && rc == S2B_OK) rc = set_error(S2B_ECUDA, "ncclAllReduce: %s", n.GetErrorString(nr));
    }
    ncclResult_t nr = n.GroupEnd();
    if (nr != ncclSuccess && rc == S2B_OK) rc = set_error(S2B_ECUDA, "ncclGroupEnd: %s", n.GetErrorString(nr));
  } else {
    PeerParts parts{};
    parts.n = R;
    for (int k = 0; k < R; ++k) parts.p[k] = g->part[k];
    for (int k = 0; k < R && rc == S2B_OK; ++k) {
      cudaSetDevice(g->devices[k]);
      cudaError_t e = cudaEventRecord(g->done[k], g->streams[k]);
      if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventRecord");
    }
    for (int k = 0; k < R && rc == S2B_OK; ++k) {
      cudaSetDevice(g->devices[k]);
      for (int j = 0; j < R; ++j)
        if (j != k) cudaStreamWaitEvent(g->streams[k], g->done[j], 0);
      peer_sum_kernel<<<(S + 255) / 256 < 64 ? (S + 255) / 256 : 64, 256, 0, g->streams[k]>>>(parts, g->total[k], S);
      count_launch();
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaEventRecord(g->read_done[k], g->streams[k]);
      if (e != cudaSuccess) rc = cuda_fail(e, "peer_sum launch");
    }
  }
  cudaSetDevice(cur);
  return rc;
}

// Before replica k's partial histogram is cleared for a new batch, every peer must have finished reading the previous one
// (peer reduction; an event that was never recorded counts as complete).
static void wait_for_peer_reads(S2bReplicaSet* g, int k) {
  for (int j = 0; j < (int)g->rep.size(); ++j)
    if (j != k) cudaStreamWaitEvent(g->streams[k], g->read_done[j], 0);
}

static S2bReplicaSet* group_of(const S2bIndex* ix) { return ix ? ix->replicas : nullptr; }

// Device-resident shards: shard k (n[k] points at xyz_dev[k]) lives on replica k's device. Enqueues, per replica, the
// histogram kernels on that replica's own stream, then the all-reduce; counts_out_dev[k] (nullable per entry) receives
// the GLOBAL histogram on device k. Returns after enqueueing unless `synchronize` is set.
int s2b_shape_histogram_multi_dev(const S2bIndex* ix, const double* const* xyz_dev, const size_t* n, int model,
                                  uint64_t* const* counts_out_dev, int synchronize) {
  S2bReplicaSet* g = group_of(ix);
  if (!g) return set_error(S2B_EINVAL, "the index has no replicas (s2b_index_replicate)");
  if (!xyz_dev || !n) return set_error(S2B_EINVAL, "null argument");
  const int R = (int)g->rep.size();
  const size_t S = (size_t)(ix->view.num_shape_ids > 0 ? ix->view.num_shape_ids : 1);
  // every GPU's launches are enqueued by its own worker thread (a round of the join is ~7 launches; from one thread the
  // enqueue of 8 GPUs would cost more than a small shard takes to run)
  int rc = on_all_devices(g->devices.data(), R, [&](int k) -> int {
    wait_for_peer_reads(g, k);
    S2B_CUDA(cudaMemsetAsync(g->part[k], 0, 8 * S, g->streams[k]));
    return n[k] ? s2b_shape_histogram_dev(g->rep[k], xyz_dev[k], n[k], model, (uint64_t*)g->part[k], g->streams[k]) : S2B_OK;
  });
  if (rc == S2B_OK && R > 1) rc = enqueue_allreduce(g);
  if (rc != S2B_OK) return rc;
  return on_all_devices(g->devices.data(), R, [&](int k) -> int {
    const unsigned long long* src = R > 1 ? g->total[k] : g->part[k];
    if (counts_out_dev && counts_out_dev[k]) S2B_CUDA(cudaMemcpyAsync(counts_out_dev[k], src, 8 * S, cudaMemcpyDeviceToDevice, g->streams[k]));
    if (synchronize) S2B_CUDA(cudaStreamSynchronize(g->streams[k]));
    return S2B_OK;
  });
}

// Host-pointer form: the batch is split into one contiguous range per replica; each GPU streams its range through its
// own chunked pipeline on its worker thread; the partial histograms are all-reduced and the result is read from GPU 0.
int s2b_shape_histogram_multi(const S2bIndex* ix, const double* xyz, size_t n, int model, uint64_t* counts) {
  S2bReplicaSet* g = group_of(ix);
  if (!g) return s2b_shape_histogram(ix, xyz, n, model, counts);
  if (!counts || (n && !xyz)) return set_error(S2B_EINVAL, "null pointer");
  const int R = (int)g->rep.size();
  const size_t S = (size_t)(ix->view.num_shape_ids > 0 ? ix->view.num_shape_ids : 1);
  int rc = on_all_devices(g->devices.data(), R, [&](int k) -> int {
    size_t lo, hi;
    shard_range(n, k, R, &lo, &hi);
    S2B_CUDA(cudaMemsetAsync(g->part[k], 0, 8 * S, g->streams[k]));
    S2B_CUDA(cudaStreamSynchronize(g->streams[k]));
    if (hi == lo) return S2B_OK;
    return run_pipelined(xyz + 3 * lo, 24, hi - lo, nullptr, 0, [&](int, const void* in_dev, size_t cnt, void**, cudaStream_t st) {
      return s2b_shape_histogram_dev(g->rep[k], (const double*)in_dev, cnt, model, (uint64_t*)g->part[k], st);
    });   // drained: part[k] is complete
  });
  if (rc) return rc;
  if (R > 1) {
    rc = enqueue_allreduce(g);
    if (rc) return rc;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(g->devices[0]);
  cudaError_t e = cudaMemcpyAsync(counts, R > 1 ? g->total[0] : g->part[0], 8 * (size_t)ix->view.num_shape_ids, cudaMemcpyDeviceToHost, g->streams[0]);
  for (int k = 0; k < R && e == cudaSuccess; ++k) {   // every rank's collective has completed before the call returns
    cudaSetDevice(g->devices[k]);
    e = cudaStreamSynchronize(g->streams[k]);
  }
  cudaSetDevice(cur);
  return e == cudaSuccess ? S2B_OK : cuda_fail(e, "histogram all-reduce");
}

// S2ContainsPointQuery::Contains for a host batch sharded over the replicas: no exchange, each GPU writes its range.
int s2b_contains_multi(const S2bIndex* ix, const double* xyz, size_t n, int model, uint8_t* out) {
  S2bReplicaSet* g = group_of(ix);
  if (!g) return s2b_contains(ix, xyz, n, model, out);
  if (n && (!xyz || !out)) return set_error(S2B_EINVAL, "null pointer");
  const int R = (int)g->rep.size();
  return on_all_devices(g->devices.data(), R, [&](int k) -> int {
    size_t lo, hi;
    shard_range(n, k, R, &lo, &hi);
    return hi == lo ? S2B_OK : s2b_contains(g->rep[k], xyz + 3 * lo, hi - lo, model, out + lo);
  });
}

// Encoding shards over any set of devices and needs no index and no exchange.
int s2b_encode_points_multi(const int* devices, int ndev, const double* xyz, size_t n, int level, uint64_t* ids_out) {
  int rc = check_devices(devices, ndev);
  if (rc) return rc;
  if (n && (!xyz || !ids_out)) return set_error(S2B_EINVAL, "null pointer");
  return on_all_devices(devices, ndev, [&](int k) -> int {
    size_t lo, hi;
    shard_range(n, k, ndev, &lo, &hi);
    return hi == lo ? S2B_OK : s2b_encode_points(xyz + 3 * lo, hi - lo, level, ids_out + lo);
  });
}

int s2b_encode_latlng_levels_tokens_multi(const int* devices, int ndev, const double* latlng_rad, size_t n, const int* levels,
                                          int num_levels, uint64_t* ids_out, int token_level_index, char* tokens16_out, uint8_t* len_out) {
  int rc = check_devices(devices, ndev);
  if (rc) return rc;
  if (n && (!latlng_rad || !ids_out)) return set_error(S2B_EINVAL, "null pointer");
  if (token_level_index >= 0 && !tokens16_out) return set_error(S2B_EINVAL, "null token output");
  return on_all_devices(devices, ndev, [&](int k) -> int {
    size_t lo, hi;
    shard_range(n, k, ndev, &lo, &hi);
    if (hi == lo) return S2B_OK;
    // the per-level id arrays stay n apart in the caller's buffer; this GPU fills [lo, hi) of each
    return encode_host(kInLatLngRad, latlng_rad + 2 * lo, 16, hi - lo, levels, num_levels, ids_out + lo, nullptr,
                       token_level_index < 0 ? -1 : token_level_index, tokens16_out ? tokens16_out + 16 * lo : nullptr,
                       len_out ? len_out + lo : nullptr, n);
  });
}

}  // extern "C"
