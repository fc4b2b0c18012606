"""Multi-GPU driver for the point workloads: one process per GPU (torch.distributed), points sharded by contiguous
range, the index / cell union replicated on every GPU, and NCCL used ONLY where the path has a real exchange step:
the all-reduce of the per-shape hit histogram (uint64[num_shape_ids], 80 KB for 10k polygons) and of scalar hit
counts. The reference has no multi-device path; every point is independent, so there is no halo,
shuffle or index sharding.
"""
import os

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialises torch.distributed from the torchrun environment (no-op for a single process).
    Returns (rank, world, device). backend defaults to nccl when CUDA is available, else gloo (CPU tests)."""
    rank, world, local_rank = env_rank_world()
    cuda = torch is not None and torch.cuda.is_available()
    device = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if cuda else "gloo")
        if backend == "nccl":
            dist.init_process_group(backend, device_id=device)
        else:
            dist.init_process_group(backend)
    return rank, world, device


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank):
    """Pins this process to the CPUs next to its GPU, so that pinned host buffers (first touch after the bind) and the
    staging threads sit on the GPU's PCIe root. Matters for the host-pointer (end-to-end) path when 8 ranks stream through
    host memory at once. Sources, in order: the GPU's sysfs numa_node, then NVML's ideal CPU affinity
    (nvmlDeviceGetCpuAffinity, what `nvidia-smi topo -m` prints). Returns a dict describing what was done and why —
    including the common VM case where the box exposes ONE NUMA node and there is nothing to bind to."""
    import glob
    info = {"bound": False, "numa_nodes": len(glob.glob("/sys/devices/system/node/node[0-9]*")), "cpus_allowed": len(os.sched_getaffinity(0))}
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        info["sysfs_numa_node"] = node
        if node >= 0:
            with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
                allowed = os.sched_getaffinity(0) & _cpulist(f.read())
            if allowed and allowed != os.sched_getaffinity(0):
                os.sched_setaffinity(0, allowed)
                info.update(bound=True, how="sysfs numa_node %d" % node, cpus_bound=len(allowed))
                return info
    except Exception as e:
        info["sysfs_error"] = repr(e)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0) & cpus
        info["nvml_affinity_cpus"] = len(cpus)
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, how="NVML ideal CPU affinity", cpus_bound=len(allowed))
            return info
        info["why"] = "the GPU's ideal CPU set is every allowed CPU (one NUMA node exposed): nothing to bind"
    except Exception as e:
        info["nvml_error"] = repr(e)
    return info


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of n units owned by `rank`; ranges partition [0, n) and differ in size by at most 1."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(t):
    """In-place SUM all-reduce (no-op for a single process). int64 holds the uint64 counts (sums stay < 2^63)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_max_(t):
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def barrier():
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def sharded_shape_histogram(local_histogram_fn, num_shape_ids, device):
    """Spatial join across ranks: local_histogram_fn(counts) accumulates this rank's shard into the int64 tensor
    `counts` (on the GPU: ContainsPointQuery.shape_histogram(points_shard, counts=counts)); the per-shape counts are
    then summed over ranks. Every rank returns the global histogram."""
    counts = torch.zeros(int(num_shape_ids), dtype=torch.int64, device=device)
    local_histogram_fn(counts)
    return all_reduce_sum_(counts)


def sharded_count(local_count, device):
    """Global number of hits given this rank's local count (e.g. S2CellUnion containment over a shard)."""
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device)
    return int(all_reduce_sum_(t).item())
