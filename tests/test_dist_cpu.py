"""world_size-2 gloo test of the multi-GPU host logic (sharding + histogram all-reduce). The per-shard compute is
done by the test-only host build of the kernels' logic (tests/hostsim); on a GPU box the same driver functions wrap
the CUDA calls (bench.py join section)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import ctypes
    import torch
    from s2geometry_b200 import dist as sdist, synth
    from s2geometry_b200.index import build_flat_index
    from tests import scenes
    from tests.test_index_cpu import hs_query
    r, w, dev = sdist.init(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    hs = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    soup = scenes.polygons_scene(200, 41, min_radius_km=300, max_radius_km=2500)
    flat = build_flat_index(soup)                      # replicated: every rank builds the same index
    pts = synth.sphere_points(n, 77)                   # the global point array (regenerated identically per rank)
    lo, hi = sdist.shard_range(n, rank, world)

    def local(counts):
        c, _ = hs_query(hs, flat, pts[lo:hi], 1, 2)
        counts += torch.from_numpy(c.astype(np.int64))

    total = sdist.sharded_shape_histogram(local, flat.num_shape_ids, dev)
    hits = sdist.sharded_count(int(hs_query(hs, flat, pts[lo:hi], 1, 0)[0].sum()), dev)
    if rank == 0:
        full, _ = hs_query(hs, flat, pts, 1, 2)
        any_hits = int(hs_query(hs, flat, pts, 1, 0)[0].sum())
        q.put((total.numpy().tolist(), full.astype(np.int64).tolist(), hits, any_hits, (lo, hi)))
    sdist.barrier()
    torch.distributed.destroy_process_group()


def test_shard_ranges_partition():
    from s2geometry_b200.dist import shard_range
    for n in (0, 1, 7, 100, 1_000_000_007):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_histogram_world2_gloo(hostsim):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n = 60001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, full, hits, any_hits, rng = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == full and sum(full) > 100
    assert hits == any_hits
    assert rng == (0, 30001)
