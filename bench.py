#!/usr/bin/env python3
"""bench.py — points/sec of the S2 hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

Headline workload (BASELINE.json configs[1]): 100M (lat,lng) pairs per GPU -> S2CellId at levels 12/20/30 plus
ToToken of the level-30 id — S2CellId(S2LatLng), parent(12), parent(20), ToToken() per point — ONE fused kernel launch
per step (s2b_encode_latlng_levels_tokens_dev).

  value   : whole-job points/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e     : same metric through the host-pointer C ABI call with pinned host buffers (H2D + kernel + D2H inside)
  roofline: the encode kernel against the measured HBM peak (algorithmic bytes/pt: 16 in + 3x8 ids + 16 token + 1 len = 57);
            "fp64" gives the second roofline the north star asks for (measured DFMA peak vs FP64 instructions/pt).
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref/libs2ref.so) on the box's host cores.

Multi-GPU (torchrun, one rank per GPU): points are independent, each rank encodes its own 100M (weak scaling), no
data-path collective; NCCL is used only for the barrier and the max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LEVELS = [12, 20, 30]
TOKEN_LEVEL_INDEX = 2
N_POINTS = 100_000_000
DEVICE_MAX_EDGES_PER_CELL = 1   # S2B_MAX_EDGES_PER_CELL_DEVICE (include/s2b_capi.h): granularity of device-resident indexes
BYTES_PER_POINT = 16 + 8 * len(LEVELS) + 16 + 1   # algorithmic HBM bytes per point of the fused kernel
# FP64-pipe instructions (DADD/DMUL/DFMA/DSETP) executed per point by encode_kernel<latlng>, measured with ncu's
# per-instruction counts (profiles/r2_encode_latlng_ncu.json, opcode mix in its "sass_per_point"); used only for the informational fp64 roofline.
FP64_INSTR_PER_POINT = 73.7       # DFMA 31.1 + DMUL 24.2 + DADD 8.4 + DSETP 10.0 executed per point (SASS counters of the same ncu capture)
DRAM_BYTES_PER_POINT = 56.44      # ncu dram__bytes_read.sum + dram__bytes_write.sum of the encode kernel, per point
PROFILE_SOURCE = "profiles/r2_encode_latlng_ncu.json"
METRIC = "points/sec (S2CellId encode, lat/lng -> levels 12/20/30 + ToToken)"
WORKLOAD = "configs[1]: 100M lat/lng (radians, uniform on sphere) per GPU -> S2CellId at levels 12/20/30 + ToToken(level 30)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # "under load" = upper half of samples
        sm_sorted = sorted(sm)
        return {"sm_mhz": float(np.median(sm_sorted[len(sm_sorted) // 2:])), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


STREAM_LATLNG = 2   # seed of the headline lat/lng stream (s2geometry_b200/synth.py: point i = f(seed, i), any host or GPU)


def workload_config(n):
    """The `config` object of BOTH arms (the reference arm must run the product arm's config)."""
    return {"workload": WORKLOAD, "points_per_gpu": n, "levels": LEVELS,
            "input": "counter-based stream %d, rank r encodes points [r*n, (r+1)*n) (s2geometry_b200/synth.py)" % STREAM_LATLNG,
            "l2_policy": "inputs (1.6 GB) and outputs (4.1 GB) per step exceed the 126 MB L2"}


def run_reference(args):
    """The reference's own CPU implementation of the step on this box's host cores: the same 100M-point stream the
    GPU arm encodes (counter-based generator: identical bits on the host), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import reflib  # the one place bench.py may execute oracle/ besides cpu_baseline
    from s2geometry_b200 import synth
    ref = reflib.load()
    threads = ref.hardware_threads()
    n = args.points
    ll = synth.stream_sphere_latlng(STREAM_LATLNG, 0, n)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ref.encode_latlng_levels_tokens(ll, LEVELS, TOKEN_LEVEL_INDEX, threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(n),   # the product arm's config, key for key
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "reference",
                         "sample": "all %d lat/lng points of stream %d per step (one GPU's share of the stream), %d threads; unmodified reference sources + abseil stand-in, -O2"
                                   % (n, STREAM_LATLNG, threads)},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


def cpu_baseline(ll_dev, ids_dev, tok_dev, len_dev):
    """Headline cpu_baseline leg: the reference on the host cores over the SAME points the GPU just encoded (copied back
    from the device: the stream is bit-reproducible, tests/test_synth_*), timed, and every id / token compared."""
    from oracle import reflib  # cpu_baseline leg
    ref = reflib.load()
    threads = ref.hardware_threads()
    n = ll_dev.shape[0]
    sample = n if threads >= 8 else min(n, 25_000_000)
    ll = ll_dev[:sample].cpu().numpy()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        w_ids, w_tok, w_len = ref.encode_latlng_levels_tokens(ll, LEVELS, TOKEN_LEVEL_INDEX, threads)
        best = min(best, time.perf_counter() - t0)
    mismatches = 0
    for k in range(len(LEVELS)):
        mismatches += int((ids_dev[k, :sample].cpu().numpy().view(np.uint64) != w_ids[k]).sum())
    mismatches += int((tok_dev[:sample].cpu().numpy() != w_tok).any(axis=1).sum()) + int((len_dev[:sample].cpu().numpy() != w_len).sum())
    return {"value": sample / best, "unit": "points/s", "cores": threads, "kind": "reference",
            "sample": "%d of %d lat/lng points, best of 3, %d threads (reference sources + abseil stand-in)" % (sample, n, threads),
            "parity_on_sample": mismatches == 0, "mismatches": mismatches,
            "compared": "ids at levels %s + tokens + token lengths of all %d sample points" % (LEVELS, sample)}


def pip_section(args, torch, dev, rank, world):
    """BASELINE configs[2]: 100M points vs ONE 10k-vertex loop in a shape index (S2ContainsPointQuery::Contains).
    The flat index is the committed fixture tests/golden/index_loop10k.npz (built and iterated by the reference, see
    tools/gen_golden.py). Two point sets: U uniform on the sphere (1% locate a cell) and B uniform in the loop's
    bounding cap (the reference's own benchmark design, s2loop_test.cc:1635-1666)."""
    from s2geometry_b200 import _lib, synth
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex
    from s2geometry_b200.index import build_flat_index
    from s2geometry_b200.shapes import ShapeSoup
    fx = np.load(os.path.join(ROOT, "tests", "golden", "index_loop10k.npz"))
    flat_default = FlatIndex(**{k: fx[k] for k in FlatIndex.FIELDS}, num_shape_ids=int(fx["num_shape_ids"]))
    # The device index is built with MutableS2ShapeIndex::Options::max_edges_per_cell = 1 instead of the CPU default 10:
    # query results do not depend on the granularity (tests), and finer cells mean fewer points reach the edge loops.
    # The reference-built default-granularity index (the committed fixture) is measured next to it.
    lsoup = ShapeSoup(); lsoup.add_polygon([fx["loop_vertices"]])
    flat = build_flat_index(lsoup.freeze(), DEVICE_MAX_EDGES_PER_CELL)
    ix = ShapeIndex(flat)
    ix_default = ShapeIndex(flat_default)
    query = ContainsPointQuery(ix, 1)
    n = args.points
    hbm_peak, _ = peaks()
    center = np.array([0.3, -0.5, 0.8]); radius = 0.17 * 1.25
    out = {"workload": "configs[2]: %d points vs one 10k-vertex loop (index with max_edges_per_cell=%d: %d cells, %d clipped edges), SEMI_OPEN" % (
               n, DEVICE_MAX_EDGES_PER_CELL, flat.num_cells, flat.num_edges),
           "bytes_per_point": 25}
    sets = {}
    # this rank's slices of the global streams 3 (uniform on the sphere) and 7 (uniform in the loop's bounding cap)
    pts_u = synth.stream_sphere_points_cuda(3, rank * n, n, dev)
    sets["uniform"] = pts_u
    sets["near"] = synth.stream_cap_points_cuda(7, rank * n, n, center, radius, dev)
    res = torch.empty(n, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    import ctypes
    out["default_granularity"] = {"workload": "the same loop in the reference-built index with the default max_edges_per_cell=10 (%d cells, %d clipped edges)" % (
        flat_default.num_cells, flat_default.num_edges)}
    for index, dest in ((ix, out), (ix_default, out["default_granularity"])):
        for name, pts in sets.items():
            def step():
                _lib.check(lib.s2b_contains_dev(index._h, ctypes.c_void_p(pts.data_ptr()), n, 1, ctypes.c_void_p(res.data_ptr()),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            ach = 25 * n / (ms * 1e-3) / 1e9
            dest[name] = {"points_per_s": n / (ms * 1e-3), "ms_per_step": ms, "inside_fraction": float(res.float().mean().item()),
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak}}
    # end to end through the host-pointer call (pinned input, pinned output), U set
    h_pts = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); h_pts.copy_(pts_u)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    a_pts, a_out = h_pts.numpy(), h_out.numpy()
    _lib.check(lib.s2b_contains(ix._h, ctypes.c_void_p(a_pts.ctypes.data), n, 1, ctypes.c_void_p(a_out.ctypes.data)))
    t0 = time.perf_counter()
    _lib.check(lib.s2b_contains(ix._h, ctypes.c_void_p(a_pts.ctypes.data), n, 1, ctypes.c_void_p(a_out.ctypes.data)))
    dt = time.perf_counter() - t0
    out["e2e_uniform"] = {"points_per_s": n / dt, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": n}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import reflib  # cpu_baseline leg
            from s2geometry_b200.shapes import ShapeSoup
            ref = reflib.load()
            soup = ShapeSoup(); soup.add_polygon([fx["loop_vertices"]])
            ri = ref.index_create(soup.freeze())
            threads = ref.hardware_threads()
            cb = {}
            for name, pts in sets.items():
                sample = pts[:20_000_000].cpu().numpy()
                best = 1e30
                for _ in range(2):
                    t0 = time.perf_counter(); r = ri.contains(sample, 1, threads); best = min(best, time.perf_counter() - t0)
                cb[name] = {"value": len(sample) / best, "unit": "points/s", "cores": threads, "kind": "reference",
                            "sample": "first 20M points of the set, best of 2",
                            "parity_on_sample": bool((r == query.contains(pts[:20_000_000]).cpu().numpy()).all())}
            out["cpu_baseline"] = cb
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def encode_xyz_section(args, torch, dev, rank, world):
    """BASELINE configs[0] on the GPU: S2Points -> level-30 S2CellId (32 algorithmic bytes per point)."""
    from s2geometry_b200 import cellid, synth
    n = args.points
    pts = synth.stream_sphere_points_cuda(1, rank * n, n, dev)
    ids = None
    for _ in range(3):
        ids = cellid.from_points(pts, 30)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ids = cellid.from_points(pts, 30)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    hbm_peak, _ = peaks()
    ach = 32 * n / (ms * 1e-3) / 1e9
    out = {"workload": "configs[0]: %d uniform S2Points -> level-30 S2CellId (per GPU)" % n, "points_per_s": n / (ms * 1e-3), "ms_per_step": ms,
           "bytes_per_point": 32, "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import reflib  # cpu_baseline leg
            ref = reflib.load()
            sample = pts[:20_000_000].cpu().numpy()
            t0 = time.perf_counter(); want = ref.encode_points(sample, 30, ref.hardware_threads()); dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": len(sample) / dt, "unit": "points/s", "cores": ref.hardware_threads(), "kind": "reference",
                                   "sample": "first 20M points", "parity_on_sample": bool((ids[:20_000_000].cpu().numpy().view(np.uint64) == want).all())}
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def radius_join_section(args, torch, dev, rank, world):
    """SURVEY 8(f) rank 3, the radius join of the reference's doc/examples/point_index.cc: an S2PointIndex of 10M uniform
    points, targets = this rank's share of the uniform point stream, count of index points within 10 km (Earth) of every
    target = S2ClosestPointQuery with max_distance; plus FindClosestPoint (nearest index point) for every target."""
    from s2geometry_b200 import synth
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    m = 10_000_000
    n = min(args.points, 50_000_000)
    radius = 10.0 / 6371.01
    index_pts = synth.stream_sphere_points_cuda(77, 0, m, dev)
    t0 = time.perf_counter()
    ix = PointIndex(index_pts)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    targets = synth.stream_sphere_points_cuda(200, rank * n, n, dev)
    q = ClosestPointQuery(ix, radius)
    qn = ClosestPointQuery(ix)

    def timed(f):
        for _ in range(2):
            r = f()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(2, args.steps // 2)):
            r = f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / max(2, args.steps // 2), r
    ms_c, counts = timed(lambda: q.count(targets))
    ms_n, (data, l2) = timed(lambda: qn.closest(targets))
    pairs = int(counts.sum().item())
    out = {"workload": "radius join: %d targets per GPU vs an S2PointIndex of %d uniform points, max_distance 10 km (S2ClosestPointQuery, PointTarget)" % (n, m),
           "index_build_s": build_s, "count": {"targets_per_s": n / (ms_c * 1e-3), "ms_per_step": ms_c, "pairs_found": pairs},
           "closest": {"targets_per_s": n / (ms_n * 1e-3), "ms_per_step": ms_n}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import reflib  # cpu_baseline leg
            ref = reflib.load()
            ri = ref.point_index(index_pts.cpu().numpy())
            k = 2_000_000
            sample = targets[:k].cpu().numpy()
            t0 = time.perf_counter(); want = ri.count(sample, radius, ref.hardware_threads()); dt = time.perf_counter() - t0
            t0 = time.perf_counter(); wd, wl = ri.closest(sample, -1.0, ref.hardware_threads()); dt2 = time.perf_counter() - t0
            out["cpu_baseline"] = {"count": {"value": k / dt, "unit": "targets/s"}, "closest": {"value": k / dt2, "unit": "targets/s"},
                                   "cores": ref.hardware_threads(), "kind": "reference", "sample": "first 2M targets, S2ClosestPointQuery<int>, one query object per thread",
                                   "parity_on_sample": bool((counts[:k].cpu().numpy().astype(np.uint32) == want).all() and (l2[:k].cpu().numpy() == wl).all())}
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def cell_index_join_section(args, torch, dev, rank, world):
    """S2CellIndex join. 10k labelled coverings (a cell and its AppendAllNeighbors ring, built with the
    library's own cell algebra) indexed in an S2CellIndex; every point of this rank's share of the stream reports the
    (cell, label) pairs VisitIntersectingCells gives for its leaf cell -> per-label hit histogram (fused encode + lookup),
    all-reduced over the ranks. Two scenes: dense (cells of levels 5..11, ~2.9 pairs per point) and sparse (levels 9..13)."""
    import torch.distributed as dist
    from s2geometry_b200 import cellid, synth
    from s2geometry_b200.cellindex import CellIndex
    num_labels = 10000
    centers = synth.sphere_points(num_labels, 55)
    leaves = cellid.from_points(centers, 30)
    n = args.points
    pts = synth.stream_sphere_points_cuda(300, rank * n, n, dev)
    hbm_peak, _ = peaks()

    def scene(level_lo, level_hi):
        rng = np.random.default_rng(123)
        label_level = rng.integers(level_lo, level_hi + 1, num_labels)
        ids, labels = [], []
        for level in range(level_lo, level_hi + 1):
            sel = np.nonzero(label_level == level)[0]
            c = cellid.parent(leaves[sel], level)
            nb, cnt = cellid.all_neighbors(c, level)
            for k, lab in enumerate(sel):
                cells = np.concatenate([c[k:k + 1], nb[k, :cnt[k]]])
                ids.append(cells); labels.append(np.full(len(cells), lab, np.int32))
        return np.concatenate(ids).astype(np.uint64), np.concatenate(labels).astype(np.int32)

    def run(level_lo, level_hi, with_cpu):
        ids, labels = scene(level_lo, level_hi)
        t0 = time.perf_counter()
        ix = CellIndex(ids, labels)
        build_s = time.perf_counter() - t0

        def step():
            counts = ix.point_histogram(pts)
            if world > 1:
                dist.all_reduce(counts)
            return counts
        for _ in range(3):
            counts = step()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            counts = step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ach = 24 * n / (ms * 1e-3) / 1e9
        out = {"workload": "S2CellIndex join: %d uniform points per GPU vs %d (cell, label) pairs of %d labels (cells of levels %d..%d), per-label hit histogram"
                           % (n, len(ids), num_labels, level_lo, level_hi),
               "points_per_s": world * n / (ms * 1e-3), "ms_per_step": ms, "index_build_s": build_s, "pairs_found": int(counts.sum().item()),
               "bytes_per_point": 24, "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "note": "per GPU"}}
        if with_cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import reflib  # cpu_baseline leg
                ref = reflib.load()
                k = min(n, 20_000_000)
                sample = pts[:k].cpu().numpy()
                ri = reflib.RefCellIndex(ref, ids, labels)
                t0 = time.perf_counter(); want = ri.point_histogram(sample, ref.hardware_threads()); dt = time.perf_counter() - t0
                got = ix.point_histogram(pts[:k]).cpu().numpy().astype(np.uint64)
                out["cpu_baseline"] = {"value": k / dt, "unit": "points/s", "cores": ref.hardware_threads(), "kind": "reference",
                                       "sample": "first %d points, S2CellIndex::VisitIntersectingCells per point" % k, "parity_on_sample": bool((got == want).all())}
            except Exception as e:
                out["cpu_baseline"] = {"unavailable": str(e)}
        ix.close()
        return out
    out = run(5, 11, True)
    out["sparse"] = run(9, 13, True)
    if rank == 0 and world == 1:
        # Cell targets through the host-pointer API (copies included): VisitIntersectingCells for 10M cells of levels 8..16.
        ids, labels = scene(5, 11)
        ix = CellIndex(ids, labels)
        m = min(n, 10_000_000)
        tl = cellid.from_points(pts[:m], 30).cpu().numpy().view(np.uint64)
        targets = tl.copy()
        lv = np.random.default_rng(5).integers(8, 17, m).astype(np.uint64)
        lsb = np.uint64(1) << (np.uint64(2) * (np.uint64(30) - lv))
        targets = (tl & (~lsb + np.uint64(1))) | lsb
        ix.intersecting_cells(targets[:100000])
        t0 = time.perf_counter(); off, cells, labs = ix.intersecting_cells(targets); dt = time.perf_counter() - t0
        ct = {"workload": "VisitIntersectingCells for %d single-cell targets (levels 8..16) through host buffers, CSR of (cell, label) pairs out" % m,
              "targets_per_s": m / dt, "pairs_found": int(off[-1]), "seconds": dt}
        if not args.no_cpu_baseline:
            try:
                from oracle import reflib  # cpu_baseline leg
                ref = reflib.load()
                k = 1_000_000
                ri = reflib.RefCellIndex(ref, ids, labels)
                t0 = time.perf_counter(); w_off, w_cells, w_labels = ri.visit(targets[:k]); dtc = time.perf_counter() - t0
                ct["cpu_baseline"] = {"value": k / dtc, "unit": "targets/s", "cores": 1, "kind": "reference", "sample": "first %d targets" % k,
                                      "parity_on_sample": bool((off[:k + 1].astype(np.uint64) == w_off).all()
                                                               and np.array_equal(np.sort(labs[:int(off[k])]), np.sort(w_labels)))}
            except Exception as e:
                ct["cpu_baseline"] = {"unavailable": str(e)}
        out["cell_targets"] = ct
        ix.close()
    return out


def region_cells_section(args, torch, dev, rank, world):
    """S2ShapeIndexRegion cell tests, the calls an S2RegionCoverer makes on an index region. Targets:
    cells of levels 6..20 containing points drawn in the 10k-vertex loop's bounding cap (so many of them meet the
    boundary), classified with MayIntersect(S2Cell) and Contains(S2Cell) against the reference-built default-granularity
    index (the answers depend on the index's cells, so the reference's own index is the one to compare on)."""
    import ctypes
    from s2geometry_b200 import _lib, cellid, synth
    from s2geometry_b200.index import FlatIndex, ShapeIndex, ShapeIndexRegion
    fx = np.load(os.path.join(ROOT, "tests", "golden", "index_loop10k.npz"))
    flat = FlatIndex(**{k: fx[k] for k in FlatIndex.FIELDS}, num_shape_ids=int(fx["num_shape_ids"]))
    region = ShapeIndexRegion(ShapeIndex(flat))
    m = min(args.points, 20_000_000)
    center = np.array([0.3, -0.5, 0.8]); radius = 0.17 * 1.25
    pts = synth.stream_cap_points_cuda(78, rank * m, m, center, radius, dev)
    leaf = cellid.from_points(pts, 30)
    level = 6 + ((leaf >> 3) & 0xFFFF) % 15          # levels 6..20, a deterministic function of the point
    lsb = torch.ones_like(level) << (2 * (30 - level))
    targets = ((leaf & -lsb) | lsb).contiguous()
    del pts, leaf, level, lsb
    out = {"workload": "S2ShapeIndexRegion cell tests: %d target cells (levels 6..20, in the bounding cap of the 10k-vertex loop) per GPU vs the "
                       "reference-built index (%d cells, %d clipped edges)" % (m, flat.num_cells, flat.num_edges)}
    for name, fn in (("may_intersect", region.may_intersect), ("contains_cell", region.contains_cells)):
        for _ in range(3):
            res = fn(targets)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            res = fn(targets)
        e1.record(); torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
        out[name] = {"targets_per_s": world * m / (ms * 1e-3), "ms_per_step": ms, "true_fraction": float(res.float().mean().item())}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import reflib  # cpu_baseline leg
            from s2geometry_b200.shapes import ShapeSoup
            ref = reflib.load()
            soup = ShapeSoup(); soup.add_polygon([fx["loop_vertices"]])
            ri = ref.index_create(soup.freeze())
            k = min(m, 1_000_000)
            sample = targets[:k].cpu().numpy().view(np.uint64)
            base = {"cores": 1, "kind": "reference", "sample": "first %d targets, one S2ShapeIndexRegion" % k, "unit": "targets/s"}
            ok = True
            for mode, name, fn in ((0, "may_intersect", region.may_intersect), (1, "contains_cell", region.contains_cells)):
                t0 = time.perf_counter(); want = ri.region_cells(sample, mode); dt = time.perf_counter() - t0
                base[name] = k / dt
                ok = ok and bool((fn(targets[:k]).cpu().numpy() == want).all())
            base["parity_on_sample"] = ok
            out["cpu_baseline"] = base
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)}
    return out


def join_section(args, torch, dev, rank, world):
    """BASELINE configs[3]: spatial join, `join_points` uniform points (default 1B, sharded over the ranks: strong
    scaling) vs 10k synthetic polygons -> per-polygon hit counts; the only collective is the NCCL all-reduce of the
    uint64[10000] histogram. Index built by the library's own host builder and replicated on every GPU."""
    from s2geometry_b200 import dist as sdist, synth
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    from s2geometry_b200.shapes import ShapeSoup
    soup = ShapeSoup()
    for loops in synth.many_polygons(10000, 4):
        soup.add_polygon(loops)
    soup = soup.freeze()
    t0 = time.perf_counter()
    flat = build_flat_index(soup, DEVICE_MAX_EDGES_PER_CELL)   # see pip_section: finer cells than the CPU default, same answers
    build_s = time.perf_counter() - t0
    ix = ShapeIndex(flat)
    query = ContainsPointQuery(ix, 1)
    n_total = args.join_points
    lo, hi = sdist.shard_range(n_total, rank, world)
    n = hi - lo
    # This rank's slice [lo, hi) of ONE global stream of n_total points (counter-based: point i = f(seed, i)), so the
    # histogram is the same array at 1, 2, 4 and 8 GPUs and its sample can be checked against the reference at every N.
    JOIN_STREAM = 1000
    pts = synth.stream_sphere_points_cuda(JOIN_STREAM, lo, n, dev)
    counts = torch.zeros(flat.num_shape_ids, dtype=torch.int64, device=dev)

    def step():
        counts.zero_()
        query.shape_histogram(pts, counts=counts)
        sdist.all_reduce_sum_(counts)

    for _ in range(3):
        step()
    sdist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record(); sdist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    sdist.all_reduce_max_(t)
    ms = float(t.item())
    hbm_peak, _ = peaks()
    ach = 24 * n_total / world / (ms * 1e-3) / 1e9   # per-GPU algorithmic bytes over the step time
    hist = counts.cpu().numpy().astype(np.uint64)
    checksum = int((hist * (np.arange(len(hist), dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))).sum() & np.uint64(0xFFFFFFFFFFFFFFFF))
    out = {"workload": "configs[3]: %d uniform points (stream %d, sharded over %d GPU(s)) vs 10k polygons (%d vertices; index with max_edges_per_cell=1: %d cells, %d clipped edges, %.1f MB on device), per-polygon counts + all-reduce"
                       % (n_total, JOIN_STREAM, world, len(soup.vertices), flat.num_cells, flat.num_edges, ix.info()["device_bytes"] / 1e6),
           "scaling": "strong", "points_per_s": n_total / (ms * 1e-3), "ms_per_step": ms, "index_build_s": build_s,
           "total_hits": int(hist.sum()), "histogram_checksum": "%016x" % checksum, "bytes_per_point": 24,
           "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "note": "per GPU"}}
    # Parity against the reference at EVERY N: the first `k` points of the global stream live on rank 0 for any N; the GPU
    # histogram of exactly those points must equal the reference's VisitContainingShapeIds histogram of the same points.
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import reflib  # cpu_baseline leg
            ref = reflib.load()
            t0 = time.perf_counter(); ri = ref.index_create(soup); ref_build_s = time.perf_counter() - t0
            threads = ref.hardware_threads()
            k = min(n, 20_000_000)
            sample = pts[:k].cpu().numpy()
            assert np.array_equal(sample[:100_000], synth.stream_sphere_points(JOIN_STREAM, 0, 100_000)), "device stream != host stream"
            best = 1e30
            for _ in range(2 if world == 1 else 1):
                t0 = time.perf_counter(); want = ri.shape_histogram(sample, 1, threads); best = min(best, time.perf_counter() - t0)
            got = query.shape_histogram(pts[:k]).cpu().numpy().view(np.uint64)
            out["parity_vs_reference"] = bool((got == want).all())
            out["cpu_baseline"] = {"value": len(sample) / best, "unit": "points/s", "cores": threads, "kind": "reference",
                                   "sample": "first %d points of the global stream, VisitContainingShapeIds histogram" % k, "index_build_s": ref_build_s,
                                   "parity_on_sample": bool((got == want).all())}
        except Exception as e:
            out["cpu_baseline"] = {"unavailable": str(e)}
    sdist.barrier()
    del pts
    # ---- configs[4]: S2CellUnion containment of the same number of points against an S2RegionCoverer covering
    # (max_cells = 1000) of the 10k-vertex loop, computed by the reference's own coverer (tests/golden/covering_loop10k.npz,
    # tools/gen_golden.py covering); global hit count = one-element all-reduce inside the step.
    from s2geometry_b200.index import cell_union_contains
    union_ids = np.load(os.path.join(ROOT, "tests", "golden", "covering_loop10k.npz"))["cell_ids"].astype(np.uint64)
    d_ids = torch.from_numpy(union_ids.view(np.int64)).to(dev)
    UNION_STREAM = 2000
    pts = synth.stream_sphere_points_cuda(UNION_STREAM, lo, n, dev)
    res = None
    hits_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    from s2geometry_b200.index import CellUnionIndex, count_flags
    uq = ContainsPointQuery(CellUnionIndex(union_ids), 1)   # the union replicated on this GPU as an index of interior cells

    def ustep():
        nonlocal res
        res = uq.contains(pts)
        hits_dev.zero_()
        count_flags(res, out=hits_dev)                          # device-side count (s2b_count_flags_dev), no host synchronisation in the step
        sdist.all_reduce_sum_(hits_dev)

    for _ in range(3):
        ustep()
    sdist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        ustep()
    e1.record(); sdist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    sdist.all_reduce_max_(t)
    ums = float(t.item())
    hits = int(hits_dev.item())
    same_as_direct_kernel = bool((cell_union_contains(d_ids, pts) == res).all())   # the plain binary-search kernel agrees
    uach = 25 * n_total / world / (ums * 1e-3) / 1e9
    cu = {"workload": "configs[4]: S2CellUnion = S2RegionCoverer(max_cells=1000) covering of the 10k-vertex loop (%d cells, reference-built fixture), containment of %d uniform points (stream %d) over %d GPU(s), global hit count"
                      % (len(union_ids), n_total, UNION_STREAM, world),
          "scaling": "strong", "points_per_s": n_total / (ums * 1e-3), "ms_per_step": ums, "hits": hits, "bytes_per_point": 25,
          "path": "s2b_index_create_from_cellunion + s2b_contains_dev", "matches_cellunion_kernel": same_as_direct_kernel,
          "roofline": {"bound": "hbm", "achieved": uach, "peak": hbm_peak, "unit": "GB/s", "frac": uach / hbm_peak, "note": "per GPU"}}
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import reflib  # cpu_baseline leg
            ref = reflib.load()
            k = min(n, 20_000_000)
            sample = pts[:k].cpu().numpy()
            t0 = time.perf_counter(); want = ref.cellunion_contains(union_ids, sample, ref.hardware_threads()); dt = time.perf_counter() - t0
            cu["parity_vs_reference"] = bool((res[:k].cpu().numpy() == want).all())
            cu["cpu_baseline"] = {"value": len(sample) / dt, "unit": "points/s", "cores": ref.hardware_threads(), "kind": "reference",
                                  "sample": "first %d points of the global stream" % k, "parity_on_sample": cu["parity_vs_reference"]}
        except Exception as e:
            cu["cpu_baseline"] = {"unavailable": str(e)}
    sdist.barrier()
    return out, cu


def capi_multi_section(args, torch, rank, world, expected_checksum):
    """north_star (4) behind the C ABI: ONE process (rank 0) drives all `world` GPUs of the box through
    s2b_index_replicate + s2b_shape_histogram_multi_dev — shard k of the SAME global stream as join_section resident on
    GPU k, the library's own streams, and its all-reduce of the uint64[10000] histogram inside the call (NCCL bound at
    run time, then the library's peer-memory reduction kernel). The other ranks have released their GPU memory and wait
    on the rendezvous store (a host-side wait: no kernel of theirs is resident meanwhile). Timed with the host clock
    around `steps` synchronous calls (the streams are the library's own), so launch latency is included."""
    import ctypes
    import torch.distributed as dist
    from s2geometry_b200 import _lib, dist as sdist, synth
    from s2geometry_b200.index import ShapeIndex, build_flat_index
    from s2geometry_b200.shapes import ShapeSoup
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    dist.barrier(); torch.cuda.synchronize()
    store = dist.distributed_c10d._get_default_store()
    out = None
    if rank != 0:
        store.wait(["s2b_capi_multi_done"])
        return None
    try:
        lib = _lib.load()
        soup = ShapeSoup()
        for loops in synth.many_polygons(10000, 4):
            soup.add_polygon(loops)
        ix = ShapeIndex(build_flat_index(soup.freeze(), DEVICE_MAX_EDGES_PER_CELL))
        S = 10000
        devs = (ctypes.c_int * world)(*range(world))
        t0 = time.perf_counter()
        _lib.check(lib.s2b_index_replicate(ix._h, devs, world))
        replicate_s = time.perf_counter() - t0
        order = (ctypes.c_int * world)()
        _lib.check(lib.s2b_index_replica_devices(ix._h, order, world))
        assert lib.s2b_index_num_replicas(ix._h) == world
        n_total = args.join_points
        pts, outs, ns = [], [], []
        for k in range(world):
            lo, hi = sdist.shard_range(n_total, k, world)
            d = torch.device("cuda", order[k])
            pts.append(synth.stream_sphere_points_cuda(1000, lo, hi - lo, d))
            outs.append(torch.zeros(S, dtype=torch.int64, device=d))
            ns.append(hi - lo)
            torch.cuda.synchronize(d)
        xyz_p = (ctypes.c_void_p * world)(*[p_.data_ptr() for p_ in pts])
        out_p = (ctypes.c_void_p * world)(*[o_.data_ptr() for o_ in outs])
        n_p = (ctypes.c_size_t * world)(*ns)
        out = {"workload": "configs[3] through the C ABI from one process: %d points (stream 1000) over %d GPUs, s2b_index_replicate + s2b_shape_histogram_multi_dev" % (n_total, world),
               "index_replicate_s": replicate_s, "timing": "host clock around %d synchronous calls" % args.steps}
        for mode, name in ((0, "nccl"), (1, "peer_kernel")):
            prev = lib.s2b_index_set_reduce_mode(ix._h, mode)
            if prev < 0:
                out[name] = {"unavailable": lib.s2b_last_error().decode("utf-8", "replace")}
                continue

            def step():
                _lib.check(lib.s2b_shape_histogram_multi_dev(ix._h, xyz_p, n_p, 1, out_p, 1))
            for _ in range(3):
                step()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            ms = (time.perf_counter() - t0) / args.steps * 1e3
            hists = [o_.cpu().numpy().astype(np.uint64) for o_ in outs]
            hist = hists[0]
            checksum = "%016x" % int((hist * (np.arange(len(hist), dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))).sum() & np.uint64(0xFFFFFFFFFFFFFFFF))
            out[name] = {"points_per_s": n_total / (ms * 1e-3), "ms_per_step": ms, "total_hits": int(hist.sum()), "histogram_checksum": checksum,
                         "same_on_every_gpu": all(bool((h == hist).all()) for h in hists),
                         "equals_torch_distributed_driver": checksum == expected_checksum}
        del pts, outs
        ix.close()
    except Exception as e:
        out = {"error": repr(e)}
    finally:
        store.set("s2b_capi_multi_done", "1")
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from s2geometry_b200 import _lib, cellid, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from s2geometry_b200 import dist as sdist_
    placement = sdist_.bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "why": "single process"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    n = args.points
    # this rank's slice of the global lat/lng stream (weak scaling: rank r encodes points [r*n, (r+1)*n))
    ll = synth.stream_sphere_latlng_cuda(STREAM_LATLNG, rank * n, n, dev)
    ids = torch.empty((len(LEVELS), n), dtype=torch.int64, device=dev)
    tok = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    ln = torch.empty(n, dtype=torch.uint8, device=dev)

    # One step = the strict call (S2B_LATLNG_HOST_LIBM, the default): ONE kernel launch, then the ~1.6e-5 of points whose
    # leaf cell depends on the libm are re-evaluated with the host's sin/cos inside the call (s2b_encode_latlng_strict_dev).
    # The step calls the C ABI entry point itself with arguments marshalled once (cellid.from_lat_lng_levels_tokens wraps
    # the same call; its per-call Python costs ~30 us, which a synchronous 1.1 ms call cannot hide).
    import ctypes as _ct
    _step_lv = (_ct.c_int * len(LEVELS))(*LEVELS)
    _step_args = (1, _ct.c_void_p(ll.data_ptr()), n, _step_lv, len(LEVELS), _ct.c_void_p(ids.data_ptr()), None, TOKEN_LEVEL_INDEX,
                  _ct.c_void_p(tok.data_ptr()), _ct.c_void_p(ln.data_ptr()), _ct.c_void_p(torch.cuda.current_stream().cuda_stream))
    _strict_call = lib.s2b_encode_latlng_strict_dev

    def step():
        rc = _strict_call(*_step_args)
        if rc != 0:
            _lib.check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.launch_count()
    cellid.lat_lng_strict_stats(reset=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    evs[0].record()
    for k in range(args.steps):
        step()
    evs[1].record()
    barrier()
    launches = _lib.launch_count() - launches0
    strict_reevaluated, strict_changed = cellid.lat_lng_strict_stats()
    total_ms = evs[0].elapsed_time(evs[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * n / (ms_per_step * 1e-3)

    # The dominant kernel alone (roofline.achieved): the same launch through the enqueue-only entry point, CUDA events
    # on the launching stream around each launch.
    import ctypes
    lv = (ctypes.c_int * len(LEVELS))(*LEVELS)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def kernel_only():
        _lib.check(lib.s2b_encode_latlng_levels_tokens_dev(ctypes.c_void_p(ll.data_ptr()), n, lv, len(LEVELS), ctypes.c_void_p(ids.data_ptr()),
                                                           TOKEN_LEVEL_INDEX, ctypes.c_void_p(tok.data_ptr()), ctypes.c_void_p(ln.data_ptr()), st))
    for _ in range(3):
        kernel_only()
    torch.cuda.synchronize()
    kevs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kevs[0].record()
    for k in range(args.steps):
        kernel_only()
        kevs[k + 1].record()
    torch.cuda.synchronize()
    per_launch_ms = [kevs[k].elapsed_time(kevs[k + 1]) for k in range(args.steps)]
    step()   # leave the strict results in ids / tok / ln for the parity check below
    torch.cuda.synchronize()

    # ---- end to end through the host-pointer C ABI call with pinned host buffers
    h_ll = torch.empty((n, 2), dtype=torch.float64, pin_memory=True)
    h_ids = torch.empty((len(LEVELS), n), dtype=torch.int64, pin_memory=True)
    h_tok = torch.empty((n, 16), dtype=torch.uint8, pin_memory=True)
    h_len = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_ll.copy_(ll)
    a_ll, a_ids, a_tok, a_len = h_ll.numpy(), h_ids.numpy().view(np.uint64), h_tok.numpy(), h_len.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    cellid.from_lat_lng_levels_tokens(a_ll, LEVELS, TOKEN_LEVEL_INDEX, out_ids=a_ids, out_tokens=a_tok, out_len=a_len)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cellid.from_lat_lng_levels_tokens(a_ll, LEVELS, TOKEN_LEVEL_INDEX, out_ids=a_ids, out_tokens=a_tok, out_len=a_len)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * n / e2e_s
    # the e2e result equals the device-resident result (every id of every level)
    same = bool((h_ids.to(dev) == ids).all())

    # Host-link ceiling for the same bytes: every rank at once copies the step's input H2D and the step's outputs D2H
    # with plain cudaMemcpyAsync (one call per array, two streams, nothing else): what the host side of this box can move.
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copies_only():
        with torch.cuda.stream(s_in):
            ll.copy_(h_ll, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_ids.copy_(ids, non_blocking=True); h_tok.copy_(tok, non_blocking=True); h_len.copy_(ln, non_blocking=True)
    copies_only(); barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copies_only()
    barrier()
    link_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([link_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    link_s = float(t.item())
    e2e_bytes = (16 + 8 * len(LEVELS) + 17) * n
    host_link = {"points_per_s": world * n / link_s, "ms_per_step": 1e3 * link_s, "gb_per_s_all_ranks": world * e2e_bytes / link_s / 1e9,
                 "e2e_fraction_of_ceiling": link_s / e2e_s,
                 "how": "same bytes per rank, plain cudaMemcpyAsync H2D (input) and D2H (outputs) on two streams, all ranks at once"}
    del h_ll, h_ids, h_tok, h_len, a_ll, a_ids, a_tok, a_len

    headline_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            headline_cpu = cpu_baseline(ll, ids, tok, ln)
        except Exception as e:  # the oracle library is test infrastructure; its absence must not hide the GPU number
            headline_cpu = {"value": None, "unit": "points/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    del ids, tok, ln, ll
    torch.cuda.empty_cache()
    pip = None
    enc_xyz = None
    if not args.no_pip:
        try:
            enc_xyz = encode_xyz_section(args, torch, dev, rank, world)
            torch.cuda.empty_cache()
            pip = pip_section(args, torch, dev, rank, world)
        except Exception as e:
            pip = {"error": repr(e)}
    join = cellunion = None
    if not args.no_join:
        try:
            torch.cuda.empty_cache()
            join, cellunion = join_section(args, torch, dev, rank, world)
        except Exception as e:
            join = {"error": repr(e)}
    radius_join = None
    if not args.no_join:
        try:
            torch.cuda.empty_cache()
            radius_join = radius_join_section(args, torch, dev, rank, world)
            torch.cuda.empty_cache()
        except Exception as e:
            radius_join = {"error": repr(e)}
    cell_index_join = None
    if not args.no_join:
        try:
            cell_index_join = cell_index_join_section(args, torch, dev, rank, world)
            torch.cuda.empty_cache()
        except Exception as e:
            cell_index_join = {"error": repr(e)}
    region_cells = None
    if not args.no_join:
        try:
            region_cells = region_cells_section(args, torch, dev, rank, world)
            torch.cuda.empty_cache()
        except Exception as e:
            region_cells = {"error": repr(e)}
    capi_multi = None
    if world > 1 and not args.no_join:
        capi_multi = capi_multi_section(args, torch, rank, world, (join or {}).get("histogram_checksum"))
    if rank == 0:
        hbm_peak, peak_src = peaks()
        kern_ms = float(np.mean(per_launch_ms))
        achieved = BYTES_PER_POINT * n / (kern_ms * 1e-3) / 1e9
        dfma = _lib.ctypes.c_double(0)
        fp64 = None
        if lib.s2b_diag_fp64_peak(_lib.ctypes.byref(dfma)) == 0 and dfma.value > 0:
            fp64_ach = FP64_INSTR_PER_POINT * n / (kern_ms * 1e-3)
            fp64 = {"peak_dfma_per_s": dfma.value, "peak_tflops": 2 * dfma.value / 1e12, "fp64_instr_per_point": FP64_INSTR_PER_POINT,
                    "fp64_instr_source": PROFILE_SOURCE, "achieved_instr_per_s": fp64_ach, "frac": fp64_ach / dfma.value,
                    "peak_source": "measured live (s2b_diag_fp64_peak)"}

        def brief(sec, *path):
            try:
                for k in path:
                    sec = sec[k]
                return sec
            except Exception:
                return None

        def gpts(v):
            return None if v is None else round(v / 1e9, 2)

        def frac(v):
            return None if v is None else round(v, 3)
        # compact secondary results in a key the driver keeps (the full sections follow at the end of the line)
        summary = {
            "encode_latlng_Gpts": gpts(value), "encode_latlng_hbm_frac": frac(achieved / hbm_peak), "encode_latlng_fp64_frac": frac(brief(fp64, "frac")),
            "encode_latlng_parity": brief(headline_cpu, "parity_on_sample"),
            "strict_reevaluated_per_step": strict_reevaluated / max(1, args.steps), "strict_changed_per_step": strict_changed / max(1, args.steps),
            "encode_xyz_Gpts": gpts(brief(enc_xyz, "points_per_s")), "encode_xyz_hbm_frac": frac(brief(enc_xyz, "roofline", "frac")),
            "pip_uniform_Gpts": gpts(brief(pip, "uniform", "points_per_s")), "pip_uniform_hbm_frac": frac(brief(pip, "uniform", "roofline", "frac")),
            "pip_near_Gpts": gpts(brief(pip, "near", "points_per_s")), "pip_near_hbm_frac": frac(brief(pip, "near", "roofline", "frac")),
            "pip_parity": None if brief(pip, "cpu_baseline", "uniform") is None else bool(brief(pip, "cpu_baseline", "uniform", "parity_on_sample") and brief(pip, "cpu_baseline", "near", "parity_on_sample")),
            "join_Gpts": gpts(brief(join, "points_per_s")), "join_hbm_frac": frac(brief(join, "roofline", "frac")),
            "join_total_hits": brief(join, "total_hits"), "join_histogram_checksum": brief(join, "histogram_checksum"),
            "join_parity": brief(join, "parity_vs_reference"),
            "cellunion_Gpts": gpts(brief(cellunion, "points_per_s")), "cellunion_hbm_frac": frac(brief(cellunion, "roofline", "frac")),
            "cellunion_hits": brief(cellunion, "hits"), "cellunion_parity": brief(cellunion, "parity_vs_reference"),
            "cell_index_join_Gpts": gpts(brief(cell_index_join, "points_per_s")), "cell_index_join_hbm_frac": frac(brief(cell_index_join, "roofline", "frac")),
        }
        if capi_multi is not None:
            summary["capi_multi_join_nccl_Gpts"] = gpts(brief(capi_multi, "nccl", "points_per_s"))
            summary["capi_multi_join_peer_Gpts"] = gpts(brief(capi_multi, "peer_kernel", "points_per_s"))
            summary["capi_multi_join_equals_join"] = brief(capi_multi, "nccl", "equals_torch_distributed_driver")
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(n),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "step": "s2b_encode_latlng_strict_dev: encode_kernel<latlng> (1 launch: fast sin/cos filter + correctly rounded path for boundary-sensitive points, fused parent/ToToken epilogues) + host-libm re-evaluation of the libm-sensitive points (S2B_LATLNG_HOST_LIBM)",
                         # compact results of the secondary sections (configs[0], [2], [3], [4] and the widened rows): Gpts/s, fraction
                         # of the HBM roofline at their algorithmic bytes per point, parity against the reference
                         "secondary": summary,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from ncu --set full, scaled to this launch
                         "traffic": DRAM_BYTES_PER_POINT * n, "traffic_source": "profile: " + PROFILE_SOURCE,
                         "traffic_unit": "bytes per launch (ncu dram read+write, %.2f B/point)" % DRAM_BYTES_PER_POINT,
                         "peak_source": peak_src, "bytes_per_point": BYTES_PER_POINT, "kernel_ms": kern_ms,
                         "kernel": "encode_kernel<latlng,3 levels,token>, timed alone (enqueue-only entry point, CUDA events per launch)",
                         "fp64": fp64},
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": (8 * len(LEVELS) + 17) * n, "points_per_gpu": n,
                    "ms_per_step": 1e3 * e2e_s, "matches_device_result": same, "host_link_ceiling": host_link, "host_placement": placement},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if headline_cpu is not None:
            line["cpu_baseline"] = headline_cpu
        for name, sec in (("encode_xyz", enc_xyz), ("pip", pip), ("join", join), ("cellunion", cellunion), ("radius_join", radius_join),
                          ("cell_index_join", cell_index_join), ("region_cells", region_cells), ("capi_multi_join", capi_multi)):
            if sec is not None:
                line[name] = sec
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else that writes to fd 1 (NCCL's version
    banner at N > 1, library chatter) was pointed at stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS, help="points per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pip", action="store_true", help="skip the secondary point-in-polygon (configs[2]) measurements")
    ap.add_argument("--no-join", action="store_true", help="skip the spatial-join (configs[3]) and cell-union (configs[4]) measurements")
    ap.add_argument("--join-points", type=int, default=1_000_000_000, help="TOTAL points of the join / cell-union sections (sharded over ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
