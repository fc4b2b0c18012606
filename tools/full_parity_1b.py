#!/usr/bin/env python3
"""North-star check: bit-exact S2ShapeIndex point-in-polygon on 1B points (BASELINE configs[3]).
1B uniform points vs 10k polygons: the per-polygon hit histogram from the GPU (own index builder, 64M-point rounds)
is compared with the reference's S2ContainsPointQuery::VisitContainingShapeIds over the SAME points, all host threads,
100M points at a time. Also compares Contains() flags on every point. Writes gpurun_out/full_parity_1b.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import reflib
from s2geometry_b200 import synth
from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index, exact_stage_count
from s2geometry_b200.shapes import ShapeSoup

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
chunk = 100_000_000
dev = torch.device("cuda", 0)
ref = reflib.load()
soup = ShapeSoup()
for loops in synth.many_polygons(10000, 4):
    soup.add_polygon(loops)
soup = soup.freeze()
q = ContainsPointQuery(ShapeIndex(build_flat_index(soup, 1)), 1)   # device granularity (S2B_MAX_EDGES_PER_CELL_DEVICE); the reference index below keeps its default
ri = ref.index_create(soup)
threads = ref.hardware_threads()
gpu_counts = torch.zeros(10000, dtype=torch.int64, device=dev)
ref_counts = np.zeros(10000, np.uint64)
flag_mismatch = 0; gpu_s = 0.0; cpu_s = 0.0; done = 0
exact_stage_count(reset=True)
k = 0
while done < total:
    n = min(chunk, total - done)
    pts = synth.sphere_points_cuda(n, 5000 + k, dev)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    q.shape_histogram(pts, counts=gpu_counts)
    flags = q.contains(pts)
    e1.record(); torch.cuda.synchronize(); gpu_s += e0.elapsed_time(e1) * 1e-3
    host = pts.cpu().numpy()
    t0 = time.perf_counter()
    ref_counts += ri.shape_histogram(host, 1, threads)
    want = ri.contains(host, 1, threads)
    cpu_s += time.perf_counter() - t0
    flag_mismatch += int((flags.cpu().numpy() != want).sum())
    done += n; k += 1
    print("chunk", k, "done", done, flush=True)
got = gpu_counts.cpu().numpy().view(np.uint64)
out = {"points": total, "polygons": 10000, "histogram_equal": bool((got == ref_counts).all()), "histogram_mismatching_polygons": int((got != ref_counts).sum()),
       "total_hits": int(ref_counts.sum()), "contains_flag_mismatches": flag_mismatch, "gpu_seconds_hist_plus_contains": gpu_s,
       "reference_seconds_hist_plus_contains": cpu_s, "reference_threads": threads, "exact_stage_evaluations_on_gpu": exact_stage_count()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "full_parity_1b.json"), "w"), indent=1)
print(json.dumps(out))
