#!/usr/bin/env python3
"""Times the query kernels (CUDA events) on U / near / join workloads. Usage: python tools/time_query.py [n]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from s2geometry_b200 import synth
from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, build_flat_index
from s2geometry_b200.shapes import ShapeSoup

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
fx = np.load(os.path.join(ROOT, "tests", "golden", "index_loop10k.npz"))
flat = FlatIndex(**{k: fx[k] for k in FlatIndex.FIELDS}, num_shape_ids=int(fx["num_shape_ids"]))
if os.environ.get('S2B_MEPC'):
    ls = ShapeSoup(); ls.add_polygon([fx["loop_vertices"]])
    flat = build_flat_index(ls.freeze(), int(os.environ['S2B_MEPC']))
q = ContainsPointQuery(ShapeIndex(flat), 1)
u = synth.sphere_points_cuda(n, 3, dev)
b = torch.from_numpy(synth.cap_points(np.array([0.3, -0.5, 0.8]), 0.17 * 1.25, n // 4, 7)).to(dev)
soup = ShapeSoup()
for loops in synth.many_polygons(10000, 4):
    soup.add_polygon(loops)
mepc = int(os.environ.get('S2B_MEPC', '0'))
flat_j = build_flat_index(soup.freeze(), mepc)
jq = ContainsPointQuery(ShapeIndex(flat_j), 1)

def timeit(f, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {"mepc": mepc, "join_cells": int(flat_j.num_cells), "join_edges": int(flat_j.num_edges), "loop_cells": int(flat.num_cells), "loop_edges": int(flat.num_edges)}
out["uniform_gpts"] = n / timeit(lambda: q.contains(u)) / 1e6
out["near_gpts"] = (n // 4) / timeit(lambda: q.contains(b)) / 1e6
cnt = torch.zeros(10000, dtype=torch.int64, device=dev)
out["join_gpts"] = n / timeit(lambda: jq.shape_histogram(u, counts=cnt)) / 1e6
out["join_any_gpts"] = n / timeit(lambda: jq.contains(u)) / 1e6
print(json.dumps(out))
