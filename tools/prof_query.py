#!/usr/bin/env python3
"""Small, deterministic launch sequence for ncu: uniform / near point sets against the 10k-vertex loop fixture and the
10k-polygon join. One warm-up + one profiled call each. Usage: python tools/prof_query.py [n_points]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from s2geometry_b200 import synth
from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, build_flat_index
from s2geometry_b200.shapes import ShapeSoup

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda", 0)
fx = np.load(os.path.join(ROOT, "tests", "golden", "index_loop10k.npz"))
flat = FlatIndex(**{k: fx[k] for k in FlatIndex.FIELDS}, num_shape_ids=int(fx["num_shape_ids"]))
ls = ShapeSoup(); ls.add_polygon([fx["loop_vertices"]])
flat = build_flat_index(ls.freeze(), 1)
q = ContainsPointQuery(ShapeIndex(flat), 1)
u = synth.sphere_points_cuda(n, 3, dev)
b = torch.from_numpy(synth.cap_points(np.array([0.3, -0.5, 0.8]), 0.17 * 1.25, n // 4, 7)).to(dev)
soup = ShapeSoup()
for loops in synth.many_polygons(10000, 4):
    soup.add_polygon(loops)
jq = ContainsPointQuery(ShapeIndex(build_flat_index(soup.freeze(), 1)), 1)   # device granularity (S2B_MAX_EDGES_PER_CELL_DEVICE)
for rep in range(2):      # launches: rep0 = warm-up (6 kernels), rep1 = profiled (6 kernels): U(A,B) near(A,B) join(A,B)
    q.contains(u); q.contains(b); jq.shape_histogram(u)
torch.cuda.synchronize()
print("ok")
