#!/usr/bin/env python3
"""Times the index queries of the bench (join histogram vs 10k polygons, Contains vs the 10k-vertex loop: uniform and
in-cap points) for the library named by S2B_LIB_PATH (default: in-tree). Index arrays are cached in /tmp between runs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from s2geometry_b200 import synth
from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, build_flat_index
from s2geometry_b200.shapes import ShapeSoup

def cached(name, make):
    path = "/tmp/%s.npz" % name
    if os.path.exists(path):
        return FlatIndex.load(path)
    f = make(); f.save(path); return f

def join_flat():
    soup = ShapeSoup()
    for loops in synth.many_polygons(10000, 4):
        soup.add_polygon(loops)
    return build_flat_index(soup.freeze(), 1)

def loop_flat():
    fx = np.load(os.path.join(ROOT, "tests", "golden", "index_loop10k.npz"))
    s = ShapeSoup(); s.add_polygon([fx["loop_vertices"]])
    return build_flat_index(s.freeze(), 1)

def timed(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best

n = int(os.environ.get("JOIN_POINTS", 200_000_000))
name = os.environ.get("S2B_LIB_PATH", "in-tree").split("/")[-1]
jq = ContainsPointQuery(ShapeIndex(cached("join_flat", join_flat)), 1)
pts = synth.stream_sphere_points_cuda(1000, 0, n, "cuda")
counts = torch.zeros(10000, dtype=torch.int64, device="cuda")
def jstep():
    counts.zero_(); jq.shape_histogram(pts, counts=counts)
ms = timed(jstep, 5)
out = "%-20s join %.1f Gpts/s (hits %d)" % (name, n / ms / 1e6, int(counts.sum().item()))
lq = ContainsPointQuery(ShapeIndex(cached("loop_flat", loop_flat)), 1)
m = min(n, 100_000_000)
ms = timed(lambda: lq.contains(pts[:m]))
out += "  pip-uniform %.1f" % (m / ms / 1e6)
near = synth.stream_cap_points_cuda(7, 0, m, np.array([0.3, -0.5, 0.8]), 0.17 * 1.25, "cuda")
res = None
def nstep():
    global res
    res = lq.contains(near)
ms = timed(nstep)
out += "  pip-near %.1f (inside %d)" % (m / ms / 1e6, int(res.sum().item()))
print(out, flush=True)
