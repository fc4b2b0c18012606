#!/bin/bash
# Throughput of s2b_index_point_distances (batched S2ClosestEdgeQuery::GetDistance): 20M uniform points vs the 10k-polygon
# index and vs the 10k-vertex loop, 10 km and 100 km limits; reference on 16 threads for a 500k sample.
mkdir -p gpurun_out
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/distances.log
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from s2geometry_b200 import synth
from s2geometry_b200.index import ShapeIndex, build_flat_index
from s2geometry_b200.shapes import ShapeSoup
from oracle import reflib
from tests import scenes
ref = reflib.load()
soup = ShapeSoup()
for loops in synth.many_polygons(10000, 4):
    soup.add_polygon(loops)
soup = soup.freeze()
loop_soup, _ = scenes.loop_scene(10000)
n = 20_000_000
pts = synth.stream_sphere_points_cuda(5000, 0, n, "cuda")
near = synth.stream_cap_points_cuda(5001, 0, n, np.array(scenes.LOOP_CENTER, float), scenes.LOOP_RADIUS * 1.25, "cuda")
for name, sp, batch in (("10k polygons, uniform points", soup, pts), ("10k-vertex loop, points in its cap", loop_soup, near)):
    ix = ShapeIndex(build_flat_index(sp, 0))
    ri = ref.index_create(sp)
    for km in (10.0, 100.0):
        r = km / 6371.0
        ix.point_distances(batch, r); torch.cuda.synchronize()
        t0 = time.perf_counter(); d = ix.point_distances(batch, r); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        sample = batch[:500_000].cpu().numpy()
        t0 = time.perf_counter(); want = ri.point_distances(sample, r, True, ref.hardware_threads()); rt = time.perf_counter() - t0
        same = bool((d[:500_000].cpu().numpy().view(np.uint64) == want.view(np.uint64)).all())
        print("%s, limit %g km: %.1f M points/s (%.0f ms per %dM incl. the temporary point index), within limit %.3f; reference %.2f M points/s on %d threads; identical on the sample: %s"
              % (name, km, n / dt / 1e6, dt * 1e3, n // 1_000_000, float(torch.isfinite(d).double().mean()), len(sample) / rt / 1e6, ref.hardware_threads(), same))
    ix.close()
PY
