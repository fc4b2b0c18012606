#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, '.')
from s2geometry_b200 import cellid, synth
from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index, cell_union_contains
from tests import scenes
ll = synth.sphere_latlng(300001, 1); p = synth.sphere_points(300001, 2)
cellid.from_lat_lng_levels_tokens(ll, [12, 20, 30], 2); cellid.from_points(p); cellid.from_lat_lng(torch.from_numpy(ll).cuda(), 30, True)
cellid.to_token(cellid.from_points(p)); cellid.parent(cellid.from_points(p), 5)
soup, v = scenes.loop_scene(2000)
ix = ShapeIndex(build_flat_index(soup))
u, b, d = scenes.loop_points(v, 200001, 100003, 3)
for m in (0, 1, 2):
    q = ContainsPointQuery(ix, m)
    for pts in (u, b, d):
        q.contains(pts); q.shape_histogram(pts); q.containing_shape_ids(pts); q.shape_contains(0, pts)
        t = torch.from_numpy(pts).cuda(); q.contains(t); q.shape_histogram(t); q.containing_shape_ids(t)
ix2 = ShapeIndex(build_flat_index(scenes.axis_aligned_scene()))
for m in (0, 1, 2):
    ContainsPointQuery(ix2, m).shape_histogram(scenes.axis_aligned_queries())
cell_union_contains(build_flat_index(soup, 32).cell_ids, p)
torch.cuda.synchronize(); print("sanitize workload ok")
PY
compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"
tail -5 gpurun_out/memcheck.log
