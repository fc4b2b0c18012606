#!/bin/bash
mkdir -p gpurun_out
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
u_pts, b, d = scenes.loop_points(v, 200001, 100003, 3)
for m in (0, 1, 2):
    q = ContainsPointQuery(ix, m)
    for pts in (u_pts, b, d):
        q.contains(pts); q.shape_histogram(pts); q.containing_shape_ids(pts); q.shape_contains(0, pts)
        t = torch.from_numpy(pts).cuda(); q.contains(t); q.shape_histogram(t); q.containing_shape_ids(t)
ix2 = ShapeIndex(build_flat_index(scenes.axis_aligned_scene()))
for m in (0, 1, 2):
    ContainsPointQuery(ix2, m).shape_histogram(scenes.axis_aligned_queries())
cell_union_contains(build_flat_index(soup, 32).cell_ids, p)
# widened rows: cell algebra, point index, fine-granularity index with the cell sort forced on
from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
ids = cellid.from_points(p[:50001], 30); lv = cellid.parent(ids, 11)
cellid.vertex_neighbors(ids, 7); cellid.all_neighbors(lv, 11); cellid.all_neighbors(lv, 13, 20); cellid.to_lat_lng(ids); cellid.edge_neighbors(lv)
mixed = np.concatenate([lv, cellid.parent(ids, 9), ids[:1000]])
u = cellid.cell_union_normalize(mixed); cellid.cell_union_normalize(torch.from_numpy(mixed.view(np.int64)).cuda())
cellid.cell_union_contains_cells(u, ids); cellid.cell_union_contains_cells(u[:1], lv)
pi = PointIndex(p[:100003])
for r in (1e-4, 0.01, 0.2, 2.5):
    ClosestPointQuery(pi, r).count(p[200000:205001]); ClosestPointQuery(pi, r).closest(p[200000:205001])
ClosestPointQuery(pi).closest(torch.from_numpy(p[:3001]).cuda()); ClosestPointQuery(PointIndex(p[:3]), 0.5).closest(p[:1001])
import os
os.environ["S2B_SORT_BY_CELL"] = "1"
ixf = ShapeIndex(build_flat_index(soup, 1))
for m in (0, 1, 2):
    q = ContainsPointQuery(ixf, m)
    for pts in (u_pts, b, d):
        q.contains(pts); q.shape_histogram(pts); q.containing_shape_ids(pts)
torch.cuda.synchronize(); print("sanitize workload ok")
PY
compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"
tail -5 gpurun_out/memcheck.log
