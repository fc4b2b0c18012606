#!/bin/bash
# Every entry point once on small inputs (round-1 memcheck workload + the round-2 additions).
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
# round 2: strict lat/lng mode on device arrays, counter streams, edge / cell / shape-index targets, region queries over
# edges that need RobustCrossProd's extended stages, flag counting, the multi-GPU entry points on one device
from s2geometry_b200.index import ShapeIndexRegion, count_flags
from s2geometry_b200.pointindex import TARGET_CELL, TARGET_EDGE
from s2geometry_b200.shapes import ShapeSoup
dll = synth.stream_sphere_latlng_cuda(2, 0, 400003, "cuda")
cellid.from_lat_lng_levels_tokens(dll, [12, 20, 30], 2)
bl = np.radians(np.array([[0.0, 45.0], [35.26438968275466, 45.0], [90.0, 0.0], [-90.0, 180.0], [0.0, -135.0]]))
cellid.from_lat_lng_levels_tokens(torch.from_numpy(np.tile(bl, (2000, 1))).cuda(), [0, 30], 1)
synth.stream_sphere_points_cuda(3, 5, 100001, "cuda"); synth.stream_cap_points_cuda(7, 0, 100001, np.array([0.3, -0.5, 0.8]), 0.2, "cuda")
count_flags(ContainsPointQuery(ix, 1).contains(torch.from_numpy(b).cuda()))
e = np.ascontiguousarray(np.concatenate([p[:2000], p[2000:4000]], axis=1)); e[::7, 3:] = e[::7, :3]
cq = ClosestPointQuery(pi, 0.05)
cq.find(TARGET_EDGE, e, 0); cq.find(TARGET_EDGE, e, 5); cq.find(TARGET_CELL, lv[:3000], 3); cq.find(TARGET_CELL, cellid.parent(ids[:500], 3), 0)
cq.find_shape_index(ix, 0, True); ClosestPointQuery(pi).find_shape_index(ix, 7, False); ClosestPointQuery(pi, 0.01).find_shape_index(ixf, 50, True)
tiny = ShapeSoup()
a0 = np.array([1.0, 0.0, 0.0]); tiny.add_polyline(np.array([a0, [1.0, 1e-16, 0.0], [1.0, 1e-16, 3e-17]]) / 1.0)
tiny.add_polyline(np.array([[0.0, 1.0, 0.0], [1e-300, -1.0, 0.0]]))
tiny.add_polygon([v[:50]])
rix = ShapeIndex(build_flat_index(tiny.freeze()))
reg = ShapeIndexRegion(rix)
cells = np.concatenate([cellid.parent(cellid.from_points(np.array([a0, [0.0, 1.0, 0.0], [0.0, -1.0, 0.0]])), l) for l in (0, 5, 12, 20, 30)])
reg.may_intersect(cells); reg.contains_cells(cells); reg.intersecting_shapes(cells)
torch.cuda.synchronize(); print("sanitize workload ok")
PY
# compute-sanitizer is closed on this pool (round 2): the workload runs plain, with CUDA_LAUNCH_BLOCKING so that a faulting
# kernel is reported at its own call
CUDA_LAUNCH_BLOCKING=1 timeout 900 python /tmp/san.py > gpurun_out/workload.log 2>&1; echo "workload exit $?"
tail -5 gpurun_out/workload.log
