#!/bin/bash
# Region cell tests on the GPU: memcheck of a small workload, then a full ncu capture of the MayIntersect kernel.
mkdir -p gpurun_out
cat > /tmp/san_region.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from oracle import reflib
from s2geometry_b200.index import FlatIndex, ShapeIndex, ShapeIndexRegion
from tests import region_scenes
ref = reflib.load()
for name, soup in region_scenes.region_scene_list():
    flat = FlatIndex(**ref.index_create(soup).flatten())
    region = ShapeIndexRegion(ShapeIndex(flat))
    t = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 5, 1000)
    region.may_intersect(t); region.contains_cells(t); region.intersecting_shapes(t)
    region.may_intersect(torch.from_numpy(t.view(np.int64)).cuda())
torch.cuda.synchronize(); print("sanitize region workload ok")
PY
compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_region.py > gpurun_out/memcheck_region.log 2>&1; echo "memcheck exit $?"
tail -4 gpurun_out/memcheck_region.log
python tools/region_bench.py --no-cpu-baseline > gpurun_out/plain_region.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:region_cells_kernel -s 3 -c 1 -o gpurun_out/prof_region python tools/region_bench.py --no-cpu-baseline > gpurun_out/ncu_region.log 2>&1
ls -la gpurun_out | tail -4
