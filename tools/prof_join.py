#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from s2geometry_b200 import synth
from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
from s2geometry_b200.shapes import ShapeSoup
n = 20_000_000
dev = torch.device("cuda", 0)
soup = ShapeSoup()
for loops in synth.many_polygons(10000, 4):
    soup.add_polygon(loops)
jq = ContainsPointQuery(ShapeIndex(build_flat_index(soup.freeze())), 1)
u = synth.sphere_points_cuda(n, 3, dev)
cnt = torch.zeros(10000, dtype=torch.int64, device=dev)
for rep in range(2):
    jq.shape_histogram(u, counts=cnt)   # 3 kernels: locate, process(edges), process(interior)
torch.cuda.synchronize(); print("ok")
