#!/bin/bash
# S2CellIndex point join: parity tests, then the bench section's two scenes timed.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cellindex_gpu.py -q -m gpu --maxfail=5 2>&1 | tail -8 | tee gpurun_out/pytest_ci.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/ci_join.log
import numpy as np, torch, ctypes
from s2geometry_b200 import _lib, cellid, synth
from s2geometry_b200.cellindex import CellIndex
lib = _lib.load()
num_labels = 10000
centers = synth.sphere_points(num_labels, 55)
leaves = cellid.from_points(centers, 30)
n = 100_000_000
pts = synth.stream_sphere_points_cuda(300, 0, n, "cuda")
def scene(lo, hi):
    rng = np.random.default_rng(123)
    label_level = rng.integers(lo, hi + 1, num_labels)
    ids, labels = [], []
    for level in range(lo, hi + 1):
        sel = np.nonzero(label_level == level)[0]
        c = cellid.parent(leaves[sel], level)
        nb, cnt = cellid.all_neighbors(c, level)
        for k, lab in enumerate(sel):
            cells = np.concatenate([c[k:k + 1], nb[k, :cnt[k]]])
            ids.append(cells); labels.append(np.full(len(cells), lab, np.int32))
    return np.concatenate(ids).astype(np.uint64), np.concatenate(labels).astype(np.int32)
for lo, hi in ((5, 11), (9, 13)):
    ids, labels = scene(lo, hi)
    ix = CellIndex(ids, labels)
    for _ in range(3): c = ix.point_histogram(pts)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c = ix.point_histogram(pts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("levels %d..%d: %.1f Gpts/s  frac %.3f  pairs %d" % (lo, hi, n / ms / 1e6, 24 * n / ms / 1e6 / 6541.1, int(c.sum().item())))
PY
