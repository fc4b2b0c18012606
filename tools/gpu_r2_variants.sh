#!/bin/bash
# Kernel-only timing of the headline encode step for each experimental build in variants/ (and the in-tree library).
mkdir -p gpurun_out
shopt -s nullglob
for lib in "" variants/*.so; do
  S2B_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python - <<'PY'
import os, ctypes, torch
from s2geometry_b200 import _lib, synth
lib = _lib.load()
n = 100_000_000
ll = synth.stream_sphere_latlng_cuda(2, 0, n, "cuda")
ids = torch.empty((3, n), dtype=torch.int64, device="cuda"); tok = torch.empty((n, 16), dtype=torch.uint8, device="cuda"); ln = torch.empty(n, dtype=torch.uint8, device="cuda")
lv = (ctypes.c_int * 3)(12, 20, 30)
def k():
    _lib.check(lib.s2b_encode_latlng_levels_tokens_dev(ctypes.c_void_p(ll.data_ptr()), n, lv, 3, ctypes.c_void_p(ids.data_ptr()), 2, ctypes.c_void_p(tok.data_ptr()), ctypes.c_void_p(ln.data_ptr()), None))
for _ in range(5): k()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(20): k()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print("%-28s kernel %.4f ms  %.1f Gpts/s  frac %.3f  checksum %x" % (os.environ.get("S2B_LIB_PATH", "in-tree").split("/")[-1], best, n / best / 1e6, 57 * n / (best * 1e-3) / 1e9 / 6541.1, int(ids[2].sum().item()) & 0xFFFFFFFFFFFF))
PY
done 2>&1 | tee gpurun_out/variants.log
