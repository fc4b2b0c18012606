#!/usr/bin/env python3
"""Would overlapping the join's latency-bound tail kernels with the next round's pre-filter pay? Probe: the same 1B points
as ONE call on one stream vs TWO half-size calls enqueued from two host threads on two streams (the hardware then
co-schedules whatever fits). Prints both throughputs."""
import os, sys, threading, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from s2geometry_b200 import _lib, synth
from s2geometry_b200.index import FlatIndex, ShapeIndex, build_flat_index
from s2geometry_b200.shapes import ShapeSoup

def join_flat():
    path = "/tmp/join_flat.npz"
    if os.path.exists(path):
        return FlatIndex.load(path)
    soup = ShapeSoup()
    for loops in synth.many_polygons(10000, 4):
        soup.add_polygon(loops)
    f = build_flat_index(soup.freeze(), 1); f.save(path); return f

lib = _lib.load()
ix = ShapeIndex(join_flat())
n = int(os.environ.get("JOIN_POINTS", 1_000_000_000))
pts = synth.stream_sphere_points_cuda(1000, 0, n, "cuda")
half = n // 2
counts = [torch.zeros(10000, dtype=torch.int64, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def call(k, lo, cnt):
    _lib.check(lib.s2b_shape_histogram_dev(ix._h, ctypes.c_void_p(pts.data_ptr() + 24 * lo), cnt, 1, ctypes.c_void_p(counts[k].data_ptr()),
                                           ctypes.c_void_p(streams[k].cuda_stream)))

def one():
    counts[0].zero_(); torch.cuda.synchronize()
    call(0, 0, n)

def two():
    counts[0].zero_(); counts[1].zero_(); torch.cuda.synchronize()
    ts = [threading.Thread(target=call, args=(k, k * half, half)) for k in range(2)]
    for t in ts: t.start()
    for t in ts: t.join()

def timed(f, reps=5):
    for _ in range(2): f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        import time
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = min(best, dt * 1e3)
    return best

a = timed(one); h1 = int(counts[0].sum().item())
b = timed(two); h2 = int((counts[0] + counts[1]).sum().item())
print("one call: %.1f Gpts/s (%.2f ms), two concurrent half calls: %.1f Gpts/s (%.2f ms); hits %d / %d" % (n / a / 1e6, a, n / b / 1e6, b, h1, h2))
