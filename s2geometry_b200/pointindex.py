"""S2PointIndex + S2ClosestPointQuery, batch form (src/s2/s2point_index.h, s2closest_point_query.h), over the C ABI.

    index = PointIndex(points)                       # S2PointIndex<int>: Add(points[i], i)
    q = ClosestPointQuery(index, max_distance_rad=r) # options.max_distance = S1Angle::Radians(r) (exclusive)
    q.count(targets)                                 # len(FindClosestPoints(PointTarget(t))) per target  (radius join)
    q.closest(targets)                               # FindClosestPoint: (data, length2) per target, data = -1 if none

NumPy arrays take the host path; CUDA tensors are used in place on the index's device.
"""
import ctypes
import math

import numpy as np

from . import _lib

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def chord_length2(angle_rad):
    """S1ChordAngle(S1Angle::Radians(a)).length2() (s1chord_angle.h:339-350); inf stays inf (no limit)."""
    if math.isinf(angle_rad) and angle_rad > 0:
        return math.inf
    if angle_rad < 0:
        return -1.0
    length = 2 * math.sin(0.5 * min(math.pi, angle_rad))
    return length * length


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class PointIndex:
    def __init__(self, points):
        self._lib = _lib.load()
        pts = np.ascontiguousarray(points.detach().cpu().numpy() if _is_tensor(points) else points, np.float64).reshape(-1, 3)
        h = ctypes.c_void_p()
        _lib.check(self._lib.s2b_point_index_create(pts.ctypes.data if pts.size else None, len(pts), ctypes.byref(h)))
        self._h = h
        self.num_points = len(pts)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.s2b_point_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


TARGET_POINT, TARGET_EDGE, TARGET_CELL = 0, 1, 2   # include/s2b_capi.h S2B_TARGET_*


class ClosestPointQuery:
    """Batch form of S2ClosestPointQuery<int>: options.max_distance (exclusive) and, for the general form, max_results and
    max_error; targets are points (count / closest / closest_k) or, through `find`, edges and cells as well."""

    def __init__(self, index, max_distance_rad=math.inf, max_error_rad=0.0):
        self.index = index
        self._lib = index._lib
        self.max_length2 = chord_length2(max_distance_rad)
        self.max_error_rad = float(max_error_rad)

    def find(self, target_kind, targets, max_results):
        """FindClosestPoints for PointTarget (TARGET_POINT, (N,3) float64), EdgeTarget (TARGET_EDGE, (N,6): a, b) or
        CellTarget (TARGET_CELL, N uint64 / int64 cell ids) with options.max_results = max_results (1..32) ->
        (data (N,k) padded with -1, length2 (N,k) padded with inf, counts (N,)); max_results = 0 (unlimited) -> counts only."""
        k = int(max_results)
        kind = int(target_kind)
        if _is_tensor(targets):
            t = targets.contiguous()
            n = t.numel() if kind == TARGET_CELL else t.shape[0]
            cnt = torch.empty(n, dtype=torch.int32, device=t.device)
            data = torch.empty((n, k), dtype=torch.int32, device=t.device) if k else None
            l2 = torch.empty((n, k), dtype=torch.float64, device=t.device) if k else None
            _lib.check(self._lib.s2b_closest_points_to_targets_dev(self.index._h, kind, t.data_ptr(), n, k, self.max_length2, self.max_error_rad,
                                                                   data.data_ptr() if k else None, l2.data_ptr() if k else None, cnt.data_ptr(), _stream()))
            return (data, l2, cnt) if k else cnt
        t = np.ascontiguousarray(targets, np.uint64 if kind == TARGET_CELL else np.float64)
        n = t.size if kind == TARGET_CELL else len(t.reshape(-1, 3 if kind == TARGET_POINT else 6))
        cnt = np.zeros(n, np.uint32)
        data = np.full((n, max(k, 1)), -1, np.int32)
        l2 = np.full((n, max(k, 1)), np.inf, np.float64)
        _lib.check(self._lib.s2b_closest_points_to_targets(self.index._h, kind, t.ctypes.data if t.size else None, n, k, self.max_length2,
                                                           self.max_error_rad, data.ctypes.data, l2.ctypes.data, cnt.ctypes.data))
        return (data, l2, cnt) if k else cnt

    def find_shape_index(self, shape_index, max_results=0, include_interiors=True):
        """FindClosestPoints(ShapeIndexTarget(shape_index)): the index points within max_distance of the geometry of a
        device-resident ShapeIndex (s2geometry_b200.index.ShapeIndex on the same GPU), nearest first -> (data, length2);
        points inside a polygon of the index are at distance 0 when include_interiors. max_results = 0: no limit."""
        import ctypes
        k = int(max_results)
        total = ctypes.c_size_t(0)
        cap = k if k else 1 << 16
        while True:
            data = np.empty(cap, np.int32); l2 = np.empty(cap, np.float64)
            rc = self._lib.s2b_closest_points_to_shape_index(self.index._h, shape_index._h, 1 if include_interiors else 0, k, self.max_length2,
                                                             self.max_error_rad, data.ctypes.data, l2.ctypes.data, cap, ctypes.byref(total))
            if rc == _lib.S2B_ECAPACITY:
                cap = int(total.value)
                continue
            _lib.check(rc)
            return data[: total.value].copy(), l2[: total.value].copy()

    def count(self, targets):
        if _is_tensor(targets):
            t = targets.contiguous()
            out = torch.empty(t.shape[0], dtype=torch.int32, device=t.device)
            _lib.check(self._lib.s2b_closest_points_count_dev(self.index._h, t.data_ptr(), t.shape[0], self.max_length2, out.data_ptr(), _stream()))
            return out
        t = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
        out = np.zeros(len(t), np.uint32)
        _lib.check(self._lib.s2b_closest_points_count(self.index._h, t.ctypes.data if t.size else None, len(t), self.max_length2, out.ctypes.data))
        return out

    def closest(self, targets):
        if _is_tensor(targets):
            t = targets.contiguous()
            data = torch.empty(t.shape[0], dtype=torch.int32, device=t.device)
            l2 = torch.empty(t.shape[0], dtype=torch.float64, device=t.device)
            _lib.check(self._lib.s2b_closest_point_dev(self.index._h, t.data_ptr(), t.shape[0], self.max_length2, data.data_ptr(), l2.data_ptr(), _stream()))
            return data, l2
        t = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
        data = np.full(len(t), -1, np.int32)
        l2 = np.full(len(t), np.inf, np.float64)
        _lib.check(self._lib.s2b_closest_point(self.index._h, t.ctypes.data if t.size else None, len(t), self.max_length2, data.ctypes.data, l2.ctypes.data))
        return data, l2

    def closest_k(self, targets, k):
        """FindClosestPoints with max_results = k: (data (N,k) padded with -1, length2 (N,k) padded with inf, counts (N,))."""
        k = int(k)
        if _is_tensor(targets):
            t = targets.contiguous()
            data = torch.empty((t.shape[0], k), dtype=torch.int32, device=t.device)
            l2 = torch.empty((t.shape[0], k), dtype=torch.float64, device=t.device)
            cnt = torch.empty(t.shape[0], dtype=torch.int32, device=t.device)
            _lib.check(self._lib.s2b_closest_points_dev(self.index._h, t.data_ptr(), t.shape[0], k, self.max_length2, data.data_ptr(), l2.data_ptr(),
                                                        cnt.data_ptr(), _stream()))
            return data, l2, cnt
        t = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
        data = np.full((len(t), k), -1, np.int32)
        l2 = np.full((len(t), k), np.inf, np.float64)
        cnt = np.zeros(len(t), np.uint32)
        _lib.check(self._lib.s2b_closest_points(self.index._h, t.ctypes.data if t.size else None, len(t), k, self.max_length2, data.ctypes.data,
                                                l2.ctypes.data, cnt.ctypes.data))
        return data, l2, cnt
