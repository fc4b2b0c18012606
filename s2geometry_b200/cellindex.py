"""S2CellIndex joins, batch form (src/s2/s2cell_index.h) and the S2RegionSharder on top (s2region_sharder.h), over the C ABI.

    index = CellIndex(cell_ids, labels)            # S2CellIndex: Add(cell_ids[i], labels[i]) ...; Build()
    off, cells, labels = index.intersecting_cells(targets)   # VisitIntersectingCells per target (CSR)
    off, labels = index.intersecting_labels(targets)         # GetIntersectingLabels per target (distinct, ascending)
    counts = index.point_histogram(points)                   # per-label hits of the leaf cells of `points`
    shard = RegionSharder(shards).most_intersecting_shard(bounds, regions, default_shard)

`targets` / `regions` are S2CellUnions: a 1-D uint64 array means one cell per target, a list of arrays one union per target
(cells sorted and non-overlapping, as S2CellUnion requires). NumPy arrays take the host path; for `point_histogram` a CUDA
tensor of points is used in place on the index's device (counts come back as a CUDA tensor).
"""
import ctypes

import numpy as np

from . import _lib

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _csr(targets):
    if isinstance(targets, np.ndarray) and targets.ndim == 1:
        return np.ascontiguousarray(targets, np.uint64), None, len(targets)
    lens = [len(t) for t in targets]
    ids = np.concatenate([np.asarray(t, np.uint64) for t in targets]) if lens else np.zeros(0, np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    return np.ascontiguousarray(ids, np.uint64), off, len(lens)


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class CellIndex:
    def __init__(self, cell_ids, labels):
        self._lib = _lib.load()
        ids = np.ascontiguousarray(cell_ids, np.uint64)
        lab = np.ascontiguousarray(labels, np.int32)
        if ids.shape != lab.shape or ids.ndim != 1:
            raise ValueError("cell_ids and labels must be 1-D arrays of the same length")
        h = ctypes.c_void_p()
        _lib.check(self._lib.s2b_cell_index_create(_ptr(ids), _ptr(lab), len(ids), ctypes.byref(h)))
        self._h = h
        n, nl, nb = ctypes.c_uint64(), ctypes.c_int32(), ctypes.c_uint64()
        _lib.check(self._lib.s2b_cell_index_info(h, ctypes.byref(n), ctypes.byref(nl), ctypes.byref(nb)))
        self.num_cells, self.num_labels, self.device_bytes = n.value, nl.value, nb.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.s2b_cell_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def intersecting_cells(self, targets):
        """VisitIntersectingCells (s2cell_index.h:619-647) per target: (offsets[T+1], cell_ids, labels); order within a
        target is unspecified, as in the reference."""
        if torch is not None and isinstance(targets, torch.Tensor):
            return self._intersecting_cells_dev(targets)
        ids, off, T = _csr(targets)
        out_off = np.zeros(T + 1, np.uint32)
        total = ctypes.c_size_t(0)
        rc = self._lib.s2b_cell_index_intersecting_cells(self._h, _ptr(ids), _ptr(off) if off is not None else None, T, out_off.ctypes.data,
                                                         None, None, 0, ctypes.byref(total))
        if rc not in (_lib.S2B_OK, _lib.S2B_ECAPACITY):
            _lib.check(rc)
        cells = np.empty(total.value, np.uint64)
        labels = np.empty(total.value, np.int32)
        if total.value:
            _lib.check(self._lib.s2b_cell_index_intersecting_cells(self._h, _ptr(ids), _ptr(off) if off is not None else None, T, out_off.ctypes.data,
                                                                   cells.ctypes.data, labels.ctypes.data, total.value, ctypes.byref(total)))
        return out_off, cells, labels

    def _intersecting_cells_dev(self, targets):
        """Device path: `targets` = int64 CUDA tensor of single-cell targets; returns CUDA tensors (offsets int32, cells int64, labels int32)."""
        t = targets.contiguous()
        T = t.numel()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        off = torch.zeros(T + 1, dtype=torch.int32, device=t.device)
        total = ctypes.c_size_t(0)
        cap = max(1024, 4 * T)
        while True:
            cells = torch.empty(cap, dtype=torch.int64, device=t.device)
            labels = torch.empty(cap, dtype=torch.int32, device=t.device)
            rc = self._lib.s2b_cell_index_intersecting_cells_dev(self._h, t.data_ptr(), None, T, off.data_ptr(), cells.data_ptr(), labels.data_ptr(), cap,
                                                                 ctypes.byref(total), st)
            if rc == _lib.S2B_ECAPACITY and total.value > cap:
                cap = total.value
                continue
            _lib.check(rc)
            return off, cells[: total.value], labels[: total.value]

    def intersecting_labels(self, targets):
        """GetIntersectingLabels (s2cell_index.cc:148-163) per target: (offsets[T+1], labels ascending, distinct)."""
        ids, off, T = _csr(targets)
        out_off = np.zeros(T + 1, np.uint32)
        total = ctypes.c_size_t(0)
        cap = max(1024, 2 * T)
        while True:
            labels = np.empty(cap, np.int32)
            rc = self._lib.s2b_cell_index_intersecting_labels(self._h, _ptr(ids), _ptr(off) if off is not None else None, T, out_off.ctypes.data,
                                                              labels.ctypes.data, cap, ctypes.byref(total))
            if rc == _lib.S2B_ECAPACITY and total.value > cap:
                cap = total.value
                continue
            _lib.check(rc)
            return out_off, labels[: total.value].copy()

    def most_intersecting_label(self, bounds, regions=None, default_label=-1):
        """GetMostIntersectingShard per region: `bounds[t]` = the normalized region.GetCellUnionBound() covering, `regions[t]` =
        the region as an S2CellUnion (None: the bound is the region)."""
        ids, off, T = _csr(bounds)
        r_ids = r_off = None
        if regions is not None:
            r_ids, r_off, rT = _csr(list(regions))
            if rT != T:
                raise ValueError("bounds and regions must have the same length")
        out = np.full(T, default_label, np.int32)
        _lib.check(self._lib.s2b_cell_index_most_intersecting_label(
            self._h, _ptr(ids), _ptr(off) if off is not None else None, r_ids.ctypes.data if r_ids is not None else None,
            r_off.ctypes.data if r_off is not None else None, T, int(default_label), out.ctypes.data if T else None))
        return out

    def point_histogram(self, points):
        """counts[label] = number of (point, indexed cell) pairs with that label whose cell contains the point's leaf cell."""
        if torch is not None and isinstance(points, torch.Tensor):
            p = points.contiguous()
            counts = torch.zeros(max(self.num_labels, 1), dtype=torch.int64, device=p.device)
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(self._lib.s2b_cell_index_point_histogram_dev(self._h, p.data_ptr(), p.shape[0], counts.data_ptr(), st))
            return counts[: self.num_labels]
        p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        counts = np.zeros(max(self.num_labels, 1), np.uint64)
        _lib.check(self._lib.s2b_cell_index_point_histogram(self._h, _ptr(p), len(p), counts.ctypes.data))
        return counts[: self.num_labels]


class RegionSharder:
    """S2RegionSharder (s2region_sharder.h:36-79) for regions given as S2CellUnions: shard s = shards[s] (an S2CellUnion)."""

    def __init__(self, shards):
        ids = [np.asarray(s, np.uint64) for s in shards]
        labels = [np.full(len(s), i, np.int32) for i, s in enumerate(ids)]
        self.index = CellIndex(np.concatenate(ids) if ids else np.zeros(0, np.uint64), np.concatenate(labels) if labels else np.zeros(0, np.int32))

    def most_intersecting_shard(self, bounds, regions=None, default_shard=-1):
        """bounds[t] = S2CellUnion(region_t.GetCellUnionBound()); regions[t] = region_t as an S2CellUnion (MayIntersect filter)."""
        return self.index.most_intersecting_label(bounds, regions, default_shard)

    def intersecting_shards(self, regions):
        """GetIntersectingShards: (offsets, shard numbers ascending)."""
        return self.index.intersecting_labels(regions)
