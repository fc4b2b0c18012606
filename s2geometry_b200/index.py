"""Batched point-in-shape queries against a device-resident shape index.

Host-side mirror of the reference interface for this path (src/s2/):
  S2ShapeIndex / MutableS2ShapeIndex   s2shape_index.h, mutable_s2shape_index.h   -> ShapeIndex (flattened, on the GPU)
  S2ContainsPointQuery<Index>          s2contains_point_query.h:92-186            -> ContainsPointQuery
  S2VertexModel                        s2contains_point_query.h:54                -> VertexModel
  S2CellUnion::Contains(S2Point)       s2cell_union.cc:564-566                    -> cell_union_contains

The flat form (FlatIndex) is what crosses the C ABI (S2bFlatIndex in include/s2b_capi.h): it is obtained either by
walking a reference-built index through its public iterator or from this package's own builder.
"""
import ctypes
import enum
import math

import numpy as np

from . import _lib

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class VertexModel(enum.IntEnum):
    OPEN = 0
    SEMI_OPEN = 1
    CLOSED = 2


class S2bFlatIndexStruct(ctypes.Structure):
    _fields_ = [
        ("num_cells", ctypes.c_uint64), ("num_clips", ctypes.c_uint64), ("num_edges", ctypes.c_uint64),
        ("num_shape_ids", ctypes.c_int32),
        ("cell_ids", ctypes.c_void_p), ("cell_center", ctypes.c_void_p), ("cell_clip_begin", ctypes.c_void_p),
        ("clip_shape_id", ctypes.c_void_p), ("clip_flags", ctypes.c_void_p), ("clip_edge_begin", ctypes.c_void_p),
        ("edge_v0", ctypes.c_void_p), ("edge_v1", ctypes.c_void_p),
    ]


class FlatIndex:
    """Host arrays of a flattened shape index (see S2bFlatIndex)."""

    FIELDS = ("cell_ids", "cell_center", "cell_clip_begin", "clip_shape_id", "clip_flags", "clip_edge_begin", "edge_v0", "edge_v1")
    DTYPES = (np.uint64, np.float64, np.uint32, np.int32, np.uint8, np.uint32, np.float64, np.float64)

    def __init__(self, num_shape_ids, cell_ids, cell_clip_begin, clip_shape_id, clip_flags, clip_edge_begin, edge_v0, edge_v1,
                 cell_center=None, **_ignored):
        self.num_shape_ids = int(num_shape_ids)
        self.cell_ids = np.ascontiguousarray(cell_ids, np.uint64)
        self.cell_center = None if cell_center is None else np.ascontiguousarray(cell_center, np.float64).reshape(-1, 3)
        self.cell_clip_begin = np.ascontiguousarray(cell_clip_begin, np.uint32)
        self.clip_shape_id = np.ascontiguousarray(clip_shape_id, np.int32)
        self.clip_flags = np.ascontiguousarray(clip_flags, np.uint8)
        self.clip_edge_begin = np.ascontiguousarray(clip_edge_begin, np.uint32)
        self.edge_v0 = np.ascontiguousarray(edge_v0, np.float64).reshape(-1, 3)
        self.edge_v1 = np.ascontiguousarray(edge_v1, np.float64).reshape(-1, 3)
        if len(self.cell_clip_begin) != len(self.cell_ids) + 1 or len(self.clip_edge_begin) != len(self.clip_shape_id) + 1:
            raise ValueError("offset arrays must have one more entry than the arrays they index")

    @property
    def num_cells(self):
        return len(self.cell_ids)

    @property
    def num_clips(self):
        return len(self.clip_shape_id)

    @property
    def num_edges(self):
        return len(self.edge_v0)

    def as_struct(self):
        def p(a):
            return None if a is None or a.size == 0 else a.ctypes.data
        return S2bFlatIndexStruct(self.num_cells, self.num_clips, self.num_edges, self.num_shape_ids,
                                  p(self.cell_ids), p(self.cell_center), self.cell_clip_begin.ctypes.data,
                                  p(self.clip_shape_id), p(self.clip_flags), self.clip_edge_begin.ctypes.data,
                                  p(self.edge_v0), p(self.edge_v1))

    def save(self, path):
        d = {k: getattr(self, k) for k in self.FIELDS if getattr(self, k) is not None}
        np.savez_compressed(path, num_shape_ids=self.num_shape_ids, **d)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        return cls(**{k: z[k] for k in z.files})


def build_flat_index(soup, max_edges_per_cell=0):
    """Builds a FlatIndex from a ShapeSoup with the library's own host builder (s2b_index_build): the counterpart of
    MutableS2ShapeIndex::Add + ForceBuild for callers without a reference-built index."""
    lib = _lib.load()
    soup = soup.freeze()
    h = ctypes.c_void_p()
    keep = (soup.shape_dim, soup.shape_chain_begin, soup.chain_vertex_begin, soup.vertices)
    _lib.check(lib.s2b_index_build(int(soup.num_shapes), keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data,
                                   keep[3].ctypes.data if keep[3].size else None, int(max_edges_per_cell), ctypes.byref(h)))
    return _take_built(lib, h)


def _take_built(lib, h):
    """Copies a S2bBuiltIndex into a FlatIndex (with .edge_id) and destroys the handle."""
    try:
        st = S2bFlatIndexStruct()
        eid = ctypes.c_void_p()
        _lib.check(lib.s2b_built_index_flat(h, ctypes.byref(st), ctypes.byref(eid)))

        def arr(ptr, count, dtype):
            if not ptr or count == 0:
                return np.zeros(count, dtype)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(count,)).copy()
        C, K, E = st.num_cells, st.num_clips, st.num_edges
        flat = FlatIndex(st.num_shape_ids, arr(st.cell_ids, C, np.uint64), arr(st.cell_clip_begin, C + 1, np.uint32),
                         arr(st.clip_shape_id, K, np.int32), arr(st.clip_flags, K, np.uint8), arr(st.clip_edge_begin, K + 1, np.uint32),
                         arr(st.edge_v0, 3 * E, np.float64), arr(st.edge_v1, 3 * E, np.float64), cell_center=arr(st.cell_center, 3 * C, np.float64))
        flat.edge_id = arr(eid.value, E, np.int32)
        return flat
    finally:
        lib.s2b_built_index_destroy(h)


def decode_index(data, soup):
    """Reads the reference's serialized index (the bytes of MutableS2ShapeIndex::Encode) into a FlatIndex; `soup` supplies
    the geometry the edge ids refer to (the counterpart of MutableS2ShapeIndex::Init's ShapeFactory). Returns
    (flat_index, max_edges_per_cell, bytes_consumed)."""
    lib = _lib.load()
    soup = soup.freeze()
    buf = np.frombuffer(bytes(data), np.uint8)
    h = ctypes.c_void_p()
    max_edges = ctypes.c_int(0)
    used = ctypes.c_size_t(0)
    keep = (soup.shape_dim, soup.shape_chain_begin, soup.chain_vertex_begin, soup.vertices)
    _lib.check(lib.s2b_index_decode(buf.ctypes.data if buf.size else None, int(buf.size), int(soup.num_shapes), keep[0].ctypes.data,
                                    keep[1].ctypes.data, keep[2].ctypes.data, keep[3].ctypes.data if keep[3].size else None,
                                    ctypes.byref(h), ctypes.byref(max_edges), ctypes.byref(used)))
    return _take_built(lib, h), int(max_edges.value), int(used.value)


def encode_index(flat, max_edges_per_cell=0):
    """Serializes a FlatIndex (which must carry .edge_id) in the reference's index format (MutableS2ShapeIndex::Encode)."""
    lib = _lib.load()
    eid = getattr(flat, "edge_id", None)
    if eid is None:
        raise ValueError("encode_index needs flat.edge_id (shape-local edge ids)")
    eid = np.ascontiguousarray(eid, np.int32)
    st = flat.as_struct()
    size = ctypes.c_size_t(0)
    rc = lib.s2b_index_encode(ctypes.byref(st), eid.ctypes.data if eid.size else None, int(max_edges_per_cell), None, 0, ctypes.byref(size))
    if rc != _lib.S2B_ECAPACITY:
        _lib.check(rc)
    out = np.zeros(max(int(size.value), 1), np.uint8)
    _lib.check(lib.s2b_index_encode(ctypes.byref(st), eid.ctypes.data if eid.size else None, int(max_edges_per_cell), out.ctypes.data,
                                    int(out.size), ctypes.byref(size)))
    return out[:int(size.value)].tobytes()


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _points(xyz):
    if _is_tensor(xyz):
        if not xyz.is_cuda or xyz.dtype != torch.float64 or xyz.dim() != 2 or xyz.shape[1] != 3:
            raise ValueError("expected a CUDA float64 (N,3) tensor")
        return xyz.contiguous()
    a = np.ascontiguousarray(xyz, np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("expected an (N,3) float64 array")
    return a


def _ptr(a):
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else ctypes.c_void_p(a.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class ShapeIndex:
    """A flattened MutableS2ShapeIndex resident on the current CUDA device (one replica per process/GPU)."""

    def __init__(self, flat):
        self._lib = _lib.load()
        self.flat = flat
        h = ctypes.c_void_p()
        st = flat.as_struct()
        _lib.check(self._lib.s2b_index_create(ctypes.byref(st), ctypes.byref(h)))
        self._h = h
        eid = getattr(flat, "edge_id", None)
        if eid is not None:      # optional: shape-local edge ids enable incident-edge queries
            eid = np.ascontiguousarray(eid, np.int32)
            _lib.check(self._lib.s2b_index_set_edge_ids(h, eid.ctypes.data if eid.size else None, eid.size))
        self.num_shape_ids = flat.num_shape_ids
        self.device = torch.cuda.current_device() if torch is not None and torch.cuda.is_available() else 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.s2b_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def locate_points(self, xyz):
        """S2ShapeIndex::Iterator::Locate(S2Point) for a batch: position of the containing index cell per point, or -1."""
        p = _points(xyz)
        n = p.shape[0]
        if _is_tensor(p):
            out = torch.empty(n, dtype=torch.int32, device=p.device)
            with torch.cuda.device(p.device):
                _lib.check(self._lib.s2b_index_locate_points_dev(self._h, _ptr(p), n, _ptr(out), _stream()))
            return out
        out = np.empty(n, np.int32)
        _lib.check(self._lib.s2b_index_locate_points(self._h, _ptr(p), n, _ptr(out)))
        return out

    def locate_cells(self, cell_ids):
        """S2ShapeIndex::Iterator::Locate(S2CellId) for a batch: (relation uint8: 0 INDEXED / 1 SUBDIVIDED / 2 DISJOINT,
        cell int32: position of the index cell the iterator lands on, -1 when disjoint)."""
        if _is_tensor(cell_ids):
            ids = cell_ids.contiguous()
            rel = torch.empty(ids.numel(), dtype=torch.uint8, device=ids.device)
            cell = torch.empty(ids.numel(), dtype=torch.int32, device=ids.device)
            with torch.cuda.device(ids.device):
                _lib.check(self._lib.s2b_index_locate_cells_dev(self._h, _ptr(ids), ids.numel(), _ptr(rel), _ptr(cell), _stream()))
            return rel, cell
        ids = np.ascontiguousarray(cell_ids, np.uint64)
        rel = np.empty(ids.size, np.uint8); cell = np.empty(ids.size, np.int32)
        _lib.check(self._lib.s2b_index_locate_cells(self._h, _ptr(ids), ids.size, _ptr(rel), _ptr(cell)))
        return rel, cell

    def point_distances(self, xyz, max_distance_rad=math.inf, include_interiors=True):
        """S2ClosestEdgeQuery(index).GetDistance(PointTarget(p)) per point as S1ChordAngle::length2() (0 inside polygons when
        include_interiors; inf where the distance is not below max_distance_rad)."""
        p = _points(xyz)
        n = p.shape[0]
        limit2 = math.inf if math.isinf(max_distance_rad) else (2.0 * math.sin(0.5 * min(max_distance_rad, math.pi))) ** 2
        if _is_tensor(p):
            out = torch.empty(n, dtype=torch.float64, device=p.device)
            with torch.cuda.device(p.device):
                torch.cuda.current_stream().synchronize()
                _lib.check(self._lib.s2b_index_point_distances_dev(self._h, _ptr(p), n, 1 if include_interiors else 0, limit2, _ptr(out)))
            return out
        out = np.empty(n, np.float64)
        _lib.check(self._lib.s2b_index_point_distances(self._h, _ptr(p), n, 1 if include_interiors else 0, limit2, _ptr(out)))
        return out

    def info(self):
        c = ctypes.c_uint64(); k = ctypes.c_uint64(); e = ctypes.c_uint64(); s = ctypes.c_int32(); b = ctypes.c_uint64()
        _lib.check(self._lib.s2b_index_info(self._h, ctypes.byref(c), ctypes.byref(k), ctypes.byref(e), ctypes.byref(s), ctypes.byref(b)))
        return dict(num_cells=c.value, num_clips=k.value, num_edges=e.value, num_shape_ids=s.value, device_bytes=b.value)


class ShapeIndexRegion:
    """S2ShapeIndexRegion (s2shape_index_region.h) over a device-resident index, for batches of target cells: the
    S2Region interface an S2RegionCoverer drives, and VisitIntersectingShapeIds. numpy arrays go through the host-pointer
    C ABI, CUDA tensors through the `_dev` one."""

    def __init__(self, index):
        self.index = index
        self._lib = index._lib

    def _cells(self, cell_ids, host_fn, dev_fn):
        h = self.index._h
        if _is_tensor(cell_ids):
            ids = cell_ids.contiguous()
            out = torch.empty(ids.numel(), dtype=torch.uint8, device=ids.device)
            with torch.cuda.device(ids.device):
                _lib.check(dev_fn(h, _ptr(ids), ids.numel(), _ptr(out), _stream()))
            return out
        ids = np.ascontiguousarray(cell_ids, np.uint64)
        out = np.empty(ids.size, np.uint8)
        _lib.check(host_fn(h, _ptr(ids), ids.size, _ptr(out)))
        return out

    def may_intersect(self, cell_ids):
        """S2ShapeIndexRegion::MayIntersect(S2Cell(id)) per id (s2shape_index_region.h:344-371)."""
        return self._cells(cell_ids, self._lib.s2b_region_may_intersect_cells, self._lib.s2b_region_may_intersect_cells_dev)

    def contains_cells(self, cell_ids):
        """S2ShapeIndexRegion::Contains(S2Cell(id)) per id (s2shape_index_region.h:313-342)."""
        return self._cells(cell_ids, self._lib.s2b_region_contains_cells, self._lib.s2b_region_contains_cells_dev)

    def contains(self, xyz):
        """S2ShapeIndexRegion::Contains(S2Point) = S2ContainsPointQuery(SEMI_OPEN)::Contains (:437-445)."""
        return ContainsPointQuery(self.index, 1).contains(xyz)

    def cell_union_bound(self):
        """S2ShapeIndexRegion::GetCellUnionBound (:232-305): at most six cell ids covering the index."""
        out = np.zeros(6, np.uint64)
        cnt = ctypes.c_int(0)
        _lib.check(self._lib.s2b_region_cell_union_bound(self.index._h, _ptr(out), ctypes.byref(cnt)))
        return out[: cnt.value].copy()

    def covering_at_level(self, level):
        """S2RegionCoverer::GetCovering(region) with min_level = max_level = level: the sorted level-`level` cells that may
        intersect the index, found by subdividing from cell_union_bound() with batched MayIntersect tests on the device."""
        total = ctypes.c_size_t(0)
        rc = self._lib.s2b_region_covering_at_level(self.index._h, int(level), None, 0, ctypes.byref(total))
        if rc != _lib.S2B_ECAPACITY:
            _lib.check(rc)
            return np.zeros(0, np.uint64)
        out = np.empty(total.value, np.uint64)
        _lib.check(self._lib.s2b_region_covering_at_level(self.index._h, int(level), _ptr(out), out.size, ctypes.byref(total)))
        return out

    def intersecting_shapes(self, cell_ids):
        """VisitIntersectingShapeIds per target cell (:373-424) as (offsets uint32[n+1], shape_ids int32, contains uint8),
        pairs ascending by shape id within a target."""
        ids = np.ascontiguousarray(cell_ids, np.uint64)
        off = np.zeros(ids.size + 1, np.uint32)
        total = ctypes.c_size_t(0)
        rc = self._lib.s2b_region_intersecting_shapes(self.index._h, _ptr(ids), ids.size, _ptr(off), None, None, 0, ctypes.byref(total))
        if rc != _lib.S2B_ECAPACITY:
            _lib.check(rc)
            return off, np.zeros(0, np.int32), np.zeros(0, np.uint8)
        shp = np.empty(total.value, np.int32); con = np.empty(total.value, np.uint8)
        _lib.check(self._lib.s2b_region_intersecting_shapes(self.index._h, _ptr(ids), ids.size, _ptr(off), _ptr(shp), _ptr(con),
                                                           total.value, ctypes.byref(total)))
        return off, shp, con


class CellUnionIndex(ShapeIndex):
    """An S2CellUnion resident on the device as an index of interior cells: ContainsPointQuery(CellUnionIndex(ids)).contains(p)
    is S2CellUnion::Contains(S2Point) per point, on the fast Locate path."""

    def __init__(self, sorted_ids):
        self._lib = _lib.load()
        ids = np.ascontiguousarray(sorted_ids, np.uint64)
        h = ctypes.c_void_p()
        _lib.check(self._lib.s2b_index_create_from_cellunion(ids.ctypes.data if ids.size else None, ids.size, ctypes.byref(h)))
        self._h = h
        self.flat = None
        self.num_shape_ids = 1
        self.device = torch.cuda.current_device() if torch is not None and torch.cuda.is_available() else 0


class _DeviceGuard:
    """Makes the index's device current for the duration of a call (kernels run on the device that holds the replica)."""

    def __init__(self, device):
        self.ctx = torch.cuda.device(device) if torch is not None and torch.cuda.is_available() else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


class ContainsPointQuery:
    """Batch form of S2ContainsPointQuery: every method takes (N,3) points (NumPy = host path, CUDA tensor = device path)."""

    def __init__(self, index, vertex_model=VertexModel.SEMI_OPEN):
        self.index = index
        self.model = int(vertex_model)
        self._lib = _lib.load()

    def _check_device(self, p):
        if _is_tensor(p) and p.device.index != self.index.device:
            raise ValueError("points are on cuda:%s but the index replica lives on cuda:%s" % (p.device.index, self.index.device))
        return _DeviceGuard(self.index.device)

    def contains(self, xyz):
        """S2ContainsPointQuery::Contains per point -> uint8 flags."""
        p = _points(xyz)
        n = p.shape[0]
        with self._check_device(p):
            if _is_tensor(p):
                out = torch.empty(n, dtype=torch.uint8, device=p.device)
                _lib.check(self._lib.s2b_contains_dev(self.index._h, _ptr(p), n, self.model, _ptr(out), _stream()))
                return out
            out = np.empty(n, np.uint8)
            _lib.check(self._lib.s2b_contains(self.index._h, _ptr(p), n, self.model, _ptr(out)))
            return out

    def shape_contains(self, shape_id, xyz):
        """S2ContainsPointQuery::ShapeContains(shape_id, p) per point -> uint8 flags."""
        p = _points(xyz)
        n = p.shape[0]
        with self._check_device(p):
            if _is_tensor(p):
                out = torch.empty(n, dtype=torch.uint8, device=p.device)
                _lib.check(self._lib.s2b_shape_contains_dev(self.index._h, int(shape_id), _ptr(p), n, self.model, _ptr(out), _stream()))
                return out
            out = np.empty(n, np.uint8)
            _lib.check(self._lib.s2b_shape_contains(self.index._h, int(shape_id), _ptr(p), n, self.model, _ptr(out)))
            return out

    def shape_histogram(self, xyz, counts=None):
        """Spatial join: number of points contained by each shape (VisitContainingShapeIds accumulated).
        Device path: `counts` (int64 CUDA tensor of num_shape_ids, accumulated into) may be passed to chain batches."""
        p = _points(xyz)
        n = p.shape[0]
        with self._check_device(p):
            S = self.index.num_shape_ids
            if _is_tensor(p):
                if counts is None:
                    counts = torch.zeros(S, dtype=torch.int64, device=p.device)
                _lib.check(self._lib.s2b_shape_histogram_dev(self.index._h, _ptr(p), n, self.model, _ptr(counts), _stream()))
                return counts
            out = np.zeros(S, np.uint64)
            _lib.check(self._lib.s2b_shape_histogram(self.index._h, _ptr(p), n, self.model, _ptr(out)))
            return out

    def incident_edges(self, xyz):
        """VisitIncidentEdges for a batch in CSR form: (offsets (N+1,), shape_ids, edge_ids) — the edges of each point's
        index cell that have the point as an endpoint. The index must have been created from a FlatIndex with .edge_id."""
        p = _points(xyz)
        n = p.shape[0]
        with self._check_device(p):
            if _is_tensor(p):
                cnt = torch.empty(n, dtype=torch.int32, device=p.device)
                _lib.check(self._lib.s2b_incident_edges_count_dev(self.index._h, _ptr(p), n, _ptr(cnt), _stream()))
                off = torch.zeros(n + 1, dtype=torch.int64, device=p.device)
                torch.cumsum(cnt, 0, out=off[1:])
                total = int(off[-1].item())
                off32 = off.to(torch.int32)
                sid = torch.empty(max(total, 1), dtype=torch.int32, device=p.device)
                eid = torch.empty(max(total, 1), dtype=torch.int32, device=p.device)
                if total:
                    _lib.check(self._lib.s2b_incident_edges_fill_dev(self.index._h, _ptr(p), n, _ptr(off32), _ptr(sid), _ptr(eid), _stream()))
                return off32, sid[:total], eid[:total]
            off = np.empty(n + 1, np.uint32)
            total = ctypes.c_size_t(0)
            cap = max(1024, n // 8)
            while True:
                sid = np.empty(cap, np.int32); eid = np.empty(cap, np.int32)
                rc = self._lib.s2b_incident_edges(self.index._h, _ptr(p), n, _ptr(off), _ptr(sid), _ptr(eid), cap, ctypes.byref(total))
                if rc == -4 and total.value > cap:
                    cap = total.value
                    continue
                _lib.check(rc)
                return off, sid[: total.value].copy(), eid[: total.value].copy()

    def containing_shape_ids(self, xyz):
        """GetContainingShapeIds for a batch in CSR form: (offsets (N+1,), shape_ids)."""
        p = _points(xyz)
        n = p.shape[0]
        with self._check_device(p):
            if _is_tensor(p):
                cnt = torch.empty(n, dtype=torch.int32, device=p.device)
                _lib.check(self._lib.s2b_containing_shapes_count_dev(self.index._h, _ptr(p), n, self.model, _ptr(cnt), _stream()))
                off = torch.zeros(n + 1, dtype=torch.int64, device=p.device)
                torch.cumsum(cnt, 0, out=off[1:])
                total = int(off[-1].item())
                off32 = off.to(torch.int32)
                ids = torch.empty(max(total, 1), dtype=torch.int32, device=p.device)
                if total:
                    _lib.check(self._lib.s2b_containing_shapes_fill_dev(self.index._h, _ptr(p), n, self.model, _ptr(off32), _ptr(ids), _stream()))
                return off32, ids[:total]
            off = np.empty(n + 1, np.uint32)
            total = ctypes.c_size_t(0)
            cap = max(1024, n // 8)
            while True:
                ids = np.empty(cap, np.int32)
                rc = self._lib.s2b_containing_shapes(self.index._h, _ptr(p), n, self.model, _ptr(off), _ptr(ids), cap, ctypes.byref(total))
                if rc == -4 and total.value > cap:
                    cap = total.value
                    continue
                _lib.check(rc)
                return off, ids[: total.value].copy()


def cell_union_contains(sorted_ids, xyz):
    """S2CellUnion::Contains(S2Point) per point; sorted_ids = S2CellUnion::cell_ids() (sorted, non-overlapping)."""
    lib = _lib.load()
    p = _points(xyz)
    n = p.shape[0]
    if _is_tensor(p):
        ids = sorted_ids if _is_tensor(sorted_ids) else torch.from_numpy(np.ascontiguousarray(sorted_ids, np.uint64).view(np.int64)).to(p.device)
        out = torch.empty(n, dtype=torch.uint8, device=p.device)
        _lib.check(lib.s2b_cellunion_contains_dev(_ptr(ids), ids.numel(), _ptr(p), n, _ptr(out), _stream()))
        return out
    ids = np.ascontiguousarray(sorted_ids, np.uint64)
    out = np.empty(n, np.uint8)
    _lib.check(lib.s2b_cellunion_contains(_ptr(ids), ids.size, _ptr(p), n, _ptr(out)))
    return out


def count_flags(flags, out=None):
    """Number of non-zero entries of a uint8 CUDA tensor of flags (the hit count of a contains() batch), accumulated into
    `out` (int64 CUDA tensor of one element; allocated and zeroed when None). No host synchronisation."""
    lib = _lib.load()
    if out is None:
        out = torch.zeros(1, dtype=torch.int64, device=flags.device)
    with torch.cuda.device(flags.device):
        _lib.check(lib.s2b_count_flags_dev(_ptr(flags), flags.numel(), _ptr(out), _stream()))
    return out


def exact_stage_count(reset=False):
    """How many predicate evaluations ran the exact (superaccumulator) stage on the device so far."""
    v = ctypes.c_uint64(0)
    _lib.check(_lib.load().s2b_exact_stage_count(ctypes.byref(v), 1 if reset else 0))
    return int(v.value)
