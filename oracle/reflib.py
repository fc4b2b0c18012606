"""ctypes wrapper over oracle/_ref/libs2ref.so — the UNMODIFIED reference (google/s2geometry) compiled from
/root/reference/src against oracle/absl_shim by oracle/build_ref.sh, behind the C ABI in oracle/ref_capi.cc.

TEST INFRASTRUCTURE ONLY: the checker for parity tests and the timed CPU baseline. Never on the product path.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libs2ref.so")

_vp = ctypes.c_void_p
_sz = ctypes.c_size_t


def _p(a):
    return _vp(a.ctypes.data) if a is not None else None


def build():
    """(Re)builds libs2ref.so when the reference sources are present (development container only)."""
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])


class Ref:
    def __init__(self, lib):
        self.lib = lib
        lib.ref_index_create.restype = _vp
        lib.ref_cover_cap.restype = _sz
        lib.ref_cellunion_normalize.restype = _sz
        lib.ref_containing_shapes.restype = ctypes.c_uint64
        lib.ref_cellid_from_face_pos_level.restype = ctypes.c_uint64
        lib.ref_cellid_from_face_pos_level.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int]

    def hardware_threads(self):
        return int(self.lib.ref_hardware_threads())

    # ---- cell ids
    def encode_points(self, xyz, level=30, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty(len(xyz), np.uint64)
        self.lib.ref_encode_points(_p(xyz), _sz(len(xyz)), level, _p(out), threads)
        return out

    def encode_latlng(self, ll, level=30, threads=0):
        ll = np.ascontiguousarray(ll, np.float64)
        out = np.empty(len(ll), np.uint64)
        self.lib.ref_encode_latlng(_p(ll), _sz(len(ll)), level, _p(out), threads)
        return out

    def encode_latlng_deg(self, ll, level=30, threads=0):
        ll = np.ascontiguousarray(ll, np.float64)
        out = np.empty(len(ll), np.uint64)
        self.lib.ref_encode_latlng_deg(_p(ll), _sz(len(ll)), level, _p(out), threads)
        return out

    def encode_latlng_e7(self, e7, level=30, threads=0):
        e7 = np.ascontiguousarray(e7, np.int32)
        out = np.empty(len(e7), np.uint64)
        self.lib.ref_encode_latlng_e7(_p(e7), _sz(len(e7)), level, _p(out), threads)
        return out

    def encode_latlng_levels_tokens(self, ll, levels, token_level_index, threads=0):
        ll = np.ascontiguousarray(ll, np.float64)
        n = len(ll)
        lv = (ctypes.c_int * len(levels))(*levels)
        ids = np.empty((len(levels), n), np.uint64); tok = np.empty((n, 16), np.uint8); ln = np.empty(n, np.uint8)
        self.lib.ref_encode_latlng_levels_tokens(_p(ll), _sz(n), lv, len(levels), _p(ids), token_level_index, _p(tok), _p(ln), threads)
        return ids, tok, ln

    def robust_cross_prod(self, a, b):
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 3); b = np.ascontiguousarray(b, np.float64).reshape(-1, 3)
        out = np.empty_like(a)
        self.lib.ref_robust_cross_prod(_p(a), _p(b), _sz(len(a)), _p(out))
        return out

    def latlng_to_point(self, ll):
        ll = np.ascontiguousarray(ll, np.float64)
        out = np.empty((len(ll), 3), np.float64)
        self.lib.ref_latlng_to_point(_p(ll), _sz(len(ll)), _p(out))
        return out

    def parent(self, ids, level):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty_like(ids)
        self.lib.ref_cellid_parent(_p(ids), _sz(ids.size), level, _p(out))
        return out

    def level(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty(ids.size, np.int32)
        self.lib.ref_cellid_level(_p(ids), _sz(ids.size), _p(out))
        return out

    def id_range(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        lo = np.empty_like(ids); hi = np.empty_like(ids)
        self.lib.ref_cellid_range(_p(ids), _sz(ids.size), _p(lo), _p(hi))
        return lo, hi

    def to_token(self, ids, threads=1):
        ids = np.ascontiguousarray(ids, np.uint64)
        tok = np.empty((ids.size, 16), np.uint8); ln = np.empty(ids.size, np.uint8)
        self.lib.ref_cellid_to_token(_p(ids), _sz(ids.size), _p(tok), _p(ln), threads)
        return tok, ln

    def from_token(self, tok, ln):
        tok = np.ascontiguousarray(tok, np.uint8); ln = np.ascontiguousarray(ln, np.uint8)
        out = np.empty(len(ln), np.uint64)
        self.lib.ref_cellid_from_token(_p(tok), _p(ln), _sz(len(ln)), _p(out))
        return out

    def to_point(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty((ids.size, 3), np.float64)
        self.lib.ref_cellid_to_point(_p(ids), _sz(ids.size), _p(out))
        return out

    def edge_neighbors(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty((ids.size, 4), np.uint64)
        self.lib.ref_edge_neighbors(_p(ids), _sz(ids.size), _p(out))
        return out

    def cellid_to_latlng(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty((ids.size, 2), np.float64)
        self.lib.ref_cellid_to_latlng(_p(ids), _sz(ids.size), _p(out))
        return out

    def vertex_neighbors(self, ids, level):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty((ids.size, 4), np.uint64); cnt = np.empty(ids.size, np.uint8)
        self.lib.ref_vertex_neighbors(_p(ids), _sz(ids.size), int(level), _p(out), _p(cnt))
        return out, cnt

    def all_neighbors(self, ids, nbr_level, stride=8):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty((ids.size, stride), np.uint64); cnt = np.empty(ids.size, np.uint32)
        self.lib.ref_all_neighbors(_p(ids), _sz(ids.size), int(nbr_level), _p(out), _sz(stride), _p(cnt))
        return out, cnt

    def cellunion_contains_cells(self, union_ids, ids):
        u = np.ascontiguousarray(union_ids, np.uint64); ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty(ids.size, np.uint8)
        self.lib.ref_cellunion_contains_cells(_p(u), _sz(u.size), _p(ids), _sz(ids.size), _p(out))
        return (out & 1) != 0, (out & 2) != 0

    def from_face_pos_level(self, face, pos, level):
        return int(self.lib.ref_cellid_from_face_pos_level(face, pos, level))

    # ---- predicates
    def sign(self, abc, kind=0):
        abc = np.ascontiguousarray(abc, np.float64).reshape(-1, 9)
        out = np.empty(len(abc), np.int8)
        self.lib.ref_sign(_p(abc), _sz(len(abc)), kind, _p(out))
        return out

    def crossing(self, abcd, kind=0):
        abcd = np.ascontiguousarray(abcd, np.float64).reshape(-1, 12)
        out = np.empty(len(abcd), np.int8)
        self.lib.ref_crossing(_p(abcd), _sz(len(abcd)), kind, _p(out))
        return out

    def normalize(self, a):
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 3)
        out = np.empty_like(a)
        self.lib.ref_normalize(_p(a), _sz(len(a)), _p(out))
        return out

    def ortho(self, a):
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 3)
        out = np.empty_like(a)
        self.lib.ref_ortho(_p(a), _sz(len(a)), _p(out))
        return out

    # ---- index
    def point_index(self, points):
        return RefPointIndex(self, points)

    def chord_length2(self, angle_rad):
        self.lib.ref_chord_length2.restype = ctypes.c_double
        return float(self.lib.ref_chord_length2(ctypes.c_double(angle_rad)))

    def tagged_shapes_encode(self, soup, compact=False, classic=False):
        """The soup as S2PointVectorShape / S2LaxPolylineShape / S2LaxPolygonShape — or, with classic=True, polygons as
        S2Polygon::Shape (loops through S2Polygon::InitOriented) and polylines as S2Polyline::Shape — encoded as
        EncodeTaggedShapes does (hint FAST, or COMPACT)."""
        self.lib.ref_tagged_shapes_encode.restype = _sz
        args = (int(soup.num_shapes), _p(soup.shape_dim), _p(soup.shape_chain_begin), _p(soup.chain_vertex_begin),
                _p(np.ascontiguousarray(soup.vertices, np.float64)), int(bool(compact)) | (2 if classic else 0))
        n = int(self.lib.ref_tagged_shapes_encode(*args, None, _sz(0)))
        out = np.zeros(max(n, 1), np.uint8)
        self.lib.ref_tagged_shapes_encode(*args, _p(out), _sz(out.size))
        return out[:n].tobytes()

    def tagged_shapes_decode_edges(self, data, max_shapes=1 << 20):
        """Decodes with the reference's Init methods -> (dims, num_edges per shape, edges (E,6)) or None if rejected."""
        self.lib.ref_tagged_shapes_decode_edges.restype = _sz
        buf = np.frombuffer(bytes(data), np.uint8)
        dims = np.zeros(max_shapes, np.int32); ne = np.zeros(max_shapes, np.uint32); ns = ctypes.c_int(0)
        cap = 1 << 16
        while True:
            edges = np.zeros((cap, 6), np.float64)
            total = self.lib.ref_tagged_shapes_decode_edges(_p(buf), _sz(buf.size), int(max_shapes), _p(dims), _p(ne), _p(edges), _sz(cap), ctypes.byref(ns))
            if total == ctypes.c_size_t(-1).value:
                return None
            if total <= cap:
                return dims[: ns.value].copy(), ne[: ns.value].copy(), edges[:total].copy()
            cap = int(total)

    def index_create(self, soup, max_edges_per_cell=0):
        return RefIndex(self, soup, max_edges_per_cell)

    # ---- edge clipping (S2ShapeIndexRegion::AnyEdgeIntersects)
    def clip_to_padded_face(self, a, b, face, padding):
        a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64); face = np.ascontiguousarray(face, np.int32)
        uv = np.zeros((len(a), 4)); ok = np.zeros(len(a), np.uint8)
        self.lib.ref_clip_to_padded_face(_p(a), _p(b), _p(face), _sz(len(a)), ctypes.c_double(padding), _p(uv), _p(ok))
        return uv, ok

    def intersects_rect(self, uv4, rect4):
        uv4 = np.ascontiguousarray(uv4, np.float64); rect4 = np.ascontiguousarray(rect4, np.float64)
        out = np.zeros(len(uv4), np.uint8)
        self.lib.ref_intersects_rect(_p(uv4), _p(rect4), _sz(len(uv4)), _p(out))
        return out

    # ---- cell unions
    def cellunion_contains(self, ids, xyz, threads=0):
        ids = np.ascontiguousarray(ids, np.uint64); xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty(len(xyz), np.uint8)
        self.lib.ref_cellunion_contains(_p(ids), _sz(ids.size), _p(xyz), _sz(len(xyz)), _p(out), threads)
        return out

    def cellunion_normalize(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty_like(ids)
        k = self.lib.ref_cellunion_normalize(_p(ids), _sz(ids.size), _p(out))
        return out[:k].copy()

    def cover_cap(self, center, radius_rad, max_cells=8, min_level=-1, max_level=-1):
        center = np.ascontiguousarray(center, np.float64)
        out = np.empty(max(4 * max_cells, 64), np.uint64)
        k = self.lib.ref_cover_cap(_p(center), ctypes.c_double(radius_rad), max_cells, min_level, max_level, _p(out), _sz(out.size))
        assert k <= out.size
        return out[:k].copy()


class RefPointIndex:
    """S2PointIndex<int> with data = position, queried through S2ClosestPointQuery<int> (the reference's own)."""

    def __init__(self, ref, points):
        self.lib = ref.lib
        self.lib.ref_point_index_create.restype = _vp
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        self.h = self.lib.ref_point_index_create(_p(pts), _sz(len(pts)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ref_point_index_destroy(_vp(self.h))
                self.h = None
        except Exception:
            pass

    def count(self, targets, max_distance_rad=-1.0, threads=0):
        t = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
        out = np.zeros(len(t), np.uint32)
        self.lib.ref_closest_points_count(_vp(self.h), _p(t), _sz(len(t)), ctypes.c_double(max_distance_rad), _p(out), int(threads))
        return out

    def closest(self, targets, max_distance_rad=-1.0, threads=0):
        t = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
        data = np.zeros(len(t), np.int32); l2 = np.zeros(len(t), np.float64)
        self.lib.ref_closest_point(_vp(self.h), _p(t), _sz(len(t)), ctypes.c_double(max_distance_rad), _p(data), _p(l2), int(threads))
        return data, l2


def _closest_k(self, targets, k, max_distance_rad=-1.0, threads=0):
    t = np.ascontiguousarray(targets, np.float64).reshape(-1, 3)
    data = np.zeros((len(t), k), np.int32); l2 = np.zeros((len(t), k), np.float64); cnt = np.zeros(len(t), np.uint32)
    self.lib.ref_closest_points(_vp(self.h), _p(t), _sz(len(t)), int(k), ctypes.c_double(max_distance_rad), _p(data), _p(l2), _p(cnt), int(threads))
    return data, l2, cnt


RefPointIndex.closest_k = _closest_k


def _closest_targets(self, kind, targets, k, max_distance_rad=-1.0, max_error_rad=0.0, threads=0):
    """FindClosestPoints for PointTarget (kind 0, (n,3)), EdgeTarget (1, (n,6)) or CellTarget (2, uint64 ids); k == 0: counts only."""
    t = np.ascontiguousarray(targets, np.uint64 if kind == 2 else np.float64)
    n = t.size if kind == 2 else len(t.reshape(-1, 3 if kind == 0 else 6))
    kk = max(k, 1)
    data = np.zeros((n, kk), np.int32); l2 = np.zeros((n, kk), np.float64); cnt = np.zeros(n, np.uint32)
    self.lib.ref_closest_points_target(_vp(self.h), int(kind), _p(t), _sz(n), int(k), ctypes.c_double(max_distance_rad), ctypes.c_double(max_error_rad),
                                       _p(data), _p(l2), _p(cnt), int(threads))
    return data, l2, cnt


RefPointIndex.closest_targets = _closest_targets


def _closest_shape_index(self, ref_index, k=0, max_distance_rad=-1.0, include_interiors=True, max_error_rad=0.0):
    """FindClosestPoints(ShapeIndexTarget(ref_index.index)) -> (data, length2), nearest first; k == 0: unlimited."""
    self.lib.ref_closest_points_shape_index.restype = _sz
    cap = k if k else 1 << 16
    while True:
        data = np.zeros(cap, np.int32); l2 = np.zeros(cap, np.float64)
        n = self.lib.ref_closest_points_shape_index(_vp(self.h), _vp(ref_index.h), 1 if include_interiors else 0, int(k), ctypes.c_double(max_distance_rad),
                                                    ctypes.c_double(max_error_rad), _p(data), _p(l2), _sz(cap))
        if n <= cap:
            return data[:n].copy(), l2[:n].copy()
        cap = int(n)


RefPointIndex.closest_shape_index = _closest_shape_index


class RefIndex:
    """A MutableS2ShapeIndex built by the reference from a shape soup (see s2geometry_b200.shapes.ShapeSoup)."""

    def __init__(self, ref, soup, max_edges_per_cell=0):
        self.ref = ref
        self.lib = ref.lib
        self._keep = soup
        self.h = self.lib.ref_index_create(int(soup.num_shapes), _p(soup.shape_dim), _p(soup.shape_chain_begin),
                                           _p(soup.chain_vertex_begin), _p(soup.vertices), int(max_edges_per_cell))
        self.num_shape_ids = int(self.lib.ref_index_num_shape_ids(_vp(self.h)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ref_index_destroy(_vp(self.h))
                self.h = None
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    def flatten(self):
        C = ctypes.c_uint64(); K = ctypes.c_uint64(); E = ctypes.c_uint64()
        self.lib.ref_index_flat_sizes(_vp(self.h), ctypes.byref(C), ctypes.byref(K), ctypes.byref(E))
        C, K, E = C.value, K.value, E.value
        f = dict(
            cell_ids=np.empty(C, np.uint64), cell_center=np.empty((C, 3), np.float64),
            cell_clip_begin=np.empty(C + 1, np.uint32), clip_shape_id=np.empty(K, np.int32),
            clip_flags=np.empty(K, np.uint8), clip_edge_begin=np.empty(K + 1, np.uint32),
            edge_id=np.empty(E, np.int32), edge_v0=np.empty((E, 3), np.float64), edge_v1=np.empty((E, 3), np.float64),
            num_shape_ids=self.num_shape_ids)
        self.lib.ref_index_flatten(_vp(self.h), _p(f["cell_ids"]), _p(f["cell_center"]), _p(f["cell_clip_begin"]),
                                   _p(f["clip_shape_id"]), _p(f["clip_flags"]), _p(f["clip_edge_begin"]),
                                   _p(f["edge_id"]), _p(f["edge_v0"]), _p(f["edge_v1"]))
        return f

    def locate_points(self, xyz, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        at = np.empty(len(xyz), np.uint64)
        self.lib.ref_index_locate_points(_vp(self.h), _p(xyz), _sz(len(xyz)), _p(at), int(threads))
        return at

    def locate_cells(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        rel = np.empty(ids.size, np.uint8); at = np.empty(ids.size, np.uint64)
        self.lib.ref_index_locate_cells(_vp(self.h), _p(ids), _sz(ids.size), _p(rel), _p(at))
        return rel, at

    def region_cells(self, ids, mode):
        """S2ShapeIndexRegion::MayIntersect (mode 0) / Contains (mode 1) of S2Cell(ids[i])."""
        ids = np.ascontiguousarray(ids, np.uint64)
        out = np.empty(ids.size, np.uint8)
        self.lib.ref_region_cells(_vp(self.h), _p(ids), _sz(ids.size), int(mode), _p(out))
        return out

    def region_cell_union_bound(self):
        out = np.zeros(8, np.uint64)
        n = int(self.lib.ref_region_cell_union_bound(_vp(self.h), _p(out)))
        return out[:n].copy()

    def region_covering(self, level):
        """S2RegionCoverer (min_level = max_level = level) over S2ShapeIndexRegion -> sorted cell ids."""
        self.lib.ref_region_covering.restype = _sz
        n = int(self.lib.ref_region_covering(_vp(self.h), int(level), None, _sz(0)))
        out = np.zeros(max(n, 1), np.uint64)
        self.lib.ref_region_covering(_vp(self.h), int(level), _p(out), _sz(n))
        return out[:n]

    def covering(self, max_cells):
        """S2RegionCoverer(max_cells).GetCovering(S2ShapeIndexRegion(index)) -> sorted cell ids (an S2CellUnion)."""
        self.lib.ref_index_covering.restype = _sz
        out = np.zeros(4 * max_cells + 64, np.uint64)
        n = int(self.lib.ref_index_covering(_vp(self.h), int(max_cells), _p(out), _sz(out.size)))
        assert n <= out.size
        return out[:n].copy()

    def region_intersecting_shapes(self, ids):
        """S2ShapeIndexRegion::VisitIntersectingShapeIds per target: (offsets, shape ids, contains flags), sorted by shape id."""
        ids = np.ascontiguousarray(ids, np.uint64)
        self.lib.ref_region_intersecting_shapes.restype = _sz
        off = np.zeros(ids.size + 1, np.uint32)
        total = int(self.lib.ref_region_intersecting_shapes(_vp(self.h), _p(ids), _sz(ids.size), _p(off), None, None, _sz(0)))
        shp = np.zeros(max(total, 1), np.int32); con = np.zeros(max(total, 1), np.uint8)
        self.lib.ref_region_intersecting_shapes(_vp(self.h), _p(ids), _sz(ids.size), _p(off), _p(shp), _p(con), _sz(total))
        return off, shp[:total], con[:total]

    def encode(self):
        """MutableS2ShapeIndex::Encode -> bytes."""
        self.lib.ref_index_encode.restype = _sz
        n = int(self.lib.ref_index_encode(_vp(self.h), None, _sz(0)))
        out = np.zeros(max(n, 1), np.uint8)
        self.lib.ref_index_encode(_vp(self.h), _p(out), _sz(out.size))
        return out[:n].tobytes()

    def init_matches(self, data):
        """True if MutableS2ShapeIndex::Init accepts `data` (with this index's shapes) and reproduces this index's cells."""
        buf = np.frombuffer(bytes(data), np.uint8)
        return bool(self.lib.ref_index_init_matches(_vp(self.h), _p(buf), _sz(buf.size)))

    def contains(self, xyz, model=1, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty(len(xyz), np.uint8)
        self.lib.ref_contains(_vp(self.h), _p(xyz), _sz(len(xyz)), model, _p(out), threads)
        return out

    def shape_contains(self, shape_id, xyz, model=1, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty(len(xyz), np.uint8)
        self.lib.ref_shape_contains(_vp(self.h), shape_id, _p(xyz), _sz(len(xyz)), model, _p(out), threads)
        return out

    def point_distances(self, xyz, max_distance_rad=-1.0, include_interiors=True, threads=0):
        """S2ClosestEdgeQuery::GetDistance(PointTarget(p)).length2() per point (inf if not below max_distance)."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty(len(xyz), np.float64)
        self.lib.ref_index_point_distances(_vp(self.h), _p(xyz), _sz(len(xyz)), 1 if include_interiors else 0, ctypes.c_double(max_distance_rad), _p(out), int(threads))
        return out

    def shape_histogram(self, xyz, model=1, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.zeros(self.num_shape_ids, np.uint64)
        self.lib.ref_shape_histogram(_vp(self.h), _p(xyz), _sz(len(xyz)), model, _p(out), threads)
        return out

    def containing_shapes(self, xyz, model=1):
        xyz = np.ascontiguousarray(xyz, np.float64)
        off = np.empty(len(xyz) + 1, np.uint32)
        cap = max(1024, 4 * len(xyz))
        while True:
            ids = np.empty(cap, np.int32)
            total = self.lib.ref_containing_shapes(_vp(self.h), _p(xyz), _sz(len(xyz)), model, _p(off), _p(ids), ctypes.c_uint64(cap))
            if total <= cap:
                return off, ids[:total].copy()
            cap = int(total)

    def incident_edges(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64)
        self.lib.ref_incident_edges.restype = ctypes.c_uint64
        off = np.empty(len(xyz) + 1, np.uint32)
        cap = max(1024, len(xyz))
        while True:
            sid = np.empty(cap, np.int32); eid = np.empty(cap, np.int32)
            total = self.lib.ref_incident_edges(_vp(self.h), _p(xyz), _sz(len(xyz)), _p(off), _p(sid), _p(eid), ctypes.c_uint64(cap))
            if total <= cap:
                return off, sid[:total].copy(), eid[:total].copy()
            cap = int(total)

    def brute_force_contains(self, shape_id, xyz, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64)
        out = np.empty(len(xyz), np.uint8)
        self.lib.ref_brute_force_contains(_vp(self.h), shape_id, _p(xyz), _sz(len(xyz)), _p(out), threads)
        return out


_ref = None


def load():
    global _ref
    if _ref is None:
        if not os.path.exists(LIB):
            build()
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/_ref/libs2ref.so is missing and the reference sources are not available to build it")
        _ref = Ref(ctypes.CDLL(LIB))
    return _ref


def _csr(targets):
    """list of uint64 arrays (or one array of single-cell targets) -> (ids, uint64 offsets)"""
    if isinstance(targets, np.ndarray) and targets.ndim == 1:
        ids = np.ascontiguousarray(targets, np.uint64)
        return ids, np.arange(len(ids) + 1, dtype=np.uint64)
    lens = [len(t) for t in targets]
    ids = np.ascontiguousarray(np.concatenate([np.asarray(t, np.uint64) for t in targets]) if lens else np.zeros(0, np.uint64), np.uint64)
    return ids, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)


class RefCellIndex:
    """S2CellIndex (the reference's own, s2cell_index.h) over (cell_id, label) pairs."""

    def __init__(self, ref, cell_ids, labels):
        self.lib = ref.lib
        self.lib.ref_cell_index_create.restype = _vp
        self.lib.ref_cell_index_visit.restype = ctypes.c_uint64
        self.lib.ref_cell_index_labels.restype = ctypes.c_uint64
        ids = np.ascontiguousarray(cell_ids, np.uint64)
        lab = np.ascontiguousarray(labels, np.int32)
        self.num_labels = int(lab.max()) + 1 if lab.size else 0
        self.h = self.lib.ref_cell_index_create(_p(ids), _p(lab), _sz(len(ids)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ref_cell_index_destroy(_vp(self.h))
                self.h = None
        except Exception:
            pass

    def visit(self, targets):
        """VisitIntersectingCells per target -> (offsets, cells, labels), each target's pairs sorted by (cell, label)."""
        ids, off = _csr(targets)
        T = len(off) - 1
        out_off = np.zeros(T + 1, np.uint64)
        cap = 1 << 16
        while True:
            cells = np.empty(cap, np.uint64); labels = np.empty(cap, np.int32)
            total = self.lib.ref_cell_index_visit(_vp(self.h), _p(ids), _p(off), _sz(T), _p(out_off), _p(cells), _p(labels), ctypes.c_uint64(cap))
            if total <= cap:
                return out_off, cells[:total].copy(), labels[:total].copy()
            cap = int(total)

    def labels(self, targets):
        """GetIntersectingLabels per target -> (offsets, labels ascending)."""
        ids, off = _csr(targets)
        T = len(off) - 1
        out_off = np.zeros(T + 1, np.uint64)
        cap = 1 << 16
        while True:
            labels = np.empty(cap, np.int32)
            total = self.lib.ref_cell_index_labels(_vp(self.h), _p(ids), _p(off), _sz(T), _p(out_off), _p(labels), ctypes.c_uint64(cap))
            if total <= cap:
                return out_off, labels[:total].copy()
            cap = int(total)

    def point_histogram(self, xyz, threads=0):
        xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
        counts = np.zeros(max(self.num_labels, 1), np.uint64)
        self.lib.ref_cell_index_point_histogram(_vp(self.h), _p(xyz), _sz(len(xyz)), _p(counts), int(self.num_labels), int(threads))
        return counts[: self.num_labels]


def ref_sharder_most_intersecting(ref, shards, regions, default_shard=-1):
    """S2RegionSharder(shards).GetMostIntersectingShard(S2CellUnion region, default) for every region."""
    s_ids, s_off = _csr(shards)
    r_ids, r_off = _csr(regions)
    out = np.empty(len(r_off) - 1, np.int32)
    ref.lib.ref_sharder_most_intersecting(_p(s_ids), _p(s_off), _sz(len(s_off) - 1), _p(r_ids), _p(r_off), _sz(len(r_off) - 1),
                                          int(default_shard), _p(out))
    return out


def ref_cellunion_bound(ref, region):
    """S2CellUnion(region.GetCellUnionBound()) for an S2CellUnion region: the covering S2RegionSharder queries its index with."""
    ids = np.ascontiguousarray(region, np.uint64)
    out = np.empty(64, np.uint64)
    ref.lib.ref_cellunion_cell_union_bound.restype = _sz
    k = ref.lib.ref_cellunion_cell_union_bound(_p(ids), _sz(len(ids)), _p(out), _sz(out.size))
    assert k <= out.size
    return out[:k].copy()
