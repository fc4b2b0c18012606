"""Shape soup: the geometry that goes into a shape index, as flat arrays.

Mirrors the reference's S2Shape plugin interface (src/s2/s2shape.h:64-260: num_edges / edge(e) / dimension /
chains) for the three lax shape kinds the path needs:
  dimension 2  closed loops, one chain per loop, interior on the left (S2LaxPolygonShape, s2lax_polygon_shape.h)
  dimension 1  polylines, n vertices -> n-1 edges                     (S2LaxPolylineShape)
  dimension 0  points, every vertex a degenerate edge (v, v)          (S2PointVectorShape)
Edge ids run over a shape's chains in order, as in the reference.
"""
import numpy as np


class ShapeSoup:
    def __init__(self):
        self._dims = []
        self._chains_per_shape = []
        self._chains = []

    def add_polygon(self, loops):
        """loops: list of (k,3) float64 arrays of unit vectors; CCW = shell, CW = hole. Returns the shape id."""
        return self._add(2, loops)

    def add_polyline(self, vertices):
        return self._add(1, [vertices])

    def add_points(self, points):
        return self._add(0, [points])

    def _add(self, dim, chains):
        chains = [np.ascontiguousarray(c, np.float64).reshape(-1, 3) for c in chains]
        self._dims.append(dim)
        self._chains_per_shape.append(len(chains))
        self._chains.extend(chains)
        return len(self._dims) - 1

    @property
    def num_shapes(self):
        return len(self._dims)

    # flat arrays in the layout the C ABI (and the oracle wrapper) take
    @property
    def shape_dim(self):
        return np.asarray(self._dims, np.int8)

    @property
    def shape_chain_begin(self):
        return np.concatenate([[0], np.cumsum(self._chains_per_shape)]).astype(np.uint32)

    @property
    def chain_vertex_begin(self):
        return np.concatenate([[0], np.cumsum([len(c) for c in self._chains])]).astype(np.uint32)

    @property
    def vertices(self):
        if not self._chains:
            return np.zeros((0, 3), np.float64)
        return np.ascontiguousarray(np.concatenate(self._chains), np.float64)

    def freeze(self):
        """A snapshot whose arrays are materialised once (attribute access above rebuilds them)."""
        return FrozenSoup(self)


class FrozenSoup:
    def __init__(self, soup):
        self.num_shapes = soup.num_shapes
        self.shape_dim = soup.shape_dim
        self.shape_chain_begin = soup.shape_chain_begin
        self.chain_vertex_begin = soup.chain_vertex_begin
        self.vertices = soup.vertices

    def freeze(self):
        return self


def decode_shapes(data):
    """Reads s2shapeutil::FastEncodeTaggedShapes / CompactEncodeTaggedShapes bytes (S2PointVectorShape, S2LaxPolylineShape,
    S2LaxPolygonShape) into a FrozenSoup. Returns (soup, bytes_consumed)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    buf = np.frombuffer(bytes(data), np.uint8)
    h = ctypes.c_void_p(); used = ctypes.c_size_t(0)
    _lib.check(lib.s2b_shapes_decode(buf.ctypes.data if buf.size else None, int(buf.size), ctypes.byref(h), ctypes.byref(used)))
    try:
        ns = ctypes.c_int(0); nv = ctypes.c_size_t(0)
        pd, pc, pv, px = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(lib.s2b_shape_soup_arrays(h, ctypes.byref(ns), ctypes.byref(pd), ctypes.byref(pc), ctypes.byref(pv), ctypes.byref(px), ctypes.byref(nv)))

        def arr(ptr, count, dtype):
            if not ptr.value or count == 0:
                return np.zeros(count, dtype)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(count,)).copy()
        soup = FrozenSoup.__new__(FrozenSoup)
        soup.num_shapes = ns.value
        soup.shape_dim = arr(pd, ns.value, np.int8)
        soup.shape_chain_begin = arr(pc, ns.value + 1, np.uint32)
        soup.chain_vertex_begin = arr(pv, int(soup.shape_chain_begin[-1]) + 1, np.uint32)
        soup.vertices = arr(px, 3 * nv.value, np.float64).reshape(-1, 3)
        return soup, int(used.value)
    finally:
        lib.s2b_shape_soup_destroy(h)


def encode_shapes(soup, compact=False):
    """Serializes a soup as s2shapeutil::FastEncodeTaggedShapes (or, with compact=True, CompactEncodeTaggedShapes) would
    serialize the corresponding lax shapes."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    soup = soup.freeze()
    v = np.ascontiguousarray(soup.vertices, np.float64)
    args = (1 if compact else 0, int(soup.num_shapes), soup.shape_dim.ctypes.data, soup.shape_chain_begin.ctypes.data, soup.chain_vertex_begin.ctypes.data,
            v.ctypes.data if v.size else None)
    size = ctypes.c_size_t(0)
    rc = lib.s2b_shapes_encode_hint(*args, None, 0, ctypes.byref(size))
    if rc != _lib.S2B_ECAPACITY:
        _lib.check(rc)
    out = np.zeros(max(int(size.value), 1), np.uint8)
    _lib.check(lib.s2b_shapes_encode_hint(*args, out.ctypes.data, int(out.size), ctypes.byref(size)))
    return out[: int(size.value)].tobytes()
