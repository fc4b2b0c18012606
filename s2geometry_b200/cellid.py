"""Batched S2CellId operations — the host-side mirror of the reference's S2CellId interface for the hot path.

Reference calls replaced (src/s2/): S2CellId(const S2Point&) s2cell_id.cc:309-315, S2CellId(const S2LatLng&)
:317, parent(level) s2cell_id.h:662-668, ToToken s2cell_id.cc:217-233. Names follow the pybind11 binding
(src/python/s2cell_id_bindings.cc): from_point(s) / from_lat_lng / parent / to_token.

Inputs may be NumPy arrays (host path: the C ABI stages them through the GPU in chunks) or CUDA torch tensors
(device path: kernels are enqueued on torch's current stream, results stay on the device).
"""
import ctypes

import numpy as np

from . import _lib

try:  # torch is plumbing only (device memory, streams); numpy inputs work without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _np_in(a, dtype, cols):
    a = np.ascontiguousarray(a, dtype=dtype)
    if cols and (a.ndim != 2 or a.shape[1] != cols):
        raise ValueError("expected an (N, %d) array" % cols)
    return a


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else ctypes.c_void_p(a.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _tensor_in(t, dtype, cols):
    if not t.is_cuda:
        raise ValueError("torch inputs must be CUDA tensors (use NumPy arrays for host data)")
    if t.dtype != dtype:
        raise ValueError("expected dtype %s, got %s" % (dtype, t.dtype))
    t = t.contiguous()
    if cols and (t.dim() != 2 or t.shape[1] != cols):
        raise ValueError("expected an (N, %d) tensor" % cols)
    return t


def _levels_arg(levels):
    lv = [int(l) for l in levels]
    if not 1 <= len(lv) <= 4:
        raise ValueError("between 1 and 4 levels")
    return (ctypes.c_int * len(lv))(*lv), len(lv)


LATLNG_HOST_LIBM, LATLNG_CORRECTLY_ROUNDED = 0, 1   # include/s2b_capi.h


def set_lat_lng_mode(mode):
    """Process-wide definition of S2CellId(S2LatLng): LATLNG_HOST_LIBM (default; ids equal the reference's evaluated
    with this process's libm, bit for bit) or LATLNG_CORRECTLY_ROUNDED (platform independent). Returns the old mode."""
    old = _lib.load().s2b_set_latlng_mode(int(mode))
    if old < 0:
        _lib.check(old)
    return old


def lat_lng_strict_stats(reset=False):
    """(points re-evaluated with the host libm, points whose id changed) since the last reset."""
    a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.check(_lib.load().s2b_latlng_strict_stats(ctypes.byref(a), ctypes.byref(b), 1 if reset else 0))
    return int(a.value), int(b.value)


def _strict_dev(lib, kind, ll, n, lv, nl, ids, flags, token_level_index, tok, ln):
    # device path with the host-libm guarantee: synchronises the current stream (s2b_encode_latlng_strict_dev)
    _lib.check(lib.s2b_encode_latlng_strict_dev(kind, _ptr(ll), n, lv, nl, _ptr(ids), _ptr(flags) if flags is not None else None,
                                                int(token_level_index), _ptr(tok) if tok is not None else None,
                                                _ptr(ln) if ln is not None else None, _stream()))


def from_points(xyz, level=30):
    """S2CellId(S2Point) for every row of xyz (N,3) float64 -> uint64 ids at `level` (int64 tensor on the device path)."""
    lib = _lib.load()
    if _is_tensor(xyz):
        xyz = _tensor_in(xyz, torch.float64, 3)
        out = torch.empty(xyz.shape[0], dtype=torch.int64, device=xyz.device)
        with torch.cuda.device(xyz.device):
            _lib.check(lib.s2b_encode_points_dev(_ptr(xyz), xyz.shape[0], int(level), _ptr(out), _stream()))
        return out
    xyz = _np_in(xyz, np.float64, 3)
    out = np.empty(xyz.shape[0], dtype=np.uint64)
    _lib.check(lib.s2b_encode_points(_ptr(xyz), xyz.shape[0], int(level), _ptr(out)))
    return out


def from_lat_lng(latlng_rad, level=30, return_boundary_flags=False):
    """S2CellId(S2LatLng::FromRadians(lat, lng)) for every row of latlng_rad (N,2) float64."""
    lib = _lib.load()
    if _is_tensor(latlng_rad):
        ll = _tensor_in(latlng_rad, torch.float64, 2)
        n = ll.shape[0]
        out = torch.empty(n, dtype=torch.int64, device=ll.device)
        flags = torch.empty(n, dtype=torch.uint8, device=ll.device) if return_boundary_flags else None
        with torch.cuda.device(ll.device):
            lv, nl = _levels_arg([level])
            _strict_dev(lib, 1, ll, n, lv, nl, out, flags, -1, None, None)
        return (out, flags) if return_boundary_flags else out
    ll = _np_in(latlng_rad, np.float64, 2)
    n = ll.shape[0]
    out = np.empty(n, dtype=np.uint64)
    flags = np.empty(n, dtype=np.uint8) if return_boundary_flags else None
    _lib.check(lib.s2b_encode_latlng(_ptr(ll), n, int(level), _ptr(out), _ptr(flags) if flags is not None else None))
    return (out, flags) if return_boundary_flags else out


def from_lat_lng_e7(latlng_e7, level=30):
    """S2CellId(S2LatLng::FromE7(lat_e7, lng_e7)) for every row of latlng_e7 (N,2) int32."""
    lib = _lib.load()
    if _is_tensor(latlng_e7):
        ll = _tensor_in(latlng_e7, torch.int32, 2)
        out = torch.empty(ll.shape[0], dtype=torch.int64, device=ll.device)
        with torch.cuda.device(ll.device):
            lv, nl = _levels_arg([level])
            _strict_dev(lib, 2, ll, ll.shape[0], lv, nl, out, None, -1, None, None)
        return out
    ll = _np_in(latlng_e7, np.int32, 2)
    out = np.empty(ll.shape[0], dtype=np.uint64)
    _lib.check(lib.s2b_encode_latlng_e7(_ptr(ll), ll.shape[0], int(level), _ptr(out), None))
    return out


def from_lat_lng_levels(latlng_rad, levels, out=None):
    """Fused encode at several levels: returns (len(levels), N) ids, row k = S2CellId(ll).parent(levels[k])."""
    lib = _lib.load()
    lv, nl = _levels_arg(levels)
    if _is_tensor(latlng_rad):
        ll = _tensor_in(latlng_rad, torch.float64, 2)
        n = ll.shape[0]
        if out is None:
            out = torch.empty((nl, n), dtype=torch.int64, device=ll.device)
        with torch.cuda.device(ll.device):
            _strict_dev(lib, 1, ll, n, lv, nl, out, None, -1, None, None)
        return out
    ll = _np_in(latlng_rad, np.float64, 2)
    n = ll.shape[0]
    if out is None:
        out = np.empty((nl, n), dtype=np.uint64)
    _lib.check(lib.s2b_encode_latlng_levels(_ptr(ll), n, lv, nl, _ptr(out)))
    return out


def from_lat_lng_levels_tokens(latlng_rad, levels, token_level_index, out_ids=None, out_tokens=None, out_len=None):
    """Fused: ids at several levels plus ToToken of the id at levels[token_level_index], one pass.
    Returns (ids (L,N), tokens (N,16) uint8, lengths (N,) uint8)."""
    lib = _lib.load()
    lv, nl = _levels_arg(levels)
    if _is_tensor(latlng_rad):
        ll = _tensor_in(latlng_rad, torch.float64, 2)
        n = ll.shape[0]
        ids = out_ids if out_ids is not None else torch.empty((nl, n), dtype=torch.int64, device=ll.device)
        tok = out_tokens if out_tokens is not None else torch.empty((n, 16), dtype=torch.uint8, device=ll.device)
        ln = out_len if out_len is not None else torch.empty(n, dtype=torch.uint8, device=ll.device)
        with torch.cuda.device(ll.device):
            _strict_dev(lib, 1, ll, n, lv, nl, ids, None, token_level_index, tok, ln)
        return ids, tok, ln
    ll = _np_in(latlng_rad, np.float64, 2)
    n = ll.shape[0]
    ids = out_ids if out_ids is not None else np.empty((nl, n), dtype=np.uint64)
    tok = out_tokens if out_tokens is not None else np.empty((n, 16), dtype=np.uint8)
    ln = out_len if out_len is not None else np.empty(n, dtype=np.uint8)
    _lib.check(lib.s2b_encode_latlng_levels_tokens(_ptr(ll), n, lv, nl, _ptr(ids), int(token_level_index), _ptr(tok), _ptr(ln)))
    return ids, tok, ln


def from_points_levels(xyz, levels):
    """Device path only: fused S2CellId(S2Point).parent(level) at several levels -> (len(levels), N) int64 tensor."""
    lib = _lib.load()
    lv, nl = _levels_arg(levels)
    xyz = _tensor_in(xyz, torch.float64, 3)
    out = torch.empty((nl, xyz.shape[0]), dtype=torch.int64, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.s2b_encode_points_levels_dev(_ptr(xyz), xyz.shape[0], lv, nl, _ptr(out), _stream()))
    return out


def parent(ids, level):
    """S2CellId::parent(level) element-wise."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty_like(ids)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_parent_dev(_ptr(ids), ids.numel(), int(level), _ptr(out), _stream()))
        return out
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty_like(ids)
    _lib.check(lib.s2b_cellid_parent(_ptr(ids), ids.size, int(level), _ptr(out)))
    return out


def to_token(ids, out=None, out_len=None):
    """S2CellId::ToToken element-wise. Returns (tokens, lengths): tokens is (N,16) uint8, NUL padded."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        n = ids.numel()
        tok = out if out is not None else torch.empty((n, 16), dtype=torch.uint8, device=ids.device)
        ln = out_len if out_len is not None else torch.empty(n, dtype=torch.uint8, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_to_token_dev(_ptr(ids), n, _ptr(tok), _ptr(ln), _stream()))
        return tok, ln
    ids = _np_in(ids, np.uint64, 0)
    tok = np.empty((ids.size, 16), dtype=np.uint8)
    ln = np.empty(ids.size, dtype=np.uint8)
    _lib.check(lib.s2b_cellid_to_token(_ptr(ids), ids.size, _ptr(tok), _ptr(ln)))
    return tok, ln


def to_point(ids):
    """S2CellId::ToPoint element-wise -> (N,3) float64 unit vectors (cell centres)."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty((ids.numel(), 3), dtype=torch.float64, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_to_point_dev(_ptr(ids), ids.numel(), _ptr(out), _stream()))
        return out
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty((ids.size, 3), dtype=np.float64)
    _lib.check(lib.s2b_cellid_to_point(_ptr(ids), ids.size, _ptr(out)))
    return out


def to_lat_lng(ids):
    """S2CellId::ToLatLng element-wise -> (N,2) float64 radians (lat, lng); agrees with the reference to a few ulp."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty((ids.numel(), 2), dtype=torch.float64, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_to_latlng_dev(_ptr(ids), ids.numel(), _ptr(out), _stream()))
        return out
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty((ids.size, 2), dtype=np.float64)
    _lib.check(lib.s2b_cellid_to_latlng(_ptr(ids), ids.size, _ptr(out)))
    return out


def edge_neighbors(ids):
    """S2CellId::GetEdgeNeighbors element-wise -> (N,4) ids (down, right, up, left) at each cell's own level."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty((ids.numel(), 4), dtype=torch.int64, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_edge_neighbors_dev(_ptr(ids), ids.numel(), _ptr(out), _stream()))
        return out
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty((ids.size, 4), dtype=np.uint64)
    _lib.check(lib.s2b_cellid_edge_neighbors(_ptr(ids), ids.size, _ptr(out)))
    return out


def children(ids):
    """S2CellId::child(0..3) element-wise -> (N,4) ids in Hilbert order (0 for leaf cells and invalid ids)."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty((ids.numel(), 4), dtype=torch.int64, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_children_dev(_ptr(ids), ids.numel(), _ptr(out), _stream()))
        return out
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty((ids.size, 4), dtype=np.uint64)
    _lib.check(lib.s2b_cellid_children(_ptr(ids), ids.size, _ptr(out)))
    return out


def advance(ids, steps, wrap=False):
    """S2CellId::advance(steps) (clamped at the ends of the level) or, with wrap=True, advance_wrap(steps), element-wise;
    next() / prev() are steps of +1 / -1."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        st = _tensor_in(torch.as_tensor(steps, dtype=torch.int64, device=ids.device).expand(ids.numel()).contiguous(), torch.int64, 0)
        out = torch.empty_like(ids)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_advance_dev(_ptr(ids), _ptr(st), ids.numel(), 1 if wrap else 0, _ptr(out), _stream()))
        return out
    ids = _np_in(ids, np.uint64, 0)
    st = np.ascontiguousarray(np.broadcast_to(np.asarray(steps, np.int64), ids.shape))
    out = np.empty_like(ids)
    _lib.check(lib.s2b_cellid_advance(_ptr(ids), _ptr(st), ids.size, 1 if wrap else 0, _ptr(out)))
    return out


def common_ancestor_level(a, b):
    """S2CellId::GetCommonAncestorLevel element-wise -> int8 (-1: different faces or an invalid id)."""
    lib = _lib.load()
    if _is_tensor(a):
        a = _tensor_in(a, torch.int64, 0); b = _tensor_in(b, torch.int64, 0)
        out = torch.empty(a.numel(), dtype=torch.int8, device=a.device)
        with torch.cuda.device(a.device):
            _lib.check(lib.s2b_cellid_common_ancestor_level_dev(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream()))
        return out
    a = _np_in(a, np.uint64, 0); b = _np_in(b, np.uint64, 0)
    out = np.empty(a.size, dtype=np.int8)
    _lib.check(lib.s2b_cellid_common_ancestor_level(_ptr(a), _ptr(b), a.size, _ptr(out)))
    return out


def info(ids):
    """S2CellId::is_valid / level / range_min / range_max element-wise -> (valid bool, level int8 (-1 if invalid), range_min, range_max)."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        n = ids.numel()
        v = torch.empty(n, dtype=torch.uint8, device=ids.device); l = torch.empty(n, dtype=torch.int8, device=ids.device)
        a = torch.empty(n, dtype=torch.int64, device=ids.device); b = torch.empty(n, dtype=torch.int64, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_info_dev(_ptr(ids), n, _ptr(v), _ptr(l), _ptr(a), _ptr(b), _stream()))
        return v != 0, l, a, b
    ids = _np_in(ids, np.uint64, 0)
    v = np.empty(ids.size, np.uint8); l = np.empty(ids.size, np.int8); a = np.empty(ids.size, np.uint64); b = np.empty(ids.size, np.uint64)
    _lib.check(lib.s2b_cellid_info(_ptr(ids), ids.size, _ptr(v), _ptr(l), _ptr(a), _ptr(b)))
    return v != 0, l, a, b


def vertex_neighbors(ids, level):
    """S2CellId::AppendVertexNeighbors(level) element-wise -> ((N,4) ids, (N,) counts): 3 or 4 cells of `level` around the
    closest vertex of parent(level), reference order, unused slots 0. Requires level < level(id)."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty((ids.numel(), 4), dtype=torch.int64, device=ids.device)
        cnt = torch.empty(ids.numel(), dtype=torch.uint8, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_vertex_neighbors_dev(_ptr(ids), ids.numel(), int(level), _ptr(out), _ptr(cnt), _stream()))
        return out, cnt
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty((ids.size, 4), dtype=np.uint64)
    cnt = np.empty(ids.size, dtype=np.uint8)
    _lib.check(lib.s2b_cellid_vertex_neighbors(_ptr(ids), ids.size, int(level), _ptr(out), _ptr(cnt)))
    return out, cnt


def all_neighbors(ids, nbr_level, stride=None):
    """S2CellId::AppendAllNeighbors(nbr_level) element-wise -> ((N,stride) ids, (N,) counts), reference order, unused slots
    0. stride defaults to 8 (enough when nbr_level equals every id's level); counts[i] = 4 * 2^(nbr_level - level) + 4."""
    lib = _lib.load()
    stride = 8 if stride is None else int(stride)
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty((ids.numel(), stride), dtype=torch.int64, device=ids.device)
        cnt = torch.empty(ids.numel(), dtype=torch.int32, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellid_all_neighbors_dev(_ptr(ids), ids.numel(), int(nbr_level), _ptr(out), stride, _ptr(cnt), _stream()))
        return out, cnt
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty((ids.size, stride), dtype=np.uint64)
    cnt = np.empty(ids.size, dtype=np.uint32)
    _lib.check(lib.s2b_cellid_all_neighbors(_ptr(ids), ids.size, int(nbr_level), _ptr(out), stride, _ptr(cnt)))
    return out, cnt


def cell_union_normalize(ids):
    """S2CellUnion::Normalize on the GPU: sorted, non-overlapping, sibling quads merged."""
    lib = _lib.load()
    import ctypes
    count = ctypes.c_size_t(0)
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        out = torch.empty(ids.numel(), dtype=torch.int64, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellunion_normalize_dev(_ptr(ids), ids.numel(), _ptr(out), ctypes.byref(count), _stream()))
        return out[: count.value]
    ids = _np_in(ids, np.uint64, 0)
    out = np.empty(ids.size, dtype=np.uint64)
    _lib.check(lib.s2b_cellunion_normalize(_ptr(ids), ids.size, _ptr(out), ctypes.byref(count)))
    return out[: count.value].copy()


def cell_union_contains_cells(sorted_ids, ids):
    """S2CellUnion::Contains(S2CellId) / Intersects(S2CellId) for a batch of ids -> (contains, intersects) bool arrays."""
    lib = _lib.load()
    if _is_tensor(ids):
        ids = _tensor_in(ids, torch.int64, 0)
        u = _tensor_in(sorted_ids, torch.int64, 0)
        out = torch.empty(ids.numel(), dtype=torch.uint8, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.s2b_cellunion_contains_cells_dev(_ptr(u), u.numel(), _ptr(ids), ids.numel(), _ptr(out), _stream()))
        return (out & 1) != 0, (out & 2) != 0
    ids = _np_in(ids, np.uint64, 0)
    u = _np_in(sorted_ids, np.uint64, 0)
    out = np.empty(ids.size, dtype=np.uint8)
    _lib.check(lib.s2b_cellunion_contains_cells(_ptr(u), u.size, _ptr(ids), ids.size, _ptr(out)))
    return (out & 1) != 0, (out & 2) != 0


def from_token(tokens, lengths):
    """S2CellId::FromToken element-wise: tokens (N,16) uint8 slots + lengths (N,) uint8 -> ids (0 for malformed tokens)."""
    lib = _lib.load()
    if _is_tensor(tokens):
        tokens = _tensor_in(tokens, torch.uint8, 16)
        lengths = _tensor_in(lengths, torch.uint8, 0)
        out = torch.empty(tokens.shape[0], dtype=torch.int64, device=tokens.device)
        with torch.cuda.device(tokens.device):
            _lib.check(lib.s2b_cellid_from_token_dev(_ptr(tokens), _ptr(lengths), tokens.shape[0], _ptr(out), _stream()))
        return out
    tokens = _np_in(tokens, np.uint8, 16)
    lengths = _np_in(lengths, np.uint8, 0)
    out = np.empty(tokens.shape[0], dtype=np.uint64)
    _lib.check(lib.s2b_cellid_from_token(_ptr(tokens), _ptr(lengths), tokens.shape[0], _ptr(out)))
    return out


def tokens_to_str(tokens, lengths):
    """Convenience: decode (N,16) uint8 token slots into Python strings."""
    tokens = tokens.cpu().numpy() if _is_tensor(tokens) else np.asarray(tokens)
    lengths = lengths.cpu().numpy() if _is_tensor(lengths) else np.asarray(lengths)
    return [bytes(tokens[i, : lengths[i]]).decode("ascii") for i in range(len(lengths))]
