"""S2ShapeIndexRegion cell tests (s2b_region.cuh compiled for the host by tests/hostsim) against the reference's
S2ShapeIndexRegion, S2::ClipToPaddedFace and S2::IntersectsRect."""
import ctypes

import numpy as np

from s2geometry_b200 import synth
from s2geometry_b200.index import FlatIndex
from tests import region_scenes, scenes
from tests.util import vp

KMAX = 9 * np.sqrt(0.5) * np.finfo(float).eps + 3 * np.sqrt(2.0) * np.finfo(float).eps


def hs_region_cells(hs, flat, ids, mode):
    out = np.zeros(len(ids), np.uint8)
    bad = ctypes.c_uint64(0)
    st = flat.as_struct()
    assert hs.hs_region_cells(ctypes.byref(st), vp(ids), ctypes.c_size_t(len(ids)), mode, vp(out), ctypes.byref(bad)) == 0
    return out, bad.value


def hs_region_shapes(hs, flat, ids):
    hs.hs_region_intersecting_shapes.restype = ctypes.c_size_t
    st = flat.as_struct()
    off = np.zeros(len(ids) + 1, np.uint32)
    total = hs.hs_region_intersecting_shapes(ctypes.byref(st), vp(ids), ctypes.c_size_t(len(ids)), vp(off), None, None, ctypes.c_size_t(0))
    shp = np.zeros(max(total, 1), np.int32); con = np.zeros(max(total, 1), np.uint8)
    hs.hs_region_intersecting_shapes(ctypes.byref(st), vp(ids), ctypes.c_size_t(len(ids)), vp(off), vp(shp), vp(con), ctypes.c_size_t(total))
    return off, shp[:total], con[:total]


def test_clip_to_padded_face_and_intersects_rect_bit_exact(hostsim, ref):
    rng = np.random.default_rng(21)
    n = 60000
    a = synth.sphere_points(n, 31)
    # edges of every length: long random ones, short ones, some exactly on face boundaries / cube diagonals
    step = rng.choice([1.5, 0.3, 1e-2, 1e-5, 1e-9], n)[:, None]
    b = scenes._norm(a + step * rng.standard_normal((n, 3)))
    special = scenes._norm(np.array([[1, 1, 0], [1, 1, 1], [1, 0, 0], [0, 1, 1], [1, -1, 1], [1, 1e-300, 0], [-1, -1, -1]], float))
    a[: len(special)] = special
    b[len(special): 2 * len(special)] = special
    a[-50:] = b[-50:]                       # degenerate edges (RobustCrossProd returns Ortho)
    face = rng.integers(0, 6, n).astype(np.int32)
    for padding in (0.0, KMAX, 1e-3):
        uv = np.zeros((n, 4)); ok = np.zeros(n, np.uint8); sup = np.zeros(n, np.uint8)
        hostsim.hs_clip_to_padded_face(vp(a), vp(b), vp(face), ctypes.c_size_t(n), ctypes.c_double(padding), vp(uv), vp(ok), vp(sup))
        want_uv, want_ok = ref.clip_to_padded_face(a, b, face, padding)
        s = sup.astype(bool)
        assert s.mean() > 0.99
        assert (ok[s] == want_ok[s]).all()
        hit = s & (ok == 1)
        assert hit.sum() > 5000
        assert (uv[hit].view(np.uint64) == want_uv[hit].view(np.uint64)).all()      # bit for bit
        # IntersectsRect on the clipped edges against cell-sized and face-sized rectangles
        c = rng.uniform(-1, 1, (n, 2)); h = rng.choice([1e-6, 1e-3, 0.1, 1.0], n)
        rect = np.stack([c[:, 0] - h, c[:, 0] + h, c[:, 1] - h, c[:, 1] + h], 1)
        rect[::7, 0] = uv[::7, 0]            # rectangle sides through an endpoint
        rect[::11, 3] = uv[::11, 3]
        got = np.zeros(n, np.uint8)
        hostsim.hs_intersects_rect(vp(want_uv), vp(np.ascontiguousarray(rect)), ctypes.c_size_t(n), vp(got))
        want = ref.intersects_rect(want_uv, rect)
        assert (got == want).all()
        assert 0.02 < want.mean() < 0.98


def test_region_cell_tests_match_reference(hostsim, ref):
    for name, soup in region_scenes.region_scene_list():
        ri = ref.index_create(soup)
        flat = FlatIndex(**ri.flatten())
        targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 5)
        rel, _ = ri.locate_cells(targets)
        assert all((rel == k).sum() > 50 for k in (0, 1, 2)), name
        for mode in (0, 1):
            got, bad = hs_region_cells(hostsim, flat, targets, mode)
            assert bad == 0, name
            want = ri.region_cells(targets, mode)
            assert (got == want).all(), (name, mode, int((got != want).sum()))
            inside = rel == 0
            assert 0 < want[inside].sum() < inside.sum(), (name, mode)      # both answers occur below the index cells
        off, shp, con = hs_region_shapes(hostsim, flat, targets)
        woff, wshp, wcon = ri.region_intersecting_shapes(targets)
        assert (off == woff).all() and (shp == wshp).all() and (con == wcon).all(), name
        assert 0 < wcon.sum() < len(wcon), name


def test_unsupported_edges_are_detected(hostsim, ref):
    """Edges whose RobustCrossProd needs the extended-precision stages are counted (the C ABI refuses such an index)."""
    a = scenes._norm(np.array([[1.0, 1e-3, 0.0]]))
    b = a.copy(); b[0, 1] += 2e-16
    uv = np.zeros((1, 4)); ok = np.zeros(1, np.uint8); sup = np.ones(1, np.uint8)
    hostsim.hs_clip_to_padded_face(vp(a), vp(b), vp(np.array([0], np.int32)), ctypes.c_size_t(1), ctypes.c_double(0.0), vp(uv), vp(ok), vp(sup))
    assert sup[0] == 0


def test_region_cell_tests_over_the_librarys_own_builder(hostsim, ref):
    """The library's host builder produces the reference's cells at the same granularity, so the region answers over
    its index equal the reference's too (they depend on the cell structure: SUBDIVIDED targets "may intersect")."""
    from s2geometry_b200.index import build_flat_index
    for name, soup in region_scenes.region_scene_list():
        ri = ref.index_create(soup)
        flat = build_flat_index(soup, 0)
        if len(flat.cell_ids) != len(ri.flatten()["cell_ids"]) or (flat.cell_ids != ri.flatten()["cell_ids"]).any():
            continue        # identical cells are a property of these scenes, not a guarantee of the builder
        targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 9, 500)
        for mode in (0, 1):
            got, bad = hs_region_cells(hostsim, flat, targets, mode)
            assert bad == 0 and (got == ri.region_cells(targets, mode)).all(), (name, mode)


def test_cell_union_bound_matches_reference(hostsim, ref):
    """GetCellUnionBound: single-face indexes (children of the smallest enclosing cell), multi-face ones, one cell, one
    cell per face."""
    from s2geometry_b200.shapes import ShapeSoup
    from s2geometry_b200 import synth
    cases = [soup for _, soup in region_scenes.region_scene_list()]
    for centre, radius, nv in (([1, 0.01, 0.02], 1e-3, 12), ([1, 1, 1], 0.3, 30), ([0.2, -0.9, 0.1], 1e-6, 5), ([0, 0, 1], 1.4, 40)):
        s = ShapeSoup(); s.add_polygon([synth.wobbly_loop(np.array(centre, float), radius, nv)]); cases.append(s.freeze())
    s = ShapeSoup(); s.add_points(scenes._norm(np.array([[1, 0.3, 0.2]], float))); cases.append(s.freeze())
    s = ShapeSoup(); s.add_points(scenes._norm(np.array([[1, .1, .1], [-1, .1, .1], [.1, 1, .1], [.1, -1, .1], [.1, .1, 1], [.1, .1, -1]], float))); cases.append(s.freeze())
    seen = set()
    for soup in cases:
        ri = ref.index_create(soup)
        ids = np.ascontiguousarray(ri.flatten()["cell_ids"], np.uint64)
        out = np.zeros(6, np.uint64)
        n = hostsim.hs_region_cell_union_bound(vp(ids), ctypes.c_size_t(len(ids)), vp(out))
        want = ri.region_cell_union_bound()
        assert n == len(want) and (out[:n] == want).all()
        seen.add(n)
    assert {1, 4, 6} & seen and len(seen) >= 3
