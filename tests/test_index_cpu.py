"""Flat index + containment logic on the CPU (tests/hostsim build of s2b_contains.cuh) against the reference."""
import ctypes

import numpy as np
import pytest

from s2geometry_b200.index import FlatIndex
from tests import scenes
from tests.util import vp


def hs_query(hs, flat, xyz, model=1, mode=0, shape_id=0):
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = len(xyz)
    out8 = np.zeros(n, np.uint8); counts = np.zeros(max(flat.num_shape_ids, 1), np.uint64); out32 = np.zeros(n, np.int32)
    ne = ctypes.c_uint64(0)
    err = ctypes.create_string_buffer(256)
    st = flat.as_struct()
    rc = hs.hs_index_query(ctypes.byref(st), vp(xyz), ctypes.c_size_t(n), model, mode, shape_id, vp(out8), vp(counts), vp(out32),
                           ctypes.byref(ne), err, ctypes.c_size_t(256))
    assert rc == 0, err.value
    return {0: out8, 1: out8, 2: counts[: flat.num_shape_ids], 3: out32, 4: out32}[mode], ne.value


def flat_of(ref_index):
    return FlatIndex(**ref_index.flatten())


def test_vertex_model_truth_tables(hostsim, ref):
    ka = scenes.known_answers()
    soup = scenes.text_index(ka["index"])
    ri = ref.index_create(soup)
    flat = flat_of(ri)
    q = scenes.latlng_deg_to_point(ka["queries"])
    for name, m in scenes.MODELS.items():
        want = np.array(ka["contains"][name], np.uint8)
        assert (ri.contains(q, m, 1) == want).all(), name            # the oracle reproduces the reference's table
        assert (hs_query(hostsim, flat, q, m, 0)[0] == want).all(), name
    for row in ka["shape_contains"]:
        m = scenes.MODELS[row["model"]]
        p = scenes.latlng_deg_to_point([row["q"]])
        assert ri.shape_contains(row["shape"], p, m, 1)[0] == row["expected"]
        assert hs_query(hostsim, flat, p, m, 1, row["shape"])[0][0] == row["expected"]
    t = ka["closed_three_shapes"]
    ri2 = ref.index_create(scenes.text_index(t["index"]))
    p = scenes.latlng_deg_to_point([t["q"]])
    off, ids = ri2.containing_shapes(p, 2)
    assert list(ids) == t["containing_shape_ids_closed"]
    assert hs_query(hostsim, flat_of(ri2), p, 2, 4)[0][0] == 3


def test_cell_centers_match_reference(hostsim, ref):
    soup, v = scenes.loop_scene(2000)
    f = ref.index_create(soup).flatten()
    got = np.empty_like(f["cell_center"])
    hostsim.hs_cellid_to_point(vp(f["cell_ids"]), ctypes.c_size_t(len(f["cell_ids"])), vp(got))
    assert (got == f["cell_center"]).all()
    rng = np.random.default_rng(3)
    leaf = ref.encode_points(np.random.default_rng(4).standard_normal((20000, 3)), 30, 1)
    ids = np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(rng.integers(0, 31, len(leaf)))], np.uint64)
    got = np.empty((len(ids), 3))
    hostsim.hs_cellid_to_point(vp(ids), ctypes.c_size_t(len(ids)), vp(got))
    assert (got == ref.to_point(ids)).all()


def test_loop_10k_all_point_sets(hostsim, ref):
    soup, v = scenes.loop_scene(10000)
    ri = ref.index_create(soup)
    flat = flat_of(ri)
    assert flat.num_cells > 1000
    u, b, d = scenes.loop_points(v, 200000, 200000, 31)
    for pts in (u, b, d):
        for m in (0, 1, 2):
            got, _ = hs_query(hostsim, flat, pts, m, 0)
            assert (got == ri.contains(pts, m)).all()
    # D set also against the reference's ground truth (brute force, semi-open)
    assert (hs_query(hostsim, flat, d, 1, 0)[0] == ri.brute_force_contains(0, d)).all()
    # Locate: a point is in an index cell iff the reference index cell ids contain its leaf id
    loc, _ = hs_query(hostsim, flat, b, 1, 3)
    leaf = ref.encode_points(b, 30)
    lo, hi = ref.id_range(flat.cell_ids)
    j = np.searchsorted(lo, leaf, side="right") - 1
    hit = (j >= 0) & (leaf <= hi[np.maximum(j, 0)])
    assert (np.where(hit, j, -1) == loc).all()


def test_axis_aligned_exact_stage(hostsim, ref):
    soup = scenes.axis_aligned_scene()
    ri = ref.index_create(soup)
    flat = flat_of(ri)
    q = scenes.axis_aligned_queries()
    n_exact_total = 0
    for m in (0, 1, 2):
        got, ne = hs_query(hostsim, flat, q, m, 0)
        n_exact_total += ne
        assert (got == ri.contains(q, m, 1)).all(), m
        for s in range(flat.num_shape_ids):
            assert (hs_query(hostsim, flat, q, m, 1, s)[0] == ri.shape_contains(s, q, m, 1)).all(), (m, s)
        assert (hs_query(hostsim, flat, q, m, 2)[0] == ri.shape_histogram(q, m, 1)).all()
    assert n_exact_total > 100  # these queries really reach the exact stage


def test_many_polygons_histogram(hostsim, ref):
    soup = scenes.polygons_scene(300, 41, min_radius_km=200, max_radius_km=2000)
    ri = ref.index_create(soup)
    flat = flat_of(ri)
    from s2geometry_b200 import synth
    pts = synth.sphere_points(300000, 42)
    got, _ = hs_query(hostsim, flat, pts, 1, 2)
    want = ri.shape_histogram(pts, 1)
    assert want.sum() > 10000
    assert (got == want).all()
    off, ids = ri.containing_shapes(pts[:50000], 1)
    assert (hs_query(hostsim, flat, pts[:50000], 1, 4)[0] == np.diff(off.astype(np.int64))).all()


# ---------------------------------------------------------------------------------------------------
# The library's own host builder (s2b_index_build): query results over ITS index must equal the reference's.

def own_flat(soup, max_edges=0):
    from s2geometry_b200.index import build_flat_index
    return build_flat_index(soup, max_edges)


def check_index_invariants(flat, soup, ref):
    """Structure checks in the spirit of the reference's QuadraticValidate/ValidateInterior
    (mutable_s2shape_index_test.cc:127-229): sorted disjoint cells, sorted clips, and contains_center equal to the
    reference's brute-force containment of the cell centre."""
    lo, hi = ref.id_range(flat.cell_ids)
    assert (lo[1:] > hi[:-1]).all()
    assert (ref.to_point(flat.cell_ids) == flat.cell_center).all()
    ri = ref.index_create(soup)
    cell_of_clip = np.repeat(np.arange(flat.num_cells), np.diff(flat.cell_clip_begin.astype(np.int64)))
    for c in range(flat.num_cells):
        ids = flat.clip_shape_id[flat.cell_clip_begin[c]:flat.cell_clip_begin[c + 1]]
        assert (np.diff(ids) > 0).all()
    dims = soup.shape_dim
    for s in range(flat.num_shape_ids):
        if dims[s] != 2:
            continue
        want = ri.brute_force_contains(s, flat.cell_center, 1)
        got = np.zeros(flat.num_cells, np.uint8)
        sel = flat.clip_shape_id == s
        got[cell_of_clip[sel]] = flat.clip_flags[sel] & 1
        assert (got == want).all(), s


def test_own_builder_truth_tables(hostsim, ref):
    ka = scenes.known_answers()
    soup = scenes.text_index(ka["index"])
    flat = own_flat(soup)
    q = scenes.latlng_deg_to_point(ka["queries"])
    for name, m in scenes.MODELS.items():
        assert (hs_query(hostsim, flat, q, m, 0)[0] == np.array(ka["contains"][name], np.uint8)).all(), name
    for row in ka["shape_contains"]:
        p = scenes.latlng_deg_to_point([row["q"]])
        assert hs_query(hostsim, flat, p, scenes.MODELS[row["model"]], 1, row["shape"])[0][0] == row["expected"]
    check_index_invariants(flat, soup, ref)


def test_own_builder_loop_10k(hostsim, ref):
    soup, v = scenes.loop_scene(10000)
    flat = own_flat(soup)
    ri = ref.index_create(soup)
    rflat = flat_of(ri)
    # comparable size to the reference's index (same subdivision rule, conservative edge lists)
    assert 0.8 * rflat.num_cells <= flat.num_cells <= 1.3 * rflat.num_cells
    assert flat.num_edges <= 1.3 * rflat.num_edges
    check_index_invariants(flat, soup, ref)
    u, b, d = scenes.loop_points(v, 200000, 200000, 33)
    centers = flat.cell_center
    for pts in (u, b, d, centers):
        for m in (0, 1, 2):
            assert (hs_query(hostsim, flat, pts, m, 0)[0] == ri.contains(pts, m)).all()


def test_own_builder_axis_aligned_and_polygons(hostsim, ref):
    soup = scenes.axis_aligned_scene()
    flat = own_flat(soup)
    check_index_invariants(flat, soup, ref)
    ri = ref.index_create(soup)
    q = scenes.axis_aligned_queries()
    for m in (0, 1, 2):
        assert (hs_query(hostsim, flat, q, m, 0)[0] == ri.contains(q, m, 1)).all()
        assert (hs_query(hostsim, flat, q, m, 2)[0] == ri.shape_histogram(q, m, 1)).all()
    # many polygons, some with holes, some overlapping, tiny max_edges_per_cell to force deep subdivision
    from s2geometry_b200 import synth
    soup2 = scenes.polygons_scene(300, 41, min_radius_km=200, max_radius_km=2000)
    ri2 = ref.index_create(soup2)
    pts = synth.sphere_points(300000, 43)
    for max_edges in (0, 3):
        flat2 = own_flat(soup2, max_edges)
        assert (hs_query(hostsim, flat2, pts, 1, 2)[0] == ri2.shape_histogram(pts, 1)).all()
    check_index_invariants(own_flat(soup2), soup2, ref)


def test_own_builder_whole_face_and_full_polygon(hostsim, ref):
    """A polygon larger than a hemisphere (contains whole cube faces) and its complement."""
    from s2geometry_b200 import synth
    from s2geometry_b200.shapes import ShapeSoup
    small = synth.wobbly_loop([0.1, 0.2, 1.0], 0.3, 40)
    soup = ShapeSoup()
    soup.add_polygon([small[::-1].copy()])   # clockwise: everything EXCEPT the small cap
    soup.add_polygon([small])
    soup = soup.freeze()
    flat = own_flat(soup)
    check_index_invariants(flat, soup, ref)
    ri = ref.index_create(soup)
    pts = np.concatenate([synth.sphere_points(200000, 44), synth.cap_points([0.1, 0.2, 1.0], 0.4, 100000, 45), small])
    assert (hs_query(hostsim, flat, pts, 1, 2)[0] == ri.shape_histogram(pts, 1)).all()
    got, _ = hs_query(hostsim, flat, pts, 1, 4)
    assert (got[: 300000] == 1).all()   # away from the boundary every point is in exactly one of the two


def test_cells_with_hundreds_of_edges(hostsim, ref):
    """max_edges_per_cell = 1000 puts hundreds of edges in one clip: exercises the 64-edge chunks of the pending mask."""
    soup, v = scenes.loop_scene(3000)
    flat = own_flat(soup, 1000)
    assert np.diff(flat.clip_edge_begin.astype(np.int64)).max() > 200
    ri = ref.index_create(soup)
    u, b, d = scenes.loop_points(v, 20000, 100000, 35)
    for pts in (b, d):
        for m in (0, 1, 2):
            assert (hs_query(hostsim, flat, pts, m, 0)[0] == ri.contains(pts, m)).all()


def test_degenerate_shapes(hostsim, ref):
    """Degenerate geometry the lax shapes allow: repeated vertices (zero-length edges), two-vertex loops (an edge and its
    reverse), a loop given twice, the full polygon (one empty chain), shapes without chains, a point shape without points."""
    from s2geometry_b200 import synth
    from s2geometry_b200.shapes import ShapeSoup
    c = np.array([0.2, 0.3, 0.9])
    loop = synth.wobbly_loop(c, 0.2, 12)
    soup = ShapeSoup()
    soup.add_polygon([np.concatenate([loop[:5], loop[4:5], loop[5:]])])          # repeated vertex
    soup.add_polygon([loop[:2]])                                                  # sibling pair: zero area
    soup.add_polygon([loop, loop])                                                # same loop twice
    soup.add_polygon([np.zeros((0, 3))])                                          # full polygon
    soup.add_polygon([])                                                          # empty polygon (no chains)
    soup.add_points(np.zeros((0, 3)))
    soup.add_polygon([loop, loop[::-1].copy()])                                   # a loop and its reverse: every edge matched
    soup.add_polyline(loop[:1])                                                   # one-vertex polyline: no edges
    soup = soup.freeze()
    ri = ref.index_create(soup)
    flat = own_flat(soup)
    pts = np.concatenate([synth.cap_points(c, 0.3, 4000, 5), synth.sphere_points(4000, 6), synth.degenerate_points(loop)])
    for m in (0, 1, 2):
        assert (hs_query(hostsim, flat, pts, m, 2)[0] == ri.shape_histogram(pts, m, 1)).all(), m
        assert (hs_query(hostsim, flat, pts, m, 0)[0] == ri.contains(pts, m, 1)).all(), m
    rflat = flat_of(ri)
    for m in (0, 1, 2):
        assert (hs_query(hostsim, rflat, pts, m, 2)[0] == ri.shape_histogram(pts, m, 1)).all(), m


def _cell_face_ij(cell_id):
    """(face, i_lo, j_lo, size in leaf cells) of one S2CellId (ToFaceIJOrientation, s2cell_id.cc:319-355)."""
    from oracle import restatement as rs
    cid = int(cell_id)
    face = cid >> 61
    bits = face & 1
    i = j = 0
    for k in range(7, -1, -1):
        nbits = 2 if k == 7 else 4
        bits += ((cid >> (k * 8 + 1)) & ((1 << (2 * nbits)) - 1)) << 2
        bits = int(rs.LOOKUP_IJ[bits])
        i += (bits >> 6) << (k * 4)
        j += ((bits >> 2) & 15) << (k * 4)
        bits &= 3
    lsb = cid & -cid
    size = 1 << ((lsb.bit_length() - 1) // 2)      # leaf cells per side
    return face, i & -size, j & -size, size


def _face_frame(face, p):
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    return [(y, z, x), (-x, z, y), (-x, -y, z), (-z, -y, -x), (-z, x, -y), (y, x, -z)][face]


def _chord_meets_rect_exactly(a, b, face, rect):
    """Exact rational decision: does the chord a + t (b - a), t in [0, 1], centrally projected on `face`, meet the closed
    rectangle [ulo, uhi] x [vlo, vhi] (the cell's bound as the doubles every query sees)? The five conditions
    w >= 0, pu - ulo w >= 0, uhi w - pu >= 0, pv - vlo w >= 0, vhi w - pv >= 0 are linear in t."""
    from fractions import Fraction as F
    au, av, aw = (F(float(c)) for c in _face_frame(face, a))
    bu, bv, bw = (F(float(c)) for c in _face_frame(face, b))
    ulo, uhi, vlo, vhi = (F(float(c)) for c in rect)
    t0, t1 = F(0), F(1)
    for f0, f1 in ((aw, bw), (au - ulo * aw, bu - ulo * bw), (uhi * aw - au, uhi * bw - bu), (av - vlo * aw, bv - vlo * bw),
                   (vhi * aw - av, vhi * bw - bv)):
        if f0 >= 0 and f1 >= 0:
            continue
        if f0 < 0 and f1 < 0:
            return False
        tz = f0 / (f0 - f1)
        if f0 < 0:
            t0 = max(t0, tz)
        else:
            t1 = min(t1, tz)
        if t0 > t1:
            return False
    return True


@pytest.mark.parametrize("scene", ["loop", "polygons_fine"])
def test_own_builder_edge_lists_are_supersets_of_the_exact_intersections(scene):
    """The builder's clipping (s2b_index_builder.h: linear inequalities with the slacks derived in its header comment) must
    never drop an edge that meets a cell: for sampled cells, every edge whose chord meets the cell's rectangle in EXACT
    rational arithmetic has to be in that cell's clipped list. (Point queries need exactly this: the segment from the cell
    centre to a query point of the cell stays inside the rectangle, so every edge it can cross is in the list.)"""
    from oracle import restatement as rs
    from s2geometry_b200 import synth
    if scene == "loop":
        soup, _ = scenes.loop_scene(3000)
        flat = own_flat(soup)
    else:
        soup = scenes.polygons_scene(120, 77, min_radius_km=50, max_radius_km=1500)
        flat = own_flat(soup, 1)
    # every edge of the scene, from the soup's chains (polygon chains are closed)
    e0, e1 = [], []
    cvb, scb = soup.chain_vertex_begin, soup.shape_chain_begin
    for s in range(soup.num_shapes):
        for c in range(scb[s], scb[s + 1]):
            v = soup.vertices[cvb[c]:cvb[c + 1]]
            if soup.shape_dim[s] == 2 and len(v):
                e0.append(v); e1.append(np.roll(v, -1, 0))
            elif soup.shape_dim[s] == 1 and len(v) > 1:
                e0.append(v[:-1]); e1.append(v[1:])
    e0 = np.concatenate(e0); e1 = np.concatenate(e1)
    rng = np.random.default_rng(5)
    clip_begin = flat.cell_clip_begin.astype(np.int64)
    edge_begin = flat.clip_edge_begin.astype(np.int64)
    checked = exact_hits = extra = 0
    cell_edges = edge_begin[clip_begin[1:]] - edge_begin[clip_begin[:-1]]
    sample = np.concatenate([rng.choice(np.flatnonzero(cell_edges > 0), 200, replace=False), rng.choice(flat.num_cells, 50, replace=False)])
    for c in sample:
        face, i, j, size = _cell_face_ij(flat.cell_ids[c])
        rect = [float(rs.st_to_uv(np.float64((1.0 / 1073741824.0) * x))) for x in (i, i + size, j, j + size)]
        rect = [rect[0], rect[1], rect[2], rect[3]]
        lo, hi = edge_begin[clip_begin[c]], edge_begin[clip_begin[c + 1]]
        have = {flat.edge_v0[k].tobytes() + flat.edge_v1[k].tobytes() for k in range(lo, hi)}
        # float screening with a generous margin picks the candidates for the exact test
        au, av, aw = _face_frame(face, e0); bu, bv, bw = _face_frame(face, e1)
        t0 = np.zeros(len(e0)); t1 = np.ones(len(e0)); alive = np.ones(len(e0), bool)
        with np.errstate(all="ignore"):
            for f0, f1 in ((aw, bw), (au - rect[0] * aw, bu - rect[0] * bw), (rect[1] * aw - au, rect[1] * bw - bu),
                           (av - rect[2] * aw, bv - rect[2] * bw), (rect[3] * aw - av, rect[3] * bw - bv)):
                f0 = f0 + 1e-9; f1 = f1 + 1e-9
                alive &= ~((f0 < 0) & (f1 < 0))
                tz = f0 / (f0 - f1)
                t0 = np.where((f0 < 0) & (f1 >= 0), np.maximum(t0, tz - 1e-9), t0)
                t1 = np.where((f0 >= 0) & (f1 < 0), np.minimum(t1, tz + 1e-9), t1)
            alive &= t0 <= t1
        for k in np.flatnonzero(alive):
            checked += 1
            if _chord_meets_rect_exactly(e0[k], e1[k], face, rect):
                exact_hits += 1
                assert e0[k].tobytes() + e1[k].tobytes() in have, (int(c), int(k))
        extra += len(have)
    assert exact_hits > 200 and checked >= exact_hits
    assert extra <= 1.5 * exact_hits + 50      # and the lists are not much larger than the exact ones
