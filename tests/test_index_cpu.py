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
