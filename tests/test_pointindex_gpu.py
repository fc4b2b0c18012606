"""GPU parity tests for the batched S2ClosestPointQuery over an S2PointIndex (s2b_pointindex.cu) against the reference's
own S2ClosestPointQuery<int> / S2PointIndex<int> (oracle/_ref): radius counts (doc/examples/point_index.cc) and
FindClosestPoint, with exact S1ChordAngle distances."""
import math

import numpy as np
import pytest

from s2geometry_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def scene(seed, m=200_000, n=40_000):
    """Index: uniform points + a dense cluster + exact duplicates. Targets: uniform, near the cluster, ON index points."""
    rng = np.random.default_rng(seed)
    c = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    pts = np.concatenate([synth.sphere_points(m, seed), synth.cap_points(c, 0.01, m // 4, seed + 1)])
    pts = np.concatenate([pts, pts[:1000]])                                  # duplicates (ties in distance)
    targets = np.concatenate([synth.sphere_points(n, seed + 2), synth.cap_points(c, 0.03, n // 4, seed + 3),
                              pts[rng.integers(0, len(pts), 2000)],           # distance exactly 0
                              np.array([[1.0, 0, 0], [0, 0, -1.0], [-1.0, 0, 0]]) / 1.0])
    return pts, targets


def test_chord_length2_matches_reference(ref):
    from s2geometry_b200.pointindex import chord_length2
    for a in (0.0, 1e-9, 1e-3, 0.02, 0.3, 1.0, 2.0, math.pi, 4.0):
        assert chord_length2(a) == ref.chord_length2(a)


def test_radius_counts_match_reference(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    pts, targets = scene(1)
    ri = ref.point_index(pts)
    ix = PointIndex(pts)
    assert ix.num_points == len(pts)
    total = 0
    # 1 km and 100 km on the Earth, a few degrees, > 90 degrees (whole-sphere covering), and the extremes
    # (the reference materialises every result, so large radii are checked on fewer targets)
    for r, k in ((0.0, None), (1.57e-4, None), (1.57e-2, 20000), (0.05, 1500), (0.3, 150), (1.2, 12), (2.0, 8), (math.pi, 8)):
        sel = targets if k is None else np.ascontiguousarray(targets[:: len(targets) // k])
        want = ri.count(sel, r)
        q = ClosestPointQuery(ix, r)
        got = q.count(sel)
        assert (got == want).all(), (r, int((got != want).sum()))
        assert (q.count(torch.from_numpy(sel).cuda()).cpu().numpy().astype(np.uint32) == want).all(), r
        total += int(want.sum())
    assert total > 2 * len(targets)
    assert (ClosestPointQuery(ix, math.pi).count(targets[:100]) <= len(pts)).all()
    # no distance limit: every index point qualifies
    assert (ClosestPointQuery(ix).count(targets[:50]) == len(pts)).all()
    assert (ri.count(targets[:2]) == len(pts)).all()


def test_closest_point_matches_reference(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    pts, targets = scene(2, m=100_000, n=30_000)
    ri = ref.point_index(pts)
    ix = PointIndex(pts)
    for r in (None, 1.0, 0.01, 1e-4):
        wd, wl = ri.closest(targets, -1.0 if r is None else r)
        q = ClosestPointQuery(ix) if r is None else ClosestPointQuery(ix, r)
        gd, gl = q.closest(targets)
        assert (gl == wl).all(), r                                  # the exact S1ChordAngle of the nearest point
        assert ((gd < 0) == (wd < 0)).all()
        ok = gd >= 0
        # the returned point really is at that distance (equidistant points — the duplicates — may differ in data)
        d = pts[gd[ok]] - targets[ok]
        l2 = np.minimum(4.0, ((0.0 + d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        assert (l2 == gl[ok]).all()
        assert (gd[ok] == wd[ok]).mean() > 0.95
        dd, dl = q.closest(torch.from_numpy(targets).cuda())
        assert (dd.cpu().numpy() == gd).all() and (dl.cpu().numpy() == gl).all()
        if r is None:
            assert ok.all()
        if r == 1e-4:
            assert (~ok).any() and ok.any()
    # a tiny index far from most targets: the search has to widen to the whole sphere
    few = synth.cap_points(np.array([0.0, 0.0, 1.0]), 0.001, 7, 5)
    rf = ref.point_index(few)
    wd, wl = rf.closest(targets[:5000])
    gd, gl = ClosestPointQuery(PointIndex(few)).closest(targets[:5000])
    assert (gl == wl).all() and (gd == wd).all()


def test_k_closest_points_match_reference(torch_cuda, ref):
    """FindClosestPoints with max_results = k: the sorted distance lists equal the reference's; so do the data values
    wherever the distances are distinct (the index holds duplicates, i.e. genuine ties)."""
    torch = torch_cuda
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    pts, targets = scene(3, m=100_000, n=8_000)
    ri = ref.point_index(pts)
    ix = PointIndex(pts)
    for k, r in ((1, None), (5, None), (32, None), (8, 0.02), (3, 1e-4), (16, 2.0)):
        wd, wl, wc = ri.closest_k(targets, k, -1.0 if r is None else r)
        q = ClosestPointQuery(ix) if r is None else ClosestPointQuery(ix, r)
        gd, gl, gc = q.closest_k(targets, k)
        assert (gc == wc).all(), (k, r)
        assert (gl == wl).all(), (k, r)                               # exact S1ChordAngle lists, nearest first
        assert (np.diff(np.where(np.isinf(gl), 4.0, gl), axis=1) >= 0).all()
        distinct = np.ones_like(gl, bool)
        distinct[:, 1:] &= gl[:, 1:] != gl[:, :-1]
        distinct[:, :-1] &= gl[:, :-1] != gl[:, 1:]
        last = np.take_along_axis(gl, np.maximum(gc.astype(np.int64) - 1, 0)[:, None], axis=1)
        distinct &= gl != last                                       # the k-th place may be tied with a point left out
        assert (gd[distinct] == wd[distinct]).all(), (k, r)
        assert ((gd >= 0).sum(axis=1) == gc).all()
        dd, dl, dc = q.closest_k(torch.from_numpy(targets).cuda(), k)
        assert (dd.cpu().numpy() == gd).all() and (dl.cpu().numpy() == gl).all() and (dc.cpu().numpy().astype(np.uint32) == gc).all()
    # k = 1 agrees with closest()
    d1, l1 = ClosestPointQuery(ix).closest(targets)
    dk, lk, _ = ClosestPointQuery(ix).closest_k(targets, 1)
    assert (l1 == lk[:, 0]).all() and (d1 == dk[:, 0]).all()
    from s2geometry_b200 import _lib
    with pytest.raises(_lib.S2BError):
        ClosestPointQuery(ix).closest_k(targets[:4], 33)


def test_large_batches_take_the_sorted_order_path(torch_cuda, ref):
    """Batches of >= 2^18 targets are processed in cell-sorted order internally; the results stay in the caller's order."""
    torch = torch_cuda
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    pts, _ = scene(4, m=150_000, n=10)
    targets = np.concatenate([synth.sphere_points(300_000, 91), pts[::3][:40_000]])
    assert len(targets) >= 1 << 18
    ri = ref.point_index(pts)
    ix = PointIndex(pts)
    td = torch.from_numpy(targets).cuda()
    for r in (3e-3, 0.03):
        want = ri.count(targets, r)
        q = ClosestPointQuery(ix, r)
        assert (q.count(targets) == want).all() and (q.count(td).cpu().numpy().astype(np.uint32) == want).all(), r
    wd, wl = ri.closest(targets)
    gd, gl = ClosestPointQuery(ix).closest(td)
    assert (gl.cpu().numpy() == wl).all() and ((gd.cpu().numpy() == wd).mean() > 0.95)
    kd, kl, kc = ri.closest_k(targets, 4, 0.01)
    gd4, gl4, gc4 = ClosestPointQuery(ix, 0.01).closest_k(targets, 4)
    assert (gl4 == kl).all() and (gc4 == kc).all()


def test_empty_index_and_empty_targets(torch_cuda):
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    ix = PointIndex(np.zeros((0, 3)))
    t = synth.sphere_points(100, 3)
    assert (ClosestPointQuery(ix, 0.5).count(t) == 0).all()
    d, l = ClosestPointQuery(ix).closest(t)
    assert (d == -1).all() and np.isinf(l).all()
    ix2 = PointIndex(t)
    assert ClosestPointQuery(ix2, 0.5).count(np.zeros((0, 3))).size == 0


def target_scene(ref, seed, m=120_000):
    """An index with a dense cluster + duplicates, and edge / cell targets of every size around and far from it: short
    and long edges (some touching index points, some degenerate a == b, a few nearly antipodal), cells of levels 0..30
    (containing index points, empty, far away)."""
    rng = np.random.default_rng(seed)
    c = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    pts = np.concatenate([synth.sphere_points(m, seed), synth.cap_points(c, 0.02, m // 3, seed + 1)])
    pts = np.concatenate([pts, pts[:500]])
    n = 6000
    a = np.concatenate([synth.sphere_points(n // 2, seed + 2), synth.cap_points(c, 0.05, n // 2, seed + 3)])
    step = 10.0 ** rng.uniform(-7, 0.3, (n, 1)) * rng.standard_normal((n, 3))
    b = a + step
    b /= np.linalg.norm(b, axis=1)[:, None]
    a[::50] = pts[rng.integers(0, len(pts), len(a[::50]))]            # an endpoint ON an index point
    b[::77] = a[::77]                                                  # degenerate edges
    b[5::400] = -a[5::400] + 1e-9 * rng.standard_normal(a[5::400].shape); b /= np.linalg.norm(b, axis=1)[:, None]   # nearly antipodal
    edges = np.ascontiguousarray(np.concatenate([a, b], axis=1))
    leaf = ref.encode_points(np.concatenate([pts[rng.integers(0, len(pts), 3000)], synth.sphere_points(3000, seed + 5)]), 30)
    lv = rng.integers(0, 31, len(leaf))
    cells = np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(lv)], np.uint64)
    return pts, edges, cells


def test_edge_and_cell_targets_match_reference(torch_cuda, ref):
    """S2ClosestPointQuery with EdgeTarget and CellTarget (and PointTarget through the same entry point): counts within
    max_distance, the k nearest (distances identical; data identical wherever the distances are distinct), no limit,
    through the host-pointer and the device-pointer calls."""
    torch = torch_cuda
    from s2geometry_b200.pointindex import TARGET_CELL, TARGET_EDGE, TARGET_POINT, ClosestPointQuery, PointIndex
    pts, edges, cells = target_scene(ref, 3)
    ri = ref.point_index(pts)
    ix = PointIndex(pts)
    th = ref.hardware_threads()
    sets = ((TARGET_EDGE, edges), (TARGET_CELL, cells), (TARGET_POINT, np.ascontiguousarray(edges[:2000, :3])))
    checked = 0
    for kind, targets in sets:
        for r in (1e-4, 5e-3, 0.06):
            q = ClosestPointQuery(ix, r)
            sub = targets if r < 0.05 else targets[::8]
            _, _, want = ri.closest_targets(kind, sub, 0, r, 0.0, th)
            got = q.find(kind, sub, 0)
            assert (got == want).all(), (kind, r, int((got != want).sum()))
            dev = torch.from_numpy(sub.view(np.int64) if kind == TARGET_CELL else sub).cuda()
            assert (q.find(kind, dev, 0).cpu().numpy().astype(np.uint32) == want).all(), (kind, r)
            checked += int(want.sum())
        for k, r in ((1, math.inf), (5, math.inf), (3, 0.01), (32, 0.03)):
            q = ClosestPointQuery(ix, r)
            sub = targets[::3]
            wd, wl, wc = ri.closest_targets(kind, sub, k, -1.0 if math.isinf(r) else r, 0.0, th)
            gd, gl, gc = q.find(kind, sub, k)
            assert (gc == wc).all(), (kind, k, r)
            assert (gl == wl).all(), (kind, k, r, int((gl != wl).sum()))
            # the reference orders equidistant results arbitrarily (the index holds 500 duplicated points, and a cell target
            # is at distance 0 from every point inside it): the data must agree wherever a row's distances are distinct
            # and none of its points is one of the duplicates
            dup = (wd < 500) | (wd >= len(pts) - 500)
            distinct = np.array([len(set(row[:c])) == c for row, c in zip(wl, wc)]) & ~dup.any(axis=1) & (wl[:, 0] > 0)
            assert (gd[distinct] == wd[distinct]).all(), (kind, k, r)
            assert distinct.sum() >= 50, (kind, k, r, int(distinct.sum()))   # the comparison is not vacuous
            dd, dl, dc = q.find(kind, torch.from_numpy(sub.view(np.int64) if kind == TARGET_CELL else sub).cuda(), k)
            assert (dl.cpu().numpy() == wl).all() and (dc.cpu().numpy().astype(np.uint32) == wc).all()
    assert checked > 100_000
    # max_error: the reference may return points up to max_error further away; the exact answer is never further than that
    q = ClosestPointQuery(ix, math.inf, max_error_rad=0.01)
    sub = edges[::10]
    _, wl, _ = ri.closest_targets(TARGET_EDGE, sub, 1, -1.0, 0.01, th)
    _, gl, _ = q.find(TARGET_EDGE, sub, 1)
    exact = ri.closest_targets(TARGET_EDGE, sub, 1, -1.0, 0.0, th)[1]
    assert (gl == exact).all() and (np.sqrt(gl) <= np.sqrt(wl) + 1e-12).all()


def _same_results(got, want, truncated):
    """(data, length2) lists: the distance sequences must be identical; the data must agree group by group of equal
    distance (the reference orders ties arbitrarily), except for a last group cut by max_results."""
    gd, gl = got
    wd, wl = want
    assert len(gd) == len(wd), (len(gd), len(wd))
    assert (gl.view(np.uint64) == wl.view(np.uint64)).all(), int((gl != wl).sum())
    if len(gl) == 0:
        return
    starts = np.flatnonzero(np.concatenate([[True], gl[1:] != gl[:-1]]))
    ends = np.concatenate([starts[1:], [len(gl)]])
    for a, b in zip(starts, ends):
        if truncated and b == len(gl):
            continue
        assert sorted(gd[a:b].tolist()) == sorted(wd[a:b].tolist()), (a, b)


def test_shape_index_target_matches_reference(torch_cuda, ref):
    """S2ClosestPointQuery with a ShapeIndexTarget: distances to a whole shape index (polygons with holes, a polyline, a
    point shape), with and without polygon interiors, with a distance limit, a result limit (<= 32 and > 32), both, and
    neither (small index) — identical distances to the reference's, index built at the default and the device granularity."""
    from s2geometry_b200.index import ShapeIndex, build_flat_index
    from s2geometry_b200.pointindex import ClosestPointQuery, PointIndex
    from s2geometry_b200.shapes import ShapeSoup
    soup = ShapeSoup()
    for loops in synth.many_polygons(60, 21, min_vertices=8, max_vertices=90, min_radius_km=150.0, max_radius_km=1200.0):
        soup.add_polygon(loops)
    c = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    soup.add_polyline(synth.cap_points(c, 0.3, 40, 5))
    soup.add_points(synth.sphere_points(25, 6))
    soup = soup.freeze()
    rsi = ref.index_create(soup)
    pts = np.concatenate([synth.sphere_points(250_000, 31), synth.cap_points(c, 0.4, 150_000, 32), soup.vertices[::7]])
    rpi = ref.point_index(pts)
    pix = PointIndex(pts)
    small_pts = pts[:3000]
    rpi_small = ref.point_index(small_pts)
    pix_small = PointIndex(small_pts)
    for max_edges in (0, 1):
        six = ShapeIndex(build_flat_index(soup, max_edges))
        for k, r, interiors in ((0, 0.004, True), (0, 0.004, False), (10, math.inf, True), (32, 0.02, False), (100, 0.003, True),
                                (5, math.inf, False), (1, math.inf, True), (200, math.inf, False), (0, 1e-9, True)):
            want = rpi.closest_shape_index(rsi, k, -1.0 if math.isinf(r) else r, interiors)
            got = ClosestPointQuery(pix, r).find_shape_index(six, k, interiors)
            _same_results(got, want, truncated=k > 0 and len(want[0]) == k)
            if k == 0 and r > 1e-6:
                assert len(want[0]) > 1000
        # no limit at all: every index point with its distance (small index)
        want = rpi_small.closest_shape_index(rsi, 0, -1.0, True)
        got = ClosestPointQuery(pix_small, math.inf).find_shape_index(six, 0, True)
        assert len(got[0]) == len(small_pts)
        _same_results(got, want, truncated=False)
        six.close()


def test_point_distances_to_a_shape_index_match_the_reference_edge_query(torch_cuda, ref):
    """S2ClosestEdgeQuery::GetDistance(PointTarget(p)) for batches of points (s2b_index_point_distances): polygons with
    holes + a polyline + a point shape, with and without interiors, several distance limits and none, host and device
    arrays — bit-identical distances (and the same points beyond the limit)."""
    torch = torch_cuda
    from s2geometry_b200.index import ShapeIndex, build_flat_index
    from s2geometry_b200.shapes import ShapeSoup
    soup = ShapeSoup()
    for loops in synth.many_polygons(60, 21, min_vertices=8, max_vertices=90, min_radius_km=150.0, max_radius_km=1200.0):
        soup.add_polygon(loops)
    c = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    soup.add_polyline(synth.cap_points(c, 0.3, 40, 5))
    soup.add_points(synth.sphere_points(25, 6))
    soup = soup.freeze()
    rsi = ref.index_create(soup)
    pts = np.concatenate([synth.sphere_points(200_000, 41), synth.cap_points(c, 0.4, 100_000, 42), soup.vertices[::5]])
    th = ref.hardware_threads()
    for max_edges in (0, 1):
        six = ShapeIndex(build_flat_index(soup, max_edges))
        for r, interiors in ((0.003, True), (0.05, True), (0.05, False), (1e-9, True), (0.5, True)):
            want = rsi.point_distances(pts, r, interiors, th)
            got = six.point_distances(pts, r, interiors)
            assert (got.view(np.uint64) == want.view(np.uint64)).all(), (max_edges, r, interiors, int((got != want).sum()))
            assert np.isinf(want).any() and (want == 0).sum() > (5000 if interiors else 100)   # inside polygons / on vertices
        dev = six.point_distances(torch.from_numpy(pts).cuda(), 0.05, True).cpu().numpy()
        assert (dev.view(np.uint64) == rsi.point_distances(pts, 0.05, True, th).view(np.uint64)).all()
        small = pts[::150]
        want = rsi.point_distances(small, -1.0, True, th)                     # no limit: every point gets its distance
        got = six.point_distances(small)
        assert (got.view(np.uint64) == want.view(np.uint64)).all() and np.isfinite(got).all()
        six.close()
