"""GPU parity tests for K2/K3/K4 (Locate, containment, histogram, cell union) through the C ABI."""
import numpy as np
import pytest

from tests import scenes, util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def make_index(ref_index):
    from s2geometry_b200.index import FlatIndex, ShapeIndex
    return ShapeIndex(FlatIndex(**ref_index.flatten()))


def test_vertex_model_truth_tables(torch_cuda, ref):
    from s2geometry_b200.index import ContainsPointQuery
    ka = scenes.known_answers()
    ri = ref.index_create(scenes.text_index(ka["index"]))
    ix = make_index(ri)
    q = scenes.latlng_deg_to_point(ka["queries"])
    for name, m in scenes.MODELS.items():
        query = ContainsPointQuery(ix, m)
        want = np.array(ka["contains"][name], np.uint8)
        assert (query.contains(q) == want).all(), name
        assert (query.contains(torch_cuda.from_numpy(q).cuda()).cpu().numpy() == want).all(), name
    for row in ka["shape_contains"]:
        query = ContainsPointQuery(ix, scenes.MODELS[row["model"]])
        assert query.shape_contains(row["shape"], scenes.latlng_deg_to_point([row["q"]]))[0] == row["expected"]
    t = ka["closed_three_shapes"]
    ix2 = make_index(ref.index_create(scenes.text_index(t["index"])))
    off, ids = ContainsPointQuery(ix2, 2).containing_shape_ids(scenes.latlng_deg_to_point([t["q"]]))
    assert list(off) == [0, 3] and list(ids) == t["containing_shape_ids_closed"]


def test_golden_index_fixture(torch_cuda):
    """Committed fixture: the reference-built index of the 10k-vertex loop and the reference's answers."""
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex
    fx = util.fixture("index_loop10k.npz")
    flat = FlatIndex(**{k: fx[k] for k in FlatIndex.FIELDS}, num_shape_ids=int(fx["num_shape_ids"]))
    ix = ShapeIndex(flat)
    info = ix.info()
    assert info["num_cells"] == flat.num_cells and info["num_edges"] == flat.num_edges
    q = fx["queries"]
    for name, m in scenes.MODELS.items():
        assert (ContainsPointQuery(ix, m).contains(q) == fx["contains_" + name]).all(), name
    assert (ContainsPointQuery(ix, 1).contains(q) == fx["brute_force"]).all()
    # cell centres recomputed by the library (cell_center = NULL) give the same answers
    flat2 = FlatIndex(**{k: fx[k] for k in FlatIndex.FIELDS if k != "cell_center"}, num_shape_ids=int(fx["num_shape_ids"]))
    assert (ContainsPointQuery(ShapeIndex(flat2), 1).contains(q) == fx["contains_SEMI_OPEN"]).all()


def test_config3_loop_10k_vs_reference(torch_cuda, ref):
    """BASELINE config 3 at 2M points per set: U (uniform), B (near the loop), D (on vertices / edges)."""
    from s2geometry_b200.index import ContainsPointQuery, exact_stage_count
    soup, v = scenes.loop_scene(10000)
    ri = ref.index_create(soup)
    ix = make_index(ri)
    u, b, d = scenes.loop_points(v, 2_000_000, 2_000_000, 3)
    for pts in (u, b, d):
        for m in (0, 1, 2):
            got = ContainsPointQuery(ix, m).contains(pts)
            assert (got == ri.contains(pts, m)).all(), m
    dev = torch_cuda.from_numpy(b).cuda()
    assert (ContainsPointQuery(ix, 1).contains(dev).cpu().numpy() == ri.contains(b, 1)).all()
    assert (ContainsPointQuery(ix, 1).contains(d) == ri.brute_force_contains(0, d)).all()
    assert (ContainsPointQuery(ix, 1).shape_contains(0, b) == ri.shape_contains(0, b, 1)).all()


def test_axis_aligned_exact_stage_on_device(torch_cuda, ref):
    from s2geometry_b200.index import ContainsPointQuery, exact_stage_count
    ri = ref.index_create(scenes.axis_aligned_scene())
    ix = make_index(ri)
    q = scenes.axis_aligned_queries()
    exact_stage_count(reset=True)
    for m in (0, 1, 2):
        query = ContainsPointQuery(ix, m)
        assert (query.contains(q) == ri.contains(q, m, 1)).all(), m
        for s in range(ix.num_shape_ids):
            assert (query.shape_contains(s, q) == ri.shape_contains(s, q, m, 1)).all(), (m, s)
        assert (query.shape_histogram(q) == ri.shape_histogram(q, m, 1)).all()
    assert exact_stage_count() > 100   # the exact (superaccumulator) stage ran on the GPU — there is no CPU stage


def test_polygons_histogram_and_csr(torch_cuda, ref):
    from s2geometry_b200 import synth
    from s2geometry_b200.index import ContainsPointQuery
    soup = scenes.polygons_scene(2000, 41, min_radius_km=100, max_radius_km=1500)
    ri = ref.index_create(soup)
    ix = make_index(ri)
    pts = synth.sphere_points(3_000_000, 42)
    query = ContainsPointQuery(ix, 1)
    want = ri.shape_histogram(pts, 1)
    assert want.sum() > 100000
    assert (query.shape_histogram(pts) == want).all()
    dev = torch_cuda.from_numpy(pts).cuda()
    got_dev = query.shape_histogram(dev).cpu().numpy().view(np.uint64)
    assert (got_dev == want).all()
    sub = pts[:300000]
    off, ids = query.containing_shape_ids(sub)
    roff, rids = ri.containing_shapes(sub, 1)
    assert (off == roff).all() and (ids == rids).all()
    doff, dids = query.containing_shape_ids(dev[:300000])
    assert (doff.cpu().numpy().astype(np.uint32) == roff).all() and (dids.cpu().numpy() == rids).all()
    assert (query.contains(sub) == (np.diff(roff.astype(np.int64)) > 0)).all()


def test_cell_union_contains(torch_cuda, ref):
    from s2geometry_b200 import synth
    from s2geometry_b200.index import cell_union_contains
    center = np.array([0.4, 0.5, 0.76])
    ids = ref.cover_cap(center, 0.05, max_cells=1000)
    assert 100 < len(ids) <= 1000
    pts = np.concatenate([synth.cap_points(center, 0.08, 1_000_000, 9), synth.sphere_points(1_000_000, 10), ref.to_point(ids)])
    want = ref.cellunion_contains(ids, pts)
    assert 0.1 < want.mean() < 0.8
    assert (cell_union_contains(ids, pts) == want).all()
    assert (cell_union_contains(ids, torch_cuda.from_numpy(pts).cuda()).cpu().numpy() == want).all()
    # Trondheim known answer (python/s2geometry_test.py:254-259)
    ka = util.known_answers()["latlng_deg_in_cells"][0]
    cells = np.sort(np.array([int(c, 16) for c in ka["cells"]], np.uint64))
    p = ref.latlng_to_point(np.radians(np.array([[ka["lat"], ka["lng"]]])))
    assert cell_union_contains(cells, p)[0] == 1
    from s2geometry_b200.index import CellUnionIndex, ContainsPointQuery
    assert (ContainsPointQuery(CellUnionIndex(ids), 1).contains(pts) == want).all()   # the union as an index of interior cells
    big = ref.cover_cap(center, 0.15, max_cells=20000, min_level=9, max_level=9)   # 8868 ids (> 4096): global-memory path
    assert len(big) > 4096
    pts2 = synth.cap_points(center, 0.2, 500_000, 11)
    assert (cell_union_contains(big, pts2) == ref.cellunion_contains(big, pts2)).all()


def test_empty_inputs(torch_cuda, ref):
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, cell_union_contains
    from s2geometry_b200.shapes import ShapeSoup
    empty = FlatIndex(0, np.zeros(0, np.uint64), np.zeros(1, np.uint32), np.zeros(0, np.int32), np.zeros(0, np.uint8),
                      np.zeros(1, np.uint32), np.zeros((0, 3)), np.zeros((0, 3)))
    ix = ShapeIndex(empty)
    q = scenes.latlng_deg_to_point([[0, 0], [10, 10]])
    assert (ContainsPointQuery(ix).contains(q) == 0).all()
    assert ContainsPointQuery(ix).contains(np.zeros((0, 3))).shape == (0,)
    assert (cell_union_contains(np.zeros(0, np.uint64), q) == 0).all()
    # invalid index data is rejected, not crashed on
    from s2geometry_b200 import S2BError
    with pytest.raises(S2BError):
        ShapeIndex(FlatIndex(1, np.array([5, 3], np.uint64), np.array([0, 0, 0], np.uint32), np.zeros(0, np.int32), np.zeros(0, np.uint8),
                             np.zeros(1, np.uint32), np.zeros((0, 3)), np.zeros((0, 3))))


def test_full_size_properties_config3(torch_cuda, ref):
    """100M uniform points against the 10k-vertex loop, on the device (BASELINE config 3 size): size-independent
    properties + an id-by-id check of a 4M sample against the reference."""
    torch = torch_cuda
    from s2geometry_b200 import synth
    from s2geometry_b200.index import ContainsPointQuery
    soup, v = scenes.loop_scene(10000)
    ri = ref.index_create(soup)
    ix = make_index(ri)
    n = 100_000_000
    pts = synth.sphere_points_cuda(n, 3, "cuda")
    query = ContainsPointQuery(ix, 1)
    inside = query.contains(pts)
    again = query.contains(pts)
    assert bool((inside == again).all())                                  # deterministic
    hist = query.shape_histogram(pts)
    assert int(hist[0].item()) == int(inside.sum().item())                # histogram == number of contained points
    assert bool((query.shape_contains(0, pts) == inside).all())           # one shape: Contains == ShapeContains(0)
    frac = float(inside.double().mean().item())
    area = 2 * np.pi * (1 - np.cos(scenes.LOOP_RADIUS)) / (4 * np.pi)     # cap of the mean radius, within a few %
    assert 0.9 * area < frac < 1.15 * area
    # open <= semi-open <= closed pointwise
    o = ContainsPointQuery(ix, 0).contains(pts); c = ContainsPointQuery(ix, 2).contains(pts)
    assert bool((o <= inside).all()) and bool((inside <= c).all())
    sample = pts[:4_000_000].cpu().numpy()
    assert (inside[:4_000_000].cpu().numpy() == ri.contains(sample, 1)).all()


def test_own_builder_index_on_gpu(torch_cuda, ref):
    """Index built by the library's own host builder (no reference involved in the product path) -> same answers
    as the reference's S2ContainsPointQuery over the reference-built index."""
    from s2geometry_b200 import synth
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    soup, v = scenes.loop_scene(10000)
    ix = ShapeIndex(build_flat_index(soup))
    ri = ref.index_create(soup)
    u, b, d = scenes.loop_points(v, 1_000_000, 1_000_000, 13)
    for pts in (u, b, d):
        for m in (0, 1, 2):
            assert (ContainsPointQuery(ix, m).contains(pts) == ri.contains(pts, m)).all()
    soup2 = scenes.polygons_scene(3000, 51, min_radius_km=100, max_radius_km=1500)
    ix2 = ShapeIndex(build_flat_index(soup2))
    ri2 = ref.index_create(soup2)
    pts = synth.sphere_points(3_000_000, 52)
    want = ri2.shape_histogram(pts, 1)
    assert want.sum() > 100000
    assert (ContainsPointQuery(ix2, 1).shape_histogram(pts) == want).all()
    soup3 = scenes.axis_aligned_scene()
    ix3 = ShapeIndex(build_flat_index(soup3))
    ri3 = ref.index_create(soup3)
    q = scenes.axis_aligned_queries()
    for m in (0, 1, 2):
        assert (ContainsPointQuery(ix3, m).shape_histogram(q) == ri3.shape_histogram(q, m, 1)).all()


def test_device_granularity_index_gives_the_same_answers(torch_cuda, ref):
    """The granularity recommended for device-resident indexes (max_edges_per_cell = 1, S2B_MAX_EDGES_PER_CELL_DEVICE):
    far more, far smaller cells — every answer still equals the reference's over its default-granularity index, for all
    vertex models, incl. every vertex / edge midpoint / cell centre of the loop and the exact-stage scene."""
    from s2geometry_b200 import synth
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    soup, v = scenes.loop_scene(10000)
    fine = build_flat_index(soup, 1)
    assert fine.num_cells > 5 * build_flat_index(soup).num_cells
    ix = ShapeIndex(fine)
    ri = ref.index_create(soup)
    u, b, d = scenes.loop_points(v, 1_000_000, 2_000_000, 29)
    centres = fine.cell_center.reshape(-1, 3)                       # the fine index's own cell centres are queries too
    for pts in (u, b, d, centres):
        for m in (0, 1, 2):
            assert (ContainsPointQuery(ix, m).contains(pts) == ri.contains(pts, m)).all()
    soup2 = scenes.polygons_scene(3000, 53, min_radius_km=100, max_radius_km=1500)
    ix2 = ShapeIndex(build_flat_index(soup2, 1))
    ri2 = ref.index_create(soup2)
    pts = synth.sphere_points(3_000_000, 54)
    for m in (0, 1, 2):
        assert (ContainsPointQuery(ix2, m).shape_histogram(pts) == ri2.shape_histogram(pts, m)).all()
    off, ids = ContainsPointQuery(ix2, 1).containing_shape_ids(pts[:300_000])
    woff, wids = ri2.containing_shapes(pts[:300_000], 1)
    assert (np.asarray(off) == woff).all() and (np.asarray(ids) == wids).all()
    soup3 = scenes.axis_aligned_scene()
    ix3 = ShapeIndex(build_flat_index(soup3, 1))
    ri3 = ref.index_create(soup3)
    q = scenes.axis_aligned_queries()
    for m in (0, 1, 2):
        assert (ContainsPointQuery(ix3, m).shape_histogram(q) == ri3.shape_histogram(q, m, 1)).all()


def test_incident_edges_match_reference(torch_cuda, ref):
    """VisitIncidentEdges: for polygon vertices (two incident edges each, more where polygons share vertices), polyline
    ends, point shapes, -0.0 coordinates and ordinary points (none)."""
    torch = torch_cuda
    from s2geometry_b200 import S2BError, synth
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, build_flat_index
    from tests.test_index_wire import mixed_soup
    for name, soup in (("mixed", mixed_soup(51, 40)), ("axis", scenes.axis_aligned_scene()), ("loop", scenes.loop_scene(3000)[0])):
        ri = ref.index_create(soup)
        verts = np.asarray(soup.vertices).reshape(-1, 3)
        flipped = verts[:200].copy(); flipped[flipped == 0] = -0.0                # operator== treats -0 and +0 alike
        q = np.concatenate([verts, flipped, synth.sphere_points(100_000, 5)])
        woff, wsid, weid = ri.incident_edges(q)
        assert len(wsid) >= len(verts)
        for flat in (FlatIndex(**ri.flatten()), build_flat_index(soup, 1)):       # reference-built and own fine-grained index
            flat.edge_id = np.asarray(flat.edge_id if hasattr(flat, "edge_id") else ri.flatten()["edge_id"], np.int32)
            query = ContainsPointQuery(ShapeIndex(flat), 1)
            off, sid, eid = query.incident_edges(q)
            assert (off == woff).all(), name
            # the reference visits clipped shapes in cell order and edges in clipped order: same order for the same cells;
            # a different (finer) index may clip differently, so compare per-point sets there
            key = lambda o, s, e: [sorted(zip(s[o[i]:o[i + 1]].tolist(), e[o[i]:o[i + 1]].tolist())) for i in range(0, len(o) - 1, 97)]
            assert key(off, sid, eid) == key(woff, wsid, weid), name
            d_off, d_sid, d_eid = query.incident_edges(torch.from_numpy(q).cuda())
            assert (d_off.cpu().numpy().astype(np.uint32) == off).all() and (d_sid.cpu().numpy() == sid).all() and (d_eid.cpu().numpy() == eid).all()
        same = FlatIndex(**ri.flatten()); same.edge_id = ri.flatten()["edge_id"]
        off, sid, eid = ContainsPointQuery(ShapeIndex(same), 1).incident_edges(q)
        assert (sid == wsid).all() and (eid == weid).all(), name                  # identical order over identical cells
    bare = FlatIndex(**{k: v for k, v in ri.flatten().items() if k != "edge_id"})
    with pytest.raises(S2BError):
        ContainsPointQuery(ShapeIndex(bare), 1).incident_edges(q[:10])           # no edge ids attached


def test_locate_cells_matches_reference(torch_cuda, ref):
    """Iterator::Locate(S2CellId) for a batch of target cells: relation + the cell the iterator lands on."""
    torch = torch_cuda
    from s2geometry_b200.index import FlatIndex, ShapeIndex
    from tests.test_encode_cpu import random_ids_all_levels
    for soup in (scenes.loop_scene(10000)[0], scenes.polygons_scene(2000, 61, min_radius_km=100, max_radius_km=1500)):
        ri = ref.index_create(soup)
        flat = FlatIndex(**ri.flatten())
        ix = ShapeIndex(flat)
        ids = flat.cell_ids
        lsb = ids & (~ids + np.uint64(1))
        ok = lsb < (np.uint64(1) << np.uint64(60))
        parents = (ids[ok] & (~(lsb[ok] << np.uint64(2)) + np.uint64(1))) | (lsb[ok] << np.uint64(2))
        kids = (ids[lsb > 1].astype(np.int64) + (lsb[lsb > 1] >> np.uint64(2)).astype(np.int64)).astype(np.uint64)   # child 2
        targets = np.concatenate([random_ids_all_levels(ref, 500_000, 71), ids, parents, kids, parents[::7] + np.uint64(0)])
        want_rel, want_at = ri.locate_cells(targets)
        rel, cell = ix.locate_cells(targets)
        assert (rel == want_rel).all()
        hit = rel != 2
        assert (cell[~hit] == -1).all() and (ids[cell[hit]] == want_at[hit]).all()
        assert set(np.unique(rel)) == {0, 1, 2}
        drel, dcell = ix.locate_cells(torch.from_numpy(targets.view(np.int64)).cuda())
        assert (drel.cpu().numpy() == rel).all() and (dcell.cpu().numpy() == cell).all()
        # Locate(S2Point): the containing index cell of every point (uniform, near the geometry, on its vertices)
        from s2geometry_b200 import synth
        pts = np.concatenate([synth.sphere_points(1_000_000, 73), synth.cap_points(scenes.LOOP_CENTER, 0.25, 500_000, 74),
                              np.asarray(soup.vertices).reshape(-1, 3)[:100_000], flat.cell_center.reshape(-1, 3)])
        want = ri.locate_points(pts)
        got = ix.locate_points(pts)
        assert ((got >= 0) == (want != 0)).all() and (ids[got[got >= 0]] == want[want != 0]).all()
        assert (ix.locate_points(torch.from_numpy(pts).cuda()).cpu().numpy() == got).all()


def test_cells_with_hundreds_of_edges_on_gpu(torch_cuda, ref):
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    soup, v = scenes.loop_scene(3000)
    flat = build_flat_index(soup, 1000)
    assert np.diff(flat.clip_edge_begin.astype(np.int64)).max() > 200
    ix = ShapeIndex(flat)
    ri = ref.index_create(soup)
    u, b, d = scenes.loop_points(v, 20000, 300000, 35)
    for pts in (b, d):
        for m in (0, 1, 2):
            assert (ContainsPointQuery(ix, m).contains(pts) == ri.contains(pts, m)).all()
            assert (ContainsPointQuery(ix, m).shape_histogram(pts) == ri.shape_histogram(pts, m)).all()


def boundary_points(n, seed, levels=(2, 4, 6, 8, 9, 10, 12)):
    """S2Points whose leaf coordinates sit within a few thousand leaf cells of level-L cell boundaries (the FP32
    pre-filter's uncertainty radius is 2048 leaf cells), plus face edges and cube corners."""
    rng = np.random.default_rng(seed)
    out = []
    for lvl in levels:
        step = 1 << (30 - lvl)
        i = rng.integers(0, (1 << lvl) + 1, n) * step + rng.integers(-5000, 5001, n)
        j = rng.integers(0, 1 << 30, n)
        swap = rng.integers(0, 2, n).astype(bool)
        fi = np.clip(np.where(swap, j, i), 0, 1 << 30).astype(np.float64) + rng.uniform(0, 1, n)
        fj = np.clip(np.where(swap, i, j), 0, 1 << 30).astype(np.float64) + rng.uniform(0, 1, n)
        s_, t_ = np.minimum(fi / 2.0 ** 30, 1.0), np.minimum(fj / 2.0 ** 30, 1.0)
        st2uv = lambda s: np.where(s >= 0.5, (1 / 3.) * (4 * s * s - 1), (1 / 3.) * (1 - 4 * (1 - s) * (1 - s)))
        u, v = st2uv(s_), st2uv(t_)
        face = rng.integers(0, 6, n)
        one = np.ones(n)
        xyz = np.select([(face == k)[:, None] for k in range(6)],
                        [np.stack(c, 1) for c in ((one, u, v), (-u, one, v), (-u, -v, one), (-one, -v, -u), (v, -one, -u), (v, u, -one))])
        out.append(xyz / np.sqrt((xyz * xyz).sum(1))[:, None])
    return np.ascontiguousarray(np.vstack(out))


def test_fp32_prefilter_equals_exact_locate(torch_cuda, ref):
    """Production Locate (FP32 pre-filter + FP64 pass over the undecided points) == all-FP64 Locate, on uniform points,
    on points hugging cell/face boundaries at every level, and on garbage inputs; and both equal the reference."""
    torch = torch_cuda
    from s2geometry_b200 import _lib, synth
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    lib = _lib.load()
    scn = {"loop": scenes.loop_scene(10000)[0], "polys": scenes.polygons_scene(3000, 51, min_radius_km=100, max_radius_km=1500),
           "axis": scenes.axis_aligned_scene()}
    garbage = np.array([[0, 0, 0], [np.nan, 1, 0], [1, np.nan, np.nan], [np.inf, 1, 1], [1e-320, 0, 0], [1e300, 1e300, 1], [-0.0, 0, 1],
                        [1, 1, 1], [1, 1, 0], [-1, -1, -1], [1e-30, 1e-30, 1e-30]], float)
    pts = np.concatenate([synth.sphere_points(3_000_000, 61), boundary_points(200_000, 62), garbage,
                          synth.cap_points(scenes.LOOP_CENTER, 0.25, 1_000_000, 63)])
    dev = torch.from_numpy(pts).cuda()
    for name, soup in scn.items():
        ix = ShapeIndex(build_flat_index(soup))
        ri = ref.index_create(soup)
        q = ContainsPointQuery(ix, 1)
        try:
            fast = (q.contains(dev), q.shape_histogram(dev), q.containing_shape_ids(dev))
            assert lib.s2b_diag_force_exact_locate(1) == 0
            exact = (q.contains(dev), q.shape_histogram(dev), q.containing_shape_ids(dev))
        finally:
            lib.s2b_diag_force_exact_locate(0)
        assert bool((fast[0] == exact[0]).all()), name
        assert bool((fast[1] == exact[1]).all()), name
        assert bool((fast[2][0] == exact[2][0]).all()) and bool((fast[2][1] == exact[2][1]).all()), name
        finite = np.isfinite(pts).all(1)
        want = ri.contains(pts[finite], 1)
        assert (fast[0].cpu().numpy()[finite] == want).all(), name


def test_error_paths_on_gpu(torch_cuda):
    """Bad arguments are reported as errors (never crashes, never a silent fallback)."""
    import ctypes
    torch = torch_cuda
    from s2geometry_b200 import _lib, S2BError
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    lib = _lib.load()
    t = torch.zeros(64, dtype=torch.float64, device="cuda")
    out = torch.zeros(16, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # lat/lng input must be 16-byte aligned for the vector loads
    assert lib.s2b_encode_latlng_dev(ctypes.c_void_p(t.data_ptr() + 8), 4, 30, ctypes.c_void_p(out.data_ptr()), None, st) == -1
    assert b"aligned" in lib.s2b_last_error()
    assert lib.s2b_encode_points_dev(ctypes.c_void_p(t.data_ptr()), 4, 31, ctypes.c_void_p(out.data_ptr()), st) == -1
    assert lib.s2b_encode_points_dev(None, 4, 30, ctypes.c_void_p(out.data_ptr()), st) == -1
    soup, v = scenes.loop_scene(200)
    ix = ShapeIndex(build_flat_index(soup))
    assert lib.s2b_contains_dev(ix._h, ctypes.c_void_p(t.data_ptr()), 4, 7, ctypes.c_void_p(out.data_ptr()), st) == -1   # vertex model
    assert lib.s2b_contains_dev(None, ctypes.c_void_p(t.data_ptr()), 4, 1, ctypes.c_void_p(out.data_ptr()), st) == -1
    with pytest.raises(ValueError):
        ContainsPointQuery(ix, 1).contains(np.zeros((4, 2)))
    # an 8-byte aligned (not 16) xyz pointer is fine: the pre-filter's cp.async path is skipped, results identical
    pts = torch.from_numpy(np.ascontiguousarray(scenes.loop_points(v, 1000, 50001, 3)[1])).cuda()
    shifted = torch.empty(pts.numel() + 1, dtype=torch.float64, device="cuda")
    shifted[1:] = pts.reshape(-1)
    view = shifted[1:].reshape(-1, 3)
    assert view.data_ptr() % 16 == 8
    q = ContainsPointQuery(ix, 1)
    assert bool((q.contains(view) == q.contains(pts)).all())
    torch.cuda.synchronize()


def test_work_lists_hold_the_worst_case_mix_of_edge_and_interior_hits(torch_cuda):
    """2^25 + 12345 uniform points in ONE round against an index that covers the whole sphere with 80 % edge-bearing and
    20 % multi-clip interior cells: nearly every warp step appends to both work lists (which share one buffer from its two
    ends) and keeps retiring partly used reservation chunks — the case the list capacity has to be sized for
    (work_list_capacity in s2b_index.cu). Count / fill / histogram / per-shape results must equal what the cell of each
    point implies, for every point."""
    torch = torch_cuda
    from s2geometry_b200 import cellid, synth
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex
    level = 2
    ids = []
    for face in range(6):
        for pos in range(4 ** level):
            ids.append((face << 61) | (pos << (61 - 2 * level)) | (1 << (60 - 2 * level)))
    ids = np.array(sorted(ids), np.uint64)
    face_center = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], float)
    clip_begin, shape, flags, edge_begin, v0, v1 = [0], [], [], [0], [], []
    interior = np.arange(len(ids)) % 5 == 0
    for k, cid in enumerate(ids):
        if interior[k]:           # two clipped shapes, both containing the centre, no edges: an "interior list" hit
            shape += [0, 1]; flags += [1 | (2 << 1)] * 2; edge_begin += [edge_begin[-1]] * 2
        else:                     # one clipped shape with one (far away, tiny) edge: an "edge list" hit that crosses nothing
            far = face_center[(int(cid >> np.uint64(61)) + 3) % 6]
            a = far + np.array([1e-3, 2e-3, 3e-3]); b = far + np.array([2e-3, 1e-3, 3e-3])
            v0.append(a / np.linalg.norm(a)); v1.append(b / np.linalg.norm(b))
            shape += [0]; flags += [1 | (2 << 1)]; edge_begin += [edge_begin[-1] + 1]
        clip_begin.append(len(shape))
    flat = FlatIndex(2, ids, clip_begin, shape, flags, edge_begin, np.array(v0), np.array(v1))
    ix = ShapeIndex(flat)
    q = ContainsPointQuery(ix, 1)
    n = (1 << 25) + 12345
    pts = synth.stream_sphere_points_cuda(31, 0, n, "cuda")
    cell = ix.locate_points(pts).long()
    assert bool((cell >= 0).all())
    is_interior = torch.from_numpy(interior).cuda()[cell]
    want_count = torch.where(is_interior, 2, 1).to(torch.int32)
    off, got_ids = q.containing_shape_ids(pts)
    assert bool(((off[1:] - off[:-1]) == want_count).all())
    assert bool((got_ids[off[:-1].long()] == 0).all())                                   # first id of every point is shape 0
    second = (off[:-1].long() + 1)[is_interior]
    assert bool((got_ids[second] == 1).all())
    hist = q.shape_histogram(pts).cpu().numpy()
    assert hist[0] == n and hist[1] == int(is_interior.sum())
    assert bool((q.shape_contains(1, pts) == is_interior.to(torch.uint8)).all())
    assert bool((q.contains(pts) == 1).all())
