"""GPU parity tests for the batched S2ShapeIndexRegion cell tests (s2b_region.cu) through the C ABI against the
reference's own S2ShapeIndexRegion (oracle/_ref): MayIntersect(S2Cell), Contains(S2Cell), VisitIntersectingShapeIds."""
import numpy as np
import pytest
import torch

from tests import region_scenes, scenes

pytestmark = pytest.mark.gpu


def test_region_cell_tests_vs_reference(ref):
    from s2geometry_b200 import _lib
    from s2geometry_b200.index import FlatIndex, ShapeIndex, ShapeIndexRegion
    for name, soup in region_scenes.region_scene_list():
        ri = ref.index_create(soup)
        flat = FlatIndex(**ri.flatten())
        region = ShapeIndexRegion(ShapeIndex(flat))
        targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 5)
        before = _lib.launch_count()
        want_may, want_con = ri.region_cells(targets, 0), ri.region_cells(targets, 1)
        assert (region.may_intersect(targets) == want_may).all(), name
        assert (region.contains_cells(targets) == want_con).all(), name
        t = torch.from_numpy(targets.view(np.int64)).cuda()
        assert (region.may_intersect(t).cpu().numpy() == want_may).all(), name
        assert (region.contains_cells(t).cpu().numpy() == want_con).all(), name
        off, shp, con = region.intersecting_shapes(targets)
        woff, wshp, wcon = ri.region_intersecting_shapes(targets)
        assert (off == woff).all() and (shp == wshp).all() and (con == wcon).all(), name
        assert _lib.launch_count() >= before + 8
        assert (region.cell_union_bound() == ri.region_cell_union_bound()).all(), name
        # empty batch and a batch of one DISJOINT / one face cell
        assert len(region.may_intersect(np.zeros(0, np.uint64))) == 0
        off0, shp0, con0 = region.intersecting_shapes(np.zeros(0, np.uint64))
        assert list(off0) == [0] and len(shp0) == 0


def test_region_over_the_device_granularity_index(ref):
    """A finer index (max_edges_per_cell = 1) changes which targets are SUBDIVIDED, exactly as it does for the reference
    built with the same option: the answers are compared against the reference index of that granularity."""
    from s2geometry_b200.index import FlatIndex, ShapeIndex, ShapeIndexRegion
    soup, _ = scenes.loop_scene(3000)
    ri = ref.index_create(soup, 1)
    flat = FlatIndex(**ri.flatten())
    region = ShapeIndexRegion(ShapeIndex(flat))
    targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 6)
    assert (region.may_intersect(targets) == ri.region_cells(targets, 0)).all()
    assert (region.contains_cells(targets) == ri.region_cells(targets, 1)).all()


def test_edges_that_need_robust_cross_prods_extended_stages(ref):
    """Edges whose endpoints are 1-3 ulps apart (RobustCrossProd leaves its double-precision stage for the x87 long double
    and exact stages, which now run on the device) in a polygon and in polylines: the region cell tests over the
    reference-built index equal the reference's on every target; no index is refused any more."""
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, ShapeIndexRegion
    from s2geometry_b200.shapes import ShapeSoup
    from tests.test_region_cpu import _nudge
    rng = np.random.default_rng(4)
    loop = scenes._norm(np.array([[1.0, 1e-3, 0.0], [1.0, 1e-3, 0.0], [1.0, 0.2, 0.1], [1.0, 0.2, 0.1], [1.0, 0.0, 0.3]]))
    loop[1] = _nudge(loop[0:1], np.array([[0, 1, 0]]))[0]          # 1 ulp from vertex 0
    loop[3] = _nudge(loop[2:3], np.array([[-1, 2, 1]]))[0]
    soup = ShapeSoup(); soup.add_polygon([loop])
    for k in range(40):                                              # short polylines with a degenerate first edge across the sphere
        p = scenes._norm(rng.standard_normal((1, 3)))
        q = _nudge(p, rng.integers(-3, 4, (1, 3)))
        r = scenes._norm(p + 0.05 * rng.standard_normal((1, 3)))
        if (q == p).all():
            q = _nudge(p, np.array([[1, 0, -1]]))
        soup.add_polyline(np.concatenate([p, q, r]))
    soup = soup.freeze()
    ri = ref.index_create(soup)
    flat = FlatIndex(**ri.flatten())
    index = ShapeIndex(flat)
    region = ShapeIndexRegion(index)
    targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 6)
    # plus the cells around every degenerate vertex at all levels
    leaf = ref.encode_points(soup.vertices, 30)
    extra = np.concatenate([ref.parent(leaf, lvl) for lvl in range(0, 31, 3)])
    targets = np.concatenate([targets, extra])
    for mode, fn in ((0, region.may_intersect), (1, region.contains_cells)):
        want = ri.region_cells(targets, mode)
        assert (fn(targets) == want).all()
        assert 0 < want.sum() < len(want)
    assert (ContainsPointQuery(index, 1).contains(soup.vertices) == ri.contains(soup.vertices, 1)).all()


def test_fixed_level_covering_equals_the_reference_coverer(ref):
    """The device-side subdivision loop (GetCellUnionBound -> MayIntersect -> children ...) returns the covering
    S2RegionCoverer::GetCovering(S2ShapeIndexRegion) gives with min_level = max_level = level."""
    from s2geometry_b200.index import FlatIndex, ShapeIndex, ShapeIndexRegion
    total = 0
    for name, soup in region_scenes.region_scene_list():
        ri = ref.index_create(soup)
        region = ShapeIndexRegion(ShapeIndex(FlatIndex(**ri.flatten())))
        for level in (0, 2, 5, 8, 10):
            want = ri.region_covering(level)
            got = region.covering_at_level(level)
            assert len(got) == len(want) and (got == want).all(), (name, level, len(got), len(want))
            total += len(want)
    assert total > 10000
