"""GPU parity tests for the batched S2CellIndex joins (s2b_cellindex.cu) through the C ABI against the reference's own
S2CellIndex / S2RegionSharder (oracle/_ref): VisitIntersectingCells, GetIntersectingLabels, GetMostIntersectingShard and the
fused point -> label histogram."""
import numpy as np
import pytest

from oracle.reflib import RefCellIndex, ref_cellunion_bound, ref_sharder_most_intersecting
from tests import cellindex_scenes as cs
from tests.test_cellindex_cpu import sharder_scene

pytestmark = pytest.mark.gpu


def check_scene(ref, ids, labels, targets):
    from s2geometry_b200.cellindex import CellIndex
    ri = RefCellIndex(ref, ids, labels)
    ix = CellIndex(ids, labels)
    assert ix.num_cells == len(ids)
    w_off, w_cells, w_labels = ri.visit(targets)
    off, cells, labs = ix.intersecting_cells(targets)
    assert (off.astype(np.uint64) == w_off).all()
    cells, labs = cs.sort_pairs_per_target(off, cells, labs)
    assert (cells == w_cells).all() and (labs == w_labels).all()
    l_off, l_labels = ri.labels(targets)
    off2, labs2 = ix.intersecting_labels(targets)
    assert (off2.astype(np.uint64) == l_off).all() and (labs2 == l_labels).all()
    ix.close()
    return int(w_off[-1])


def test_reference_known_cases(ref):
    D = cs.debug_id
    faces = np.array([(f << 61) | (1 << 60) for f in range(6)], np.uint64)
    scenes = [
        ([], []),
        ([D("0/")], [0]),
        ([D("1/012301230123012301230123012301")], [12]),
        ([D("0/")] * 3 + [D("1/")], [0, 0, 1, 17]),
        ([D("0/"), D("3/"), D("5/0"), D("5/1")], [0, 0, 1, 2]),
        ([D("1/"), D("1/0"), D("1/012"), D("1/01"), D("1/0123"), D("1/0"), D("1/03"), D("1/02")], [3, 15, 9, 12, 6, 3, 4, 0]),
        ([D("1/001"), D("1/333"), D("2/00"), D("2/0232")], [1, 2, 3, 4]),
    ]
    targets = [faces, np.array([D("1/010"), D("1/3")], np.uint64), np.array([D("2/010"), D("2/011"), D("2/02")], np.uint64),
               np.array([D("1/012301230123012301230123012301")], np.uint64), np.array([D("1/0123")], np.uint64), np.zeros(0, np.uint64)]
    for ids, labels in scenes:
        ids = np.array(ids, np.uint64); labels = np.array(labels, np.int32)
        check_scene(ref, ids, labels, targets)
        if len(ids):
            check_scene(ref, ids, labels, np.unique(np.concatenate([ids] + [cs.parent(ids, l) for l in (0, 1, 2)])))


def test_random_and_semi_random_unions(ref):
    rng = np.random.default_rng(7)
    ids, labels = [], []
    for i in range(100):
        u = cs.random_union(rng, 10)
        ids.append(u); labels.append(np.full(len(u), i, np.int32))
    ids, labels = np.concatenate(ids), np.concatenate(labels)
    targets = [cs.random_union(rng, 10) for _ in range(200)] + [np.zeros(0, np.uint64)]
    assert check_scene(ref, ids, labels, targets) > 1000
    for it in range(40):
        ids, labels, target = cs.semi_random_scene(rng)
        check_scene(ref, ids, labels, [target])


def test_clustered_scene(ref):
    rng = np.random.default_rng(11)
    ids, labels = cs.clustered_scene(rng, 200000, 3000, n_seeds=400)
    # random cells, indexed cells, their ancestors at every level (a few of them very coarse: a level-0 target returns every
    # pair of its face), random leaves and the first leaf of indexed cells. The scene nests deeply (~2000 pairs per target):
    # 15k targets give ~30M pairs; eight times as many spent 100 s in the reference and the per-target sort.
    singles = np.concatenate([cs.random_cells(rng, 4000, 5, 30), cs.random_cells(rng, 100, 0, 4), ids[:4000],
                              cs.parent(ids[:3000], rng.integers(5, 31, 3000)), cs.parent(ids[:100], rng.integers(0, 5, 100)),
                              cs.random_leaves(rng, 3000), (ids[:1500] - (ids[:1500] & (~ids[:1500] + np.uint64(1))) + np.uint64(1))])
    assert check_scene(ref, ids, labels, singles) > 200000
    unions = [cs.normalize_sorted_disjoint(np.concatenate([cs.random_cells(rng, 5, 2, 30), ids[rng.integers(0, len(ids), 6)]])) for _ in range(600)]
    check_scene(ref, ids, labels, unions)


def test_device_path_matches_host_path(ref):
    import torch
    from s2geometry_b200.cellindex import CellIndex
    rng = np.random.default_rng(17)
    ids, labels = cs.clustered_scene(rng, 50000, 900, n_seeds=100)
    ix = CellIndex(ids, labels)
    targets = np.concatenate([cs.random_cells(rng, 20000), ids[:20000], cs.random_leaves(rng, 5000)])
    off, cells, labs = ix.intersecting_cells(targets)
    d_off, d_cells, d_labs = ix.intersecting_cells(torch.from_numpy(targets.view(np.int64)).cuda())
    assert (d_off.cpu().numpy().astype(np.uint32) == off).all()
    assert (d_cells.cpu().numpy().view(np.uint64) == cells).all() and (d_labs.cpu().numpy() == labs).all()
    assert off[-1] > 4 * len(targets)          # the first capacity guess was too small: the retry path ran


def test_invalid_inputs(ref):
    from s2geometry_b200 import _lib
    from s2geometry_b200.cellindex import CellIndex
    rng = np.random.default_rng(3)
    ids, labels = cs.clustered_scene(rng, 1000, 10)
    ix = CellIndex(ids, labels)
    bad = np.array([0, (7 << 61) | 1, (1 << 61) | 2], np.uint64)
    off, cells, labs = ix.intersecting_cells(bad)
    assert off[-1] == 0 and len(cells) == 0
    with pytest.raises(_lib.S2BError):
        CellIndex(np.array([0], np.uint64), np.array([0], np.int32))           # invalid cell id
    with pytest.raises(_lib.S2BError):
        CellIndex(ids[:1], np.array([-1], np.int32))                            # negative label
    empty = CellIndex(np.zeros(0, np.uint64), np.zeros(0, np.int32))
    off, cells, labs = empty.intersecting_cells(ids[:10])
    assert off[-1] == 0
    assert len(empty.point_histogram(np.random.default_rng(1).standard_normal((100, 3)))) == 0


def test_sharder_matches_reference(ref):
    from s2geometry_b200.cellindex import RegionSharder
    fpl = lambda f, p, l: ref.lib.ref_cellid_from_face_pos_level(f, p, l)
    shards = [np.array([fpl(0, 0, 10)], np.uint64), np.array([fpl(1, 1, 9), fpl(3, 0, 8)], np.uint64), np.array([fpl(5, 0, 10)], np.uint64)]
    regions = [np.array([fpl(0, 0, 11)], np.uint64), np.array([fpl(0, 0, 10), fpl(3, 0, 9), fpl(3, 1, 9)], np.uint64), np.array([fpl(4, 0, 10)], np.uint64)]
    sh = RegionSharder(shards)
    bounds = [ref_cellunion_bound(ref, r) for r in regions]
    assert list(sh.most_intersecting_shard(bounds, regions, 42)) == [0, 1, 42]          # s2region_sharder_test.cc:82-126
    assert list(sh.most_intersecting_shard(regions, None, 42)) == [0, 1, 42]
    off, labs = sh.intersecting_shards(regions)                                          # :128-163 (exact covering as the bound)
    assert [sorted(labs[off[t]:off[t + 1]].tolist()) for t in range(3)] == [[0], [0, 1], []]
    c0, c1 = cs.debug_id("0/0"), cs.debug_id("1/0")
    for pair in ([c1, c0], [c0, c1]):                                                    # :165-193 tie-breaking
        s2 = RegionSharder([np.array([pair[0]], np.uint64), np.array([pair[1]], np.uint64)])
        reg = [np.array([c0, c1], np.uint64)]
        assert list(s2.most_intersecting_shard([ref_cellunion_bound(ref, reg[0])], reg, 42)) == [0]
    shards, regions = sharder_scene(ref, np.random.default_rng(5))
    want = ref_sharder_most_intersecting(ref, shards, regions, -7)
    got = RegionSharder(shards).most_intersecting_shard([ref_cellunion_bound(ref, r) for r in regions], regions, -7)
    assert (got == want).all()


def test_point_histogram_vs_reference(ref):
    """Geofencing join: 2M points against the coverings of 2000 caps (label = cap), host and device forms."""
    import torch
    from s2geometry_b200 import synth
    from s2geometry_b200.cellindex import CellIndex
    rng = np.random.default_rng(21)
    centers = rng.standard_normal((2000, 3)); centers /= np.linalg.norm(centers, axis=1)[:, None]
    ids, labels = [], []
    for i, c in enumerate(centers):
        cov = ref.cover_cap(c, float(10 ** rng.uniform(-3.0, -0.8)), max_cells=int(rng.integers(4, 40)))
        ids.append(cov); labels.append(np.full(len(cov), i, np.int32))
    ids, labels = np.concatenate(ids), np.concatenate(labels)
    pts = synth.sphere_points(2_000_000, seed=9)
    near = centers[rng.integers(0, len(centers), 500_000)] + 0.01 * rng.standard_normal((500_000, 3))
    pts = np.ascontiguousarray(np.concatenate([pts, near / np.linalg.norm(near, axis=1)[:, None]]))
    want = RefCellIndex(ref, ids, labels).point_histogram(pts)
    ix = CellIndex(ids, labels)
    got = ix.point_histogram(pts)
    assert (got == want).all() and want.sum() > 100000
    dev = ix.point_histogram(torch.from_numpy(pts).cuda())
    assert (dev.cpu().numpy().astype(np.uint64) == want).all()
    # more labels than the CTA-private histogram holds: the global-atomics variant
    big = CellIndex(ids, labels + 30000)
    got = big.point_histogram(pts)
    assert len(got) == 32000 and (got[30000:] == want).all() and not got[:30000].any()


def test_point_histogram_on_leaf_boundaries(ref):
    """Points on / within 1e-8..1e-4 leaf cells of leaf-cell boundaries (where the filtered projection must fall back to the
    IEEE sequence) against an index whose cells are the fine ancestors of those very leaves: range boundaries coincide
    with the boundaries the points sit on."""
    from s2geometry_b200.cellindex import CellIndex
    from tests.util import leaf_boundary_points
    rng = np.random.default_rng(31)
    pts = leaf_boundary_points(400_000, 17)
    leaves = ref.encode_points(pts, 30, 0)
    k = 150_000
    ids = np.concatenate([cs.parent(leaves[:k], rng.integers(22, 31, k)), cs.parent(leaves[:k], rng.integers(2, 22, k))])
    labels = rng.integers(0, 700, len(ids)).astype(np.int32)
    want = RefCellIndex(ref, ids, labels).point_histogram(pts)
    got = CellIndex(ids, labels).point_histogram(pts)
    assert (got == want).all() and want.sum() > 2 * k
