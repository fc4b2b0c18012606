"""S2CellIndex builder + traversal (tests/hostsim build of s2b_cellindex.cuh) against the reference's own S2CellIndex /
S2RegionSharder on the CPU. The same scenes run through the C ABI on the GPU in tests/test_cellindex_gpu.py."""
import ctypes

import numpy as np

from oracle.reflib import RefCellIndex, ref_cellunion_bound, ref_sharder_most_intersecting
from tests import cellindex_scenes as cs
from tests.util import vp


def hs_visit(hs, ids, labels, targets):
    ids = np.ascontiguousarray(ids, np.uint64); labels = np.ascontiguousarray(labels, np.int32)
    if isinstance(targets, np.ndarray):
        t_ids, t_off, T = np.ascontiguousarray(targets, np.uint64), None, len(targets)
    else:
        lens = [len(t) for t in targets]
        t_ids = np.ascontiguousarray(np.concatenate(targets) if lens else np.zeros(0, np.uint64), np.uint64)
        t_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32); T = len(lens)
    hs.hs_cell_index_visit.restype = ctypes.c_uint64
    out_off = np.zeros(T + 1, np.uint64)
    cap = 1 << 16
    while True:
        cells = np.empty(cap, np.uint64); labs = np.empty(cap, np.int32)
        total = hs.hs_cell_index_visit(vp(ids), vp(labels), ctypes.c_size_t(len(ids)), vp(t_ids), vp(t_off), ctypes.c_size_t(T), vp(out_off),
                                       vp(cells), vp(labs), ctypes.c_uint64(cap))
        assert total != 2 ** 64 - 1
        if total <= cap:
            return out_off, cells[:total].copy(), labs[:total].copy()
        cap = int(total)


def check_scene(hs, ref, ids, labels, targets):
    ri = RefCellIndex(ref, ids, labels)
    want = ri.visit(targets)
    got = hs_visit(hs, ids, labels, targets)
    assert (got[0] == want[0]).all()
    assert (got[1] == want[1]).all() and (got[2] == want[2]).all()
    return int(want[0][-1])


def test_reference_known_cases(hostsim, ref):
    """s2cell_index_test.cc: Empty, OneFaceCell, OneLeafCell, DuplicateValues, DisjointCells, NestedCells, IntersectionOptimization."""
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
        check_scene(hostsim, ref, ids, labels, targets)
        # every indexed cell, its parents and children as single-cell targets
        if len(ids):
            singles = np.unique(np.concatenate([ids] + [cs.parent(ids[ids >= 0], l) for l in (0, 1, 2)]))
            check_scene(hostsim, ref, ids, labels, singles)


def test_random_and_semi_random_unions(hostsim, ref):
    """IntersectionRandomUnions / IntersectionSemiRandomUnions (s2cell_index_test.cc:400-434)."""
    rng = np.random.default_rng(7)
    ids, labels = [], []
    for i in range(100):
        u = cs.random_union(rng, 10)
        ids.append(u); labels.append(np.full(len(u), i, np.int32))
    ids, labels = np.concatenate(ids), np.concatenate(labels)
    targets = [cs.random_union(rng, 10) for _ in range(200)] + [np.zeros(0, np.uint64)]
    assert check_scene(hostsim, ref, ids, labels, targets) > 1000
    for it in range(200):
        ids, labels, target = cs.semi_random_scene(rng)
        check_scene(hostsim, ref, ids, labels, [target])


def test_clustered_scene_single_cells_and_unions(hostsim, ref):
    rng = np.random.default_rng(11)
    ids, labels = cs.clustered_scene(rng, 20000, 500)
    singles = np.concatenate([cs.random_cells(rng, 3000), ids[:3000], cs.parent(ids[:2000], rng.integers(0, 31, 2000)),
                              cs.random_leaves(rng, 2000), (ids[:1000] - (ids[:1000] & (~ids[:1000] + np.uint64(1))) + np.uint64(1))])
    n = check_scene(hostsim, ref, ids, labels, singles)
    assert n > 20000
    unions = [cs.normalize_sorted_disjoint(np.concatenate([cs.random_cells(rng, 5, 2, 30), ids[rng.integers(0, len(ids), 6)]])) for _ in range(400)]
    check_scene(hostsim, ref, ids, labels, unions)
    # invalid targets report nothing
    bad = np.array([0, (7 << 61) | 1, (1 << 61) | 2], np.uint64)
    off, cells, labs = hs_visit(hostsim, ids, labels, bad)
    assert off[-1] == 0


def sharder_scene(ref, rng):
    shards = [ref.cellunion_normalize(cs.random_cells(rng, 12, 1, 12)) for _ in range(40)]
    regions = [cs.normalize_sorted_disjoint(cs.random_cells(rng, int(rng.integers(1, 8)), 0, 14)) for _ in range(300)]
    # small regions near shard cells, so that bound cells contain / are contained in shard cells at many levels
    anchor = np.concatenate(shards)
    for _ in range(300):
        a = anchor[rng.integers(0, len(anchor))]
        leaf = (a - ((a & (~a + np.uint64(1))) - np.uint64(1))) + (rng.integers(0, 1 << 30, dtype=np.uint64) << np.uint64(1))
        regions.append(cs.normalize_sorted_disjoint(cs.parent(np.array([leaf, leaf + np.uint64(2 << 20)], np.uint64), rng.integers(8, 25, 2))))
    return shards, regions


def test_sharder_matches_reference(hostsim, ref):
    """s2region_sharder_test.cc:82-126,165-193 plus random shards, regions = S2CellUnions (bound from the reference's own
    GetCellUnionBound, which for an S2CellUnion is its cap bound's covering, s2cell_union.cc:269-273)."""
    fpl = lambda f, p, l: ref.lib.ref_cellid_from_face_pos_level(f, p, l)
    shards = [np.array([fpl(0, 0, 10)], np.uint64), np.array([fpl(1, 1, 9), fpl(3, 0, 8)], np.uint64), np.array([fpl(5, 0, 10)], np.uint64)]
    regions = [np.array([fpl(0, 0, 11)], np.uint64), np.array([fpl(0, 0, 10), fpl(3, 0, 9), fpl(3, 1, 9)], np.uint64), np.array([fpl(4, 0, 10)], np.uint64)]
    want = ref_sharder_most_intersecting(ref, shards, regions, 42)
    assert list(want) == [0, 1, 42]
    bounds = [ref_cellunion_bound(ref, r) for r in regions]
    assert list(hs_best(hostsim, shards, bounds, regions, 42)) == [0, 1, 42]
    c0, c1 = cs.debug_id("0/0"), cs.debug_id("1/0")
    for sh in ([np.array([c1], np.uint64), np.array([c0], np.uint64)], [np.array([c0], np.uint64), np.array([c1], np.uint64)]):
        reg = [np.array([c0, c1], np.uint64)]
        assert list(ref_sharder_most_intersecting(ref, sh, reg, 42)) == [0]
        assert list(hs_best(hostsim, sh, [ref_cellunion_bound(ref, reg[0])], reg, 42)) == [0]
    shards, regions = sharder_scene(ref, np.random.default_rng(5))
    want = ref_sharder_most_intersecting(ref, shards, regions, -7)
    got = hs_best(hostsim, shards, [ref_cellunion_bound(ref, r) for r in regions], regions, -7)
    assert (got == want).all()
    assert len(set(want.tolist())) > 10 and (want == -7).any()


def hs_best(hs, shards, bounds, regions, default):
    ids = np.ascontiguousarray(np.concatenate(shards), np.uint64)
    labels = np.ascontiguousarray(np.concatenate([np.full(len(s), i, np.int32) for i, s in enumerate(shards)]), np.int32)
    csr = lambda L: (np.ascontiguousarray(np.concatenate(L), np.uint64), np.concatenate([[0], np.cumsum([len(r) for r in L])]).astype(np.uint32))
    b_ids, b_off = csr(bounds)
    r_ids, r_off = csr(regions)
    out = np.empty(len(regions), np.int32)
    hs.hs_cell_index_best_label(vp(ids), vp(labels), ctypes.c_size_t(len(ids)), vp(b_ids), vp(b_off), vp(r_ids), vp(r_off),
                                ctypes.c_size_t(len(regions)), int(default), vp(out))
    return out


def test_leaf_path_matches_visit(hostsim, ref):
    """The point path (leaf range table + flat label lists / chain walk) reports the labels VisitIntersectingCells gives for
    the leaf cell, on scenes with nesting, duplicates, leaf-level pairs and range boundaries."""
    rng = np.random.default_rng(13)
    hostsim.hs_cell_index_leaf_labels.restype = ctypes.c_uint64
    for ids, labels in (cs.clustered_scene(rng, 20000, 500), cs.clustered_scene(rng, 300, 7, n_seeds=3),
                        (np.array([cs.debug_id("1/")], np.uint64), np.array([4], np.int32)), (np.zeros(0, np.uint64), np.zeros(0, np.int32))):
        ids = np.ascontiguousarray(ids, np.uint64); labels = np.ascontiguousarray(labels, np.int32)
        lsb = ids & (~ids + np.uint64(1))
        edges = np.concatenate([ids - (lsb - np.uint64(1)), ids + (lsb - np.uint64(1))])           # first / last leaf of every cell
        nxt = edges[edges < np.uint64(0xBFFFFFFFFFFFFFFF)] + np.uint64(2)
        prv = edges[edges > np.uint64(1)] - np.uint64(2)
        leaves = np.concatenate([cs.random_leaves(rng, 5000), edges, nxt, prv, np.array([1, 0xBFFFFFFFFFFFFFFF], np.uint64)])
        w_off, w_cells, w_labels = RefCellIndex(ref, ids, labels).visit(leaves)
        for t in range(len(leaves)):                                                                   # canonical order: by label
            w_labels[int(w_off[t]):int(w_off[t + 1])].sort()
        for force_chain in (0, 1):
            off = np.zeros(len(leaves) + 1, np.uint64)
            out = np.empty(max(1, len(w_labels)), np.int32)
            total = hostsim.hs_cell_index_leaf_labels(vp(ids), vp(labels), ctypes.c_size_t(len(ids)), vp(leaves), ctypes.c_size_t(len(leaves)),
                                                      force_chain, vp(off), vp(out), ctypes.c_uint64(len(out)))
            assert total == len(w_labels)
            assert (off == w_off).all() and (out[:total] == w_labels).all()
