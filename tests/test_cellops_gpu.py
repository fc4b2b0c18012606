"""GPU parity tests for the batched S2CellId / S2CellUnion algebra (s2b_cellops.cu) against the reference:
AppendVertexNeighbors, AppendAllNeighbors, S2CellUnion::Normalize, S2CellUnion::Contains/Intersects(S2CellId)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def debug_id(s):
    """S2CellId::FromDebugString("f/ddd") (s2cell_id.cc:600-640)."""
    face, digits = s.split("/")
    id_ = (int(face) << 61) | (1 << 60)
    for k, d in enumerate(digits):
        # child(d): new lsb = lsb >> 2; id + (2*d + 1 - 4) * new_lsb
        lsb = id_ & -id_
        new_lsb = lsb >> 2
        id_ = id_ + (2 * int(d) + 1 - 4) * new_lsb
    return id_


def test_reference_known_answers(torch_cuda):
    """s2cell_id_test.cc:623-642 (vertex neighbours) and :660-695 (a cube-corner cell has 7 distinct neighbours)."""
    from s2geometry_b200 import cellid
    top = cellid.from_points(np.array([[0.0, 0.0, 1.0]]))
    nb, cnt = cellid.vertex_neighbors(top, 5)
    assert cnt[0] == 4
    # the four level-5 cells meeting at the centre of face 2 = the level-5 parents of the leaves around (0, 0, 1)
    around = cellid.from_points(np.array([[sx * 1e-9, sy * 1e-9, 1.0] for sx in (-1, 1) for sy in (-1, 1)]), 5)
    assert sorted(int(x) for x in nb[0]) == sorted(int(x) for x in around) and len(set(int(x) for x in around)) == 4
    corner = np.array([(0 << 61) | 1], np.uint64)          # FromFacePosLevel(0, 0, kMaxLevel)
    nb, cnt = cellid.vertex_neighbors(corner, 0)
    assert cnt[0] == 3 and nb[0][3] == 0
    assert sorted(int(x) for x in nb[0][:3]) == [(f << 61) | (1 << 60) for f in (0, 4, 5)]
    nb, cnt = cellid.all_neighbors(np.array([debug_id("3/0000")], np.uint64), 4)
    assert cnt[0] == 8
    assert sorted(int(x) for x in nb[0]) == sorted(debug_id(s) for s in
                                                   ["1/2221", "1/2222", "2/3330", "2/3333", "2/3333", "3/0001", "3/0002", "3/0003"])
    nb, cnt = cellid.all_neighbors(np.array([debug_id("3/")], np.uint64), 0)
    assert cnt[0] == 8 and {int(x) >> 61 for x in nb[0]} == {1, 2, 4, 5}   # s2cell_id_test.cc:677-690: never itself or face 0


def test_vertex_and_all_neighbors_match_reference(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200 import cellid
    from tests.test_encode_cpu import random_ids_all_levels
    ids = random_ids_all_levels(ref, 300_000, 23)
    lsb = ids & (~ids + np.uint64(1))
    level = 30 - (np.log2(lsb.astype(np.float64)).astype(np.int64) // 2)
    for target in (0, 3, 11, 29):
        sel = ids[level > target]
        want, wcnt = ref.vertex_neighbors(sel, target)
        got, gcnt = cellid.vertex_neighbors(sel, target)
        assert (gcnt == wcnt).all() and (got == want).all(), target
        d = torch.from_numpy(sel.view(np.int64)).cuda()
        got_d, gcnt_d = cellid.vertex_neighbors(d, target)
        assert (got_d.cpu().numpy().view(np.uint64) == want).all() and (gcnt_d.cpu().numpy() == wcnt).all()
    # ids at or above the target level violate the reference's precondition: count 0, zeros
    got, gcnt = cellid.vertex_neighbors(ids[level <= 11][:1000], 11)
    assert (gcnt == 0).all() and (got == 0).all()
    # all neighbours at the cell's own level (8 slots) and a few levels deeper (4 * 2^diff + 4 slots)
    for L in (0, 1, 7, 18, 30):
        sel = ids[level == L]
        if L == 0:
            sel = np.array([(f << 61) | (1 << 60) for f in range(6)], np.uint64)
        want, wcnt = ref.all_neighbors(sel, L, 8)
        got, gcnt = cellid.all_neighbors(sel, L)
        assert (gcnt == wcnt).all() and (got == want).all(), L
    for diff in (1, 3, 5):
        sel = ids[level <= 30 - diff][:20000]
        stride = 4 * 2 ** diff + 4
        for L in np.unique(level[level <= 30 - diff][:20000]):
            grp = sel[level[level <= 30 - diff][:20000] == L]
            want, wcnt = ref.all_neighbors(grp, int(L) + diff, stride)
            got, gcnt = cellid.all_neighbors(grp, int(L) + diff, stride)
            assert (gcnt == wcnt).all() and (wcnt == stride).all() and (got == want).all(), (diff, L)
    # a stride smaller than the count truncates, the count is still reported in full
    grp = ids[level == 10][:100]
    got, gcnt = cellid.all_neighbors(grp, 12, 5)
    want, wcnt = ref.all_neighbors(grp, 12, 5)
    assert (gcnt == 20).all() and (got == want).all()


def random_union_input(ref, n, seed, min_level=7):
    """Cell ids with duplicates, nested cells and complete sibling families (at several levels, cascading)."""
    from tests.test_encode_cpu import random_ids_all_levels
    rng = np.random.default_rng(seed)
    ids = random_ids_all_levels(ref, n, seed)
    lsb = ids & (~ids + np.uint64(1))
    ids = ids[lsb <= (np.uint64(1) << np.uint64(2 * (30 - min_level)))]     # a level-0 cell would swallow everything
    lsb = ids & (~ids + np.uint64(1))
    fam = ids[lsb > np.uint64(4)][: n // 8]
    flsb = fam & (~fam + np.uint64(1))
    kids = []
    for k in range(4):    # the four children of each cell in `fam` (child k = id + (2k + 1 - 4) * lsb/4)
        kids.append((fam.astype(np.int64) + (2 * k + 1 - 4) * (flsb >> np.uint64(2)).astype(np.int64)).astype(np.uint64))
    grand = []
    g0 = kids[0]; glsb = g0 & (~g0 + np.uint64(1))
    for k in range(4):    # child 0 is given only through ITS four children: merging must cascade
        grand.append((g0.astype(np.int64) + (2 * k + 1 - 4) * (glsb >> np.uint64(2)).astype(np.int64)).astype(np.uint64))
    parts = [ids[: n // 2], ids[: n // 10], kids[1], kids[2], kids[3]] + grand
    out = np.concatenate(parts)
    rng.shuffle(out)
    return out


def test_cell_union_normalize_matches_reference(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200 import cellid
    for n, seed in ((2000, 1), (200_000, 2), (1_500_000, 3)):
        ids = random_union_input(ref, n, seed)
        want = ref.cellunion_normalize(ids)
        assert 100 < want.size < ids.size // 2                       # the input really exercises dedup / nesting / merging
        got = cellid.cell_union_normalize(ids)
        assert got.shape == want.shape and (got == want).all(), n
        d = torch.from_numpy(ids.view(np.int64)).cuda()
        got_d = cellid.cell_union_normalize(d).cpu().numpy().view(np.uint64)
        assert (got_d == want).all()
        assert (cellid.cell_union_normalize(want) == want).all()     # idempotent
    # whole sphere at level 3 collapses to the six faces; single ids and the empty union pass through
    assert cellid.cell_union_normalize(np.zeros(0, np.uint64)).size == 0
    faces = np.array([(f << 61) | (1 << 60) for f in range(6)], np.uint64)
    level3 = faces
    for _ in range(3):
        lsb = level3 & (~level3 + np.uint64(1))
        level3 = np.concatenate([(level3.astype(np.int64) + (2 * k + 1 - 4) * (lsb >> np.uint64(2)).astype(np.int64)).astype(np.uint64) for k in range(4)])
    assert level3.size == 6 * 64
    assert (cellid.cell_union_normalize(level3) == faces).all() and (ref.cellunion_normalize(level3) == faces).all()
    assert (cellid.cell_union_normalize(faces[2:3]) == faces[2:3]).all()


def test_cell_union_contains_cells_matches_reference(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200 import cellid
    from tests.test_encode_cpu import random_ids_all_levels
    union = ref.cellunion_normalize(random_union_input(ref, 30_000, 7, min_level=8))
    assert union.size > 1000
    cap = ref.cover_cap(np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8]), 0.2, 1000)
    for u in (union, cap, cap[:1], np.zeros(0, np.uint64)):
        q = np.concatenate([random_ids_all_levels(ref, 400_000, 31), u, u[::3]])
        if u.size:   # parents and children of union cells: contains vs intersects differ exactly here
            lsb = u & (~u + np.uint64(1))
            ok = lsb < (np.uint64(1) << np.uint64(60))
            par = (u[ok] & (~(lsb[ok] << np.uint64(2)) + np.uint64(1))) | (lsb[ok] << np.uint64(2))
            ch = (u[lsb > 1].astype(np.int64) - (lsb[lsb > 1] >> np.uint64(2)).astype(np.int64)).astype(np.uint64)
            q = np.concatenate([q, par, ch])
        wc, wi = ref.cellunion_contains_cells(u, q)
        gc, gi = cellid.cell_union_contains_cells(u, q)
        assert (gc == wc).all() and (gi == wi).all()
        if u.size:
            assert wc.any() and (wi & ~wc).any()
            dc, di = cellid.cell_union_contains_cells(torch.from_numpy(u.view(np.int64)).cuda(), torch.from_numpy(q.view(np.int64)).cuda())
            assert (dc.cpu().numpy() == wc).all() and (di.cpu().numpy() == wi).all()


def test_to_lat_lng_within_ulps_of_reference(torch_cuda, ref):
    """S2CellId::ToLatLng. Floating-point output through atan2: CUDA's is documented <= 2 ulp, glibc's < 1 ulp, so the
    stated tolerance is 4 ulp (of the result, with pi/2 resp. pi as the scale near zero crossings of the other argument)."""
    torch = torch_cuda
    from s2geometry_b200 import cellid
    from tests.test_encode_cpu import random_ids_all_levels
    ids = random_ids_all_levels(ref, 1_000_000, 41)
    want = ref.cellid_to_latlng(ids)
    got = cellid.to_lat_lng(ids)
    ulp = np.spacing(np.maximum(np.abs(want), 1e-300))
    assert (np.abs(got - want) <= 4 * ulp).all()
    assert (got == want).mean() > 0.6                                  # and most are identical (73 % measured)
    assert (np.signbit(got) == np.signbit(want)).all()
    d = cellid.to_lat_lng(torch.from_numpy(ids.view(np.int64)).cuda()).cpu().numpy()
    assert (d == got).all()
    # round trip: the cell of the returned lat/lng at the id's level is the id itself
    lsb = ids & (~ids + np.uint64(1))
    leaf = cellid.from_lat_lng(got)
    assert (((leaf & (~lsb + np.uint64(1))) | lsb) == ids).mean() > 0.999   # (centres of coarse cells sit on leaf boundaries)


def test_cellid_info_matches_reference(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200 import cellid
    from tests.test_encode_cpu import random_ids_all_levels
    ids = random_ids_all_levels(ref, 500_000, 83)
    want_level = ref.level(ids)
    want_min, want_max = ref.id_range(ids)
    valid, level, rmin, rmax = cellid.info(ids)
    assert valid.all() and (level == want_level).all() and (rmin == want_min).all() and (rmax == want_max).all()
    bad = np.array([0, 7 << 61, (6 << 61) | 1, 2, 8, (1 << 61) | 2, 0xFFFFFFFFFFFFFFFF], np.uint64)   # None, face >= 6, even lsb position
    v2, l2, _, _ = cellid.info(bad)
    assert (~v2).all() and (l2 == -1).all()
    dv, dl, da, db = cellid.info(torch.from_numpy(ids.view(np.int64)).cuda())
    assert dv.all().item() and (dl.cpu().numpy() == level).all() and (da.cpu().numpy().view(np.uint64) == rmin).all() and (db.cpu().numpy().view(np.uint64) == rmax).all()


def test_children_advance_and_common_ancestor_level(torch_cuda, ref):
    """S2CellId::child, advance, advance_wrap, next / prev (+ _wrap) and GetCommonAncestorLevel element-wise against the
    reference: cells of every level incl. the first and last cell of each level (where advance clamps and the wrapping
    forms wrap), steps from 0 to far beyond the level's size, pairs on the same and on different faces."""
    import ctypes
    from oracle.reflib import _p, _sz
    torch = torch_cuda
    from s2geometry_b200 import cellid
    rng = np.random.default_rng(5)
    leaf = ref.encode_points(rng.standard_normal((60000, 3)), 30, 1)
    lv = rng.integers(0, 31, len(leaf))
    ids = np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(lv)], np.uint64)
    ends = []
    for level in range(31):
        lsb = np.uint64(1) << np.uint64(2 * (30 - level))
        ends += [lsb, np.uint64(6 << 61) - lsb, lsb + (lsb << np.uint64(1)), np.uint64(3 << 61) + lsb]   # Begin, last, second, first of face 3
    ids = np.concatenate([ids, np.array(ends, np.uint64)])
    n = len(ids)
    lib = ref.lib
    # children (non-leaf cells)
    par = ids[(ids & np.uint64(1)) == 0]
    want = np.empty((len(par), 4), np.uint64)
    lib.ref_cellid_children(_p(par), _sz(len(par)), _p(want))
    assert (cellid.children(par) == want).all()
    assert (cellid.children(torch.from_numpy(par.view(np.int64)).cuda()).cpu().numpy().view(np.uint64) == want).all()
    assert (cellid.children(ids[(ids & np.uint64(1)) == 1][:100]) == 0).all()          # leaf cells have no children
    # advance / advance_wrap
    mag = 10.0 ** rng.uniform(0, 18.9, n)
    steps = (rng.choice([-1, 1], n) * mag).astype(np.int64)
    steps[::11] = rng.integers(-3, 4, len(steps[::11]))
    steps[::97] = 0
    for wrap in (0, 1):
        want = np.empty(n, np.uint64)
        lib.ref_cellid_advance(_p(ids), _p(steps), _sz(n), wrap, _p(want))
        assert (cellid.advance(ids, steps, bool(wrap)) == want).all(), wrap
        got = cellid.advance(torch.from_numpy(ids.view(np.int64)).cuda(), torch.from_numpy(steps).cuda(), bool(wrap))
        assert (got.cpu().numpy().view(np.uint64) == want).all(), wrap
    nx, pv, nxw, pvw = (np.empty(n, np.uint64) for _ in range(4))
    lib.ref_cellid_next_prev(_p(ids), _sz(n), _p(nx), _p(pv), _p(nxw), _p(pvw))
    # next() / prev() do not clamp (they may return End(level) / an id before Begin); advance(+-1) equals them except there
    inner = (ids != np.array([np.uint64(6 << 61) - (i & (~i + np.uint64(1))) for i in ids], np.uint64))
    assert (cellid.advance(ids, 1)[inner] == nx[inner]).all()
    assert (cellid.advance(ids, 1, wrap=True) == nxw).all() and (cellid.advance(ids, -1, wrap=True) == pvw).all()
    first = ids == (ids & (~ids + np.uint64(1)))
    assert (cellid.advance(ids, -1)[~first] == pv[~first]).all()
    # common ancestor level
    other = np.concatenate([ids[rng.permutation(n)[: n // 2]], ids[n // 2:]])                              # unrelated cells, and each cell with itself
    anc = rng.integers(0, 31, n)
    near = np.array([ref.parent(leaf[i % len(leaf)][None], int(min(a, 30)))[0] for i, a in enumerate(anc)], np.uint64)   # relatives of the same leaves
    for b in (other, near):
        want = np.empty(n, np.int8)
        lib.ref_cellid_common_ancestor_level(_p(ids), _p(b), _sz(n), _p(want))
        assert (cellid.common_ancestor_level(ids, b) == want).all()
        assert (cellid.common_ancestor_level(torch.from_numpy(ids.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()).cpu().numpy() == want).all()
        assert (want >= 0).sum() > 1000
