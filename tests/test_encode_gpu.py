"""GPU parity tests for K1 (cell-id encoding), through the C ABI (host-pointer path and device-pointer path)."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

KA = util.known_answers()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def test_known_answers(torch_cuda):
    from s2geometry_b200 import cellid
    for row in KA["point_to_id"]:
        assert cellid.from_points(np.array([row["xyz"]]))[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_id"]:
        assert cellid.from_lat_lng(np.radians(np.array([[row["lat"], row["lng"]]])))[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_face"]:
        ll = np.array([[(np.pi / 180) * row["lat"], (np.pi / 180) * row["lng"]]])
        assert int(cellid.from_lat_lng(ll)[0]) >> 61 == row["face"]
    ids = np.array([int(r["id"], 16) for r in KA["tokens"]], np.uint64)
    assert cellid.tokens_to_str(*cellid.to_token(ids)) == [r["token"] for r in KA["tokens"]]
    for row in KA["latlng_deg_in_cells"]:
        leaf = int(cellid.from_lat_lng(np.radians(np.array([[row["lat"], row["lng"]]])))[0])
        cells = [int(c, 16) for c in row["cells"]]
        assert any(c - ((c & -c) - 1) <= leaf <= c + ((c & -c) - 1) for c in cells)


def test_golden_fixture_host_path(torch_cuda):
    from s2geometry_b200 import cellid
    fx = util.fixture("encode_fixture.npz")
    xyz, ll, e7 = fx["xyz"], fx["latlng"], fx["e7"]
    for lvl in (30, 20, 12, 0):
        assert (cellid.from_points(xyz, lvl) == fx["xyz_ids_L%d" % lvl]).all()
        assert (cellid.from_lat_lng(ll, lvl) == fx["latlng_ids_L%d" % lvl]).all()
    assert (cellid.from_lat_lng_e7(e7) == fx["e7_ids_L30"]).all()
    multi = cellid.from_lat_lng_levels(ll, [12, 20, 30])
    for k, lvl in enumerate((12, 20, 30)):
        assert (multi[k] == fx["latlng_ids_L%d" % lvl]).all()
    ids = fx["tok_ids"]
    tok, ln = cellid.to_token(ids)
    assert (tok == fx["tok"]).all() and (ln == fx["tok_len"]).all()
    assert (cellid.parent(fx["xyz_ids_L30"], 12) == fx["xyz_ids_L12"]).all()


def test_golden_fixture_device_path(torch_cuda):
    torch = torch_cuda
    from s2geometry_b200 import cellid
    fx = util.fixture("encode_fixture.npz")
    xyz = torch.from_numpy(fx["xyz"]).cuda()
    ll = torch.from_numpy(fx["latlng"]).cuda()
    for lvl in (30, 12):
        got = cellid.from_points(xyz, lvl).cpu().numpy().view(np.uint64)
        assert (got == fx["xyz_ids_L%d" % lvl]).all()
        got = cellid.from_lat_lng(ll, lvl).cpu().numpy().view(np.uint64)
        assert (got == fx["latlng_ids_L%d" % lvl]).all()
    multi = cellid.from_points_levels(xyz, [0, 12, 20, 30]).cpu().numpy().view(np.uint64)
    for k, lvl in enumerate((0, 12, 20, 30)):
        assert (multi[k] == fx["xyz_ids_L%d" % lvl]).all()
    ids = torch.from_numpy(fx["tok_ids"].view(np.int64)).cuda()
    tok, ln = cellid.to_token(ids)
    assert (tok.cpu().numpy() == fx["tok"]).all() and (ln.cpu().numpy() == fx["tok_len"]).all()
    e7 = torch.from_numpy(fx["e7"]).cuda()
    assert (cellid.from_lat_lng_e7(e7).cpu().numpy().view(np.uint64) == fx["e7_ids_L30"]).all()


def test_empty_and_ragged_sizes(torch_cuda):
    from s2geometry_b200 import cellid, synth
    assert cellid.from_points(np.zeros((0, 3))).shape == (0,)
    p = synth.sphere_points(100003, 5)
    full = cellid.from_points(p)
    for n in (1, 31, 32, 33, 255, 257, 100003):
        assert (cellid.from_points(p[:n]) == full[:n]).all()


def test_config1_10M_points_vs_reference(torch_cuda, ref):
    """BASELINE config 1: 10M random S2Points -> level-30 ids, every id compared with the reference's CPU path."""
    from s2geometry_b200 import cellid, synth
    p = synth.sphere_points(10_000_000, 1)
    got = cellid.from_points(p)          # host path: chunked pipeline (3 chunks)
    want = ref.encode_points(p, 30)
    assert (got == want).all()


def test_points_on_leaf_cell_boundaries_vs_reference(torch_cuda, ref):
    """S2Point encoding where it is most sensitive: 4M points on / within 1e-8..1e-4 leaf cells of cell boundaries (unit
    and unnormalised), plus scaled and tiny-component vectors. Production path (bounded-error filter + IEEE sequence
    where it matters) == reference, at levels 30 and 7, and == the all-IEEE diagnostic mode."""
    from s2geometry_b200 import _lib, cellid, synth
    lib = _lib.load()
    rng = np.random.default_rng(8)
    u = synth.sphere_points(1_000_000, 3)
    p = np.concatenate([util.leaf_boundary_points(4_000_000, 2), u * 10.0 ** rng.uniform(-200, 200, (len(u), 1)),
                        u * np.array([1.0, 1e-170, 1e-300]), u * np.array([1e-310, 1.0, 1e-20])])
    for level in (30, 7):
        got = cellid.from_points(p, level)
        assert (got == ref.encode_points(p, level)).all()
    try:
        lib.s2b_diag_force_correctly_rounded(1)
        exact = cellid.from_points(p, 30)
    finally:
        lib.s2b_diag_force_correctly_rounded(0)
    assert (exact == cellid.from_points(p, 30)).all()


def test_latlng_5M_vs_reference_and_flags(torch_cuda, ref):
    """Default (S2B_LATLNG_HOST_LIBM) mode: lat/lng ids equal the reference's, evaluated with this process's libm, bit for
    bit at every level. In S2B_LATLNG_CORRECTLY_ROUNDED mode the ids are those of correctly rounded sin/cos; glibc is not
    correctly rounded, so a mismatch is possible only where the boundary flag is set."""
    from s2geometry_b200 import cellid, synth
    ll = synth.sphere_latlng(5_000_000, 2)
    cellid.lat_lng_strict_stats(reset=True)
    got, flags = cellid.from_lat_lng(ll, return_boundary_flags=True)
    want = ref.encode_latlng(ll, 30)
    assert (got != want).sum() == 0
    reevaluated, changed = cellid.lat_lng_strict_stats()
    assert reevaluated == int(flags.sum()) and 0 < reevaluated < 1e-4 * len(ll)
    for lvl in (12, 20):
        assert (cellid.from_lat_lng(ll, lvl) != ref.encode_latlng(ll, lvl)).sum() == 0
    old = cellid.set_lat_lng_mode(cellid.LATLNG_CORRECTLY_ROUNDED)
    try:
        cr, flags_cr = cellid.from_lat_lng(ll, return_boundary_flags=True)
    finally:
        cellid.set_lat_lng_mode(old)
    bad = cr != want
    assert bad.sum() <= 2 and (flags_cr[bad] == 1).all() and (flags_cr == flags).all()
    assert (cr != got).sum() == bad.sum()


def test_config2_100M_latlng_strict_vs_reference(torch_cuda, ref):
    """BASELINE configs[1] at full size: 100M lat/lng -> ids at levels 12/20/30 + ToToken(level 30), EVERY id, token and
    token length compared with the reference's S2CellId(S2LatLng) / parent / ToToken on the host: 0 mismatches, through the
    host-pointer call and through the device-pointer strict call."""
    torch = torch_cuda
    from s2geometry_b200 import cellid, synth
    n = 100_000_000
    ll = synth.sphere_latlng(n, 2)
    w_ids, w_tok, w_len = ref.encode_latlng_levels_tokens(ll, [12, 20, 30], 2, ref.hardware_threads())
    cellid.lat_lng_strict_stats(reset=True)
    ids, tok, ln = cellid.from_lat_lng_levels_tokens(ll, [12, 20, 30], 2)
    reevaluated, changed = cellid.lat_lng_strict_stats(reset=True)
    print("strict mode, 100M points: %d re-evaluated with the host libm, %d ids changed" % (reevaluated, changed))
    for k in range(3):
        assert int((ids[k] != w_ids[k]).sum()) == 0
    assert int((tok != w_tok).sum()) == 0 and int((ln != w_len).sum()) == 0
    assert 0 < reevaluated < 1e-4 * n
    del ids, tok, ln
    d_ids, d_tok, d_len = cellid.from_lat_lng_levels_tokens(torch.from_numpy(ll).cuda(), [12, 20, 30], 2)
    assert cellid.lat_lng_strict_stats()[0] == reevaluated
    for k in range(3):
        assert bool((d_ids[k].cpu() == torch.from_numpy(w_ids[k].view(np.int64))).all())
    assert bool((d_tok.cpu() == torch.from_numpy(w_tok)).all()) and bool((d_len.cpu() == torch.from_numpy(w_len)).all())


def test_latlng_strict_on_cell_boundaries_and_odd_angles(torch_cuda, ref):
    """The inputs where the libm matters most, all through the strict mode: 400k points ON leaf / face boundaries and a
    few ulps off them (more sensitive points than the per-chunk list holds: the piecewise fallback runs), E7 input,
    angles far outside S2LatLng::is_valid (|x| up to 1e300), infinities and NaN. Ids == the reference's in this process."""
    torch = torch_cuda
    from s2geometry_b200 import cellid
    from tests.test_encode_cpu import boundary_latlngs
    rng = np.random.default_rng(17)
    bl = boundary_latlngs(ref, 20000, 10)
    odd = np.concatenate([rng.uniform(-1e6, 1e6, (50000, 2)), rng.uniform(-4e6, 4e6, (50000, 2)),
                          10.0 ** rng.uniform(6, 300, (20000, 2)) * rng.choice([-1.0, 1.0], (20000, 2)),
                          np.array([[np.inf, 0.3], [0.3, -np.inf], [np.nan, 1.0], [1.0, np.nan], [0.0, 0.0], [-0.0, np.pi],
                                    [np.pi / 2, 0.0], [-np.pi / 2, 2.0], [0.0, np.pi / 2], [0.0, -np.pi / 2], [0.0, np.pi / 4]])])
    for ll in (bl, odd, np.tile(bl, (12, 1))):
        want = ref.encode_latlng(ll, 30)
        cellid.lat_lng_strict_stats(reset=True)
        assert (cellid.from_lat_lng(ll, 30) != want).sum() == 0                                          # host-pointer path
        host_stats = cellid.lat_lng_strict_stats(reset=True)
        assert (cellid.from_lat_lng(torch.from_numpy(ll).cuda(), 30).cpu().numpy().view(np.uint64) != want).sum() == 0   # device path
        assert cellid.lat_lng_strict_stats() == host_stats and host_stats[0] > 0.1 * len(bl)
        w3, wt, wl = ref.encode_latlng_levels_tokens(ll, [5, 17, 30], 1, 0)
        for src in (ll, torch.from_numpy(ll).cuda()):
            g3, gt, gl = cellid.from_lat_lng_levels_tokens(src, [5, 17, 30], 1)
            if not isinstance(g3, np.ndarray):
                g3, gt, gl = g3.cpu().numpy().view(np.uint64), gt.cpu().numpy(), gl.cpu().numpy()
            assert (g3 != w3).sum() == 0 and (gt != wt).sum() == 0 and (gl != wl).sum() == 0
    e7 = np.concatenate([np.round(np.degrees(bl[:100000]) * 1e7).astype(np.int32),
                         rng.integers(-1800000000, 1800000001, (400000, 2)).astype(np.int32),
                         np.array([[0, 0], [900000000, 0], [-900000000, 1800000000], [450000000, 450000000], [0, 900000000],
                                   [0, -900000000], [353000000, 450000000]], np.int32)])
    want = ref.encode_latlng_e7(e7, 30)
    assert (cellid.from_lat_lng_e7(e7) != want).sum() == 0
    assert (cellid.from_lat_lng_e7(torch.from_numpy(e7).cuda()).cpu().numpy().view(np.uint64) != want).sum() == 0


def test_full_size_properties(torch_cuda):
    """100M lat/lng on the device: size-independent properties (BASELINE config 2 size).
    parent(L) of the level-30 id equals the fused multi-level output; ids are valid leaf ids; token round trip."""
    torch = torch_cuda
    from s2geometry_b200 import cellid, synth
    n = 100_000_000
    ll = synth.sphere_latlng_cuda(n, 2, "cuda")
    multi = cellid.from_lat_lng_levels(ll, [12, 20, 30])
    leaf = multi[2]
    assert bool(((leaf & 1) == 1).all())                      # leaf cells have lsb 1
    assert bool((((leaf >> 61) & 7) < 6).all())               # valid face
    assert bool((cellid.parent(leaf, 12) == multi[0]).all())
    assert bool((cellid.parent(leaf, 20) == multi[1]).all())
    single = cellid.from_lat_lng(ll, 30)
    assert bool((single == leaf).all())
    # uniform points: face histogram is balanced within 1%
    faces = torch.bincount(((leaf >> 61) & 7).to(torch.int64), minlength=6).double() / n
    assert float((faces - 1 / 6).abs().max()) < 0.002
    # tokens of level-12 ids have exactly 7 hex digits (3 face bits + 24 position bits + marker -> 28 bits)
    tok, ln = cellid.to_token(multi[0][:10_000_000])
    assert bool((ln == 7).all())
    del multi, leaf, single
    # a sample compared id by id with the host path (different launch geometry, same kernel)
    sample = ll[:1_000_000].cpu().numpy()
    assert (cellid.from_lat_lng(sample).view(np.int64) == cellid.from_lat_lng(ll[:1_000_000]).cpu().numpy()).all()


def test_fast_filter_equals_correctly_rounded_on_device(torch_cuda, ref):
    """The production lat/lng path (fast sin/cos + correctly rounded re-evaluation of boundary-sensitive points) gives
    exactly the ids of the all-correctly-rounded path: 100M uniform points + 400k points sitting on cell boundaries."""
    torch = torch_cuda
    from s2geometry_b200 import _lib, cellid, synth
    from tests.test_encode_cpu import boundary_latlngs
    lib = _lib.load()
    ll = synth.sphere_latlng_cuda(100_000_000, 5, "cuda")
    bl = torch.from_numpy(boundary_latlngs(ref, 20000, 10)).cuda()
    old_mode = cellid.set_lat_lng_mode(cellid.LATLNG_CORRECTLY_ROUNDED)   # this test is about the device's own definition
    try:
        fast = cellid.from_lat_lng(ll, 30)
        fast_b, flags_b = cellid.from_lat_lng(bl, 30, return_boundary_flags=True)
        assert lib.s2b_diag_force_correctly_rounded(1) == 0
        exact = cellid.from_lat_lng(ll, 30)
        exact_b, flags_b2 = cellid.from_lat_lng(bl, 30, return_boundary_flags=True)
    finally:
        lib.s2b_diag_force_correctly_rounded(0)
        cellid.set_lat_lng_mode(old_mode)
    assert bool((fast == exact).all())
    assert bool((fast_b == exact_b).all()) and bool((flags_b == flags_b2).all())
    assert float(flags_b.float().mean()) > 0.1
    # host reference (glibc libm): differences only where the flag is set
    want = ref.encode_latlng(bl.cpu().numpy(), 30)
    got = fast_b.cpu().numpy().view(np.uint64)
    bad = got != want
    assert (flags_b.cpu().numpy()[bad] == 1).all()


def test_filter_deviation_is_far_below_the_filter_tolerance(torch_cuda, ref):
    """The filter path (fast sin/cos, approximate reciprocal / square root) must stay within a fraction of kFilterTolIj
    (1.6e-5 leaf cells) of the correctly rounded path, and never pick another face unnoticed: 200M uniform points, points
    on cell boundaries, tiny and huge (but in-domain) angles."""
    import ctypes
    torch = torch_cuda
    from s2geometry_b200 import _lib, synth
    from tests.test_encode_cpu import boundary_latlngs
    lib = _lib.load()
    rng = np.random.default_rng(3)
    special = np.concatenate([boundary_latlngs(ref, 20000, 11),
                              np.stack([rng.uniform(-1.6, 1.6, 200000), rng.uniform(-3.2, 3.2, 200000)], 1),
                              np.stack([10.0 ** rng.uniform(-300, 0, 100000), 10.0 ** rng.uniform(-300, 0, 100000)], 1),
                              rng.uniform(-1e6, 1e6, (100000, 2))])
    worst = 0.0
    for ll in (synth.sphere_latlng_cuda(200_000_000, 9, "cuda"), torch.from_numpy(np.ascontiguousarray(special)).cuda()):
        dev, flips = ctypes.c_double(0), ctypes.c_uint64(0)
        rc = lib.s2b_diag_filter_deviation_dev(1, ll.data_ptr(), ll.shape[0], ctypes.byref(dev), ctypes.byref(flips), None)
        assert rc == 0 and flips.value == 0
        worst = max(worst, dev.value)
    print("max filter deviation, lat/lng [leaf cells]:", worst)
    assert worst < 4e-6, worst          # kFilterTolIj / 4
    # S2Point input: approximate vs IEEE division / square root (kXyzFilterTolIj = 8e-6). Unit vectors, scaled vectors,
    # points on cell boundaries and axis-aligned / tiny-component points.
    u = synth.sphere_points_cuda(200_000_000, 4, "cuda")
    edge = torch.from_numpy(util.leaf_boundary_points(2_000_000, 5)).cuda()
    scaled = u[:1_000_000] * torch.from_numpy(10.0 ** rng.uniform(-140, 140, (1_000_000, 1))).cuda()
    tiny = u[:1_000_000].clone()
    tiny[:, 1] *= 1e-200
    worst = 0.0
    for pts in (u, edge, scaled.contiguous(), tiny):
        dev, flips = ctypes.c_double(0), ctypes.c_uint64(0)
        rc = lib.s2b_diag_filter_deviation_dev(0, pts.data_ptr(), pts.shape[0], ctypes.byref(dev), ctypes.byref(flips), None)
        assert rc == 0 and flips.value == 0
        worst = max(worst, dev.value)
    print("max filter deviation, xyz [leaf cells]:", worst)
    assert worst < 2e-6, worst          # kXyzFilterTolIj / 4


def test_to_point_and_from_token(torch_cuda, ref):
    """Batched inverses: S2CellId::ToPoint bit-exact, FromToken, token round trip."""
    torch = torch_cuda
    from s2geometry_b200 import cellid, synth
    rng = np.random.default_rng(8)
    leaf = cellid.from_points(synth.sphere_points(2_000_000, 21))
    lv = rng.integers(0, 31, len(leaf))
    lsb = np.uint64(1) << (2 * (30 - lv)).astype(np.uint64)
    ids = (leaf & (~lsb + np.uint64(1))) | lsb                      # parent(level) with a per-id level
    want = ref.to_point(ids)
    assert (cellid.to_point(ids) == want).all()
    d_ids = torch.from_numpy(ids.view(np.int64)).cuda()
    assert (cellid.to_point(d_ids).cpu().numpy() == want).all()
    # leaf cell of the centre of a cell is inside that cell (S2CellId.Inverses property, s2cell_id_test.cc:327-338)
    back = cellid.from_points(want)
    lo = ids - (lsb - np.uint64(1)); hi = ids + (lsb - np.uint64(1))
    assert ((lo <= back) & (back <= hi)).all()
    tok, ln = cellid.to_token(ids)
    assert (cellid.from_token(tok, ln) == ids).all()
    assert (cellid.from_token(torch.from_numpy(tok).cuda(), torch.from_numpy(ln).cuda()).cpu().numpy().view(np.uint64) == ids).all()
    assert (ref.from_token(tok, ln) == ids).all()


def test_cuda_vs_independent_restatement(torch_cuda):
    """Second, independent oracle (oracle/restatement.py: NumPy + exact rationals; no code shared with the kernels or with
    the compiled reference): cell ids on 2M points, and containment with vertex models on a small polygon."""
    from oracle import restatement as rs
    from s2geometry_b200 import cellid, synth
    from s2geometry_b200.index import ContainsPointQuery, ShapeIndex, build_flat_index
    from s2geometry_b200.shapes import ShapeSoup
    p = synth.sphere_points(2_000_000, 81)
    assert (cellid.from_points(p) == rs.encode_points(p)).all()
    assert (cellid.from_points(p, 7) == rs.encode_points(p, 7)).all()
    ids = cellid.from_points(p[:20000], 11)
    assert cellid.tokens_to_str(*cellid.to_token(ids)) == rs.to_token(ids)
    assert (cellid.to_point(ids) == rs.to_point(ids)).all()
    loop = synth.wobbly_loop(np.array([0.3, -0.5, 0.8]), 0.17, 40)
    soup = ShapeSoup(); soup.add_polygon([loop])
    ix = ShapeIndex(build_flat_index(soup.freeze()))
    pts = np.concatenate([synth.cap_points(np.array([0.3, -0.5, 0.8]), 0.22, 120, 4), synth.degenerate_points(loop)])
    for m in (0, 1, 2):
        want = np.array([rs.polygon_contains([loop], q, m) for q in pts], np.uint8)
        assert (ContainsPointQuery(ix, m).contains(pts) == want).all(), m


def test_edge_neighbors_on_gpu(torch_cuda, ref):
    torch = torch_cuda
    from s2geometry_b200 import cellid
    from tests.test_encode_cpu import random_ids_all_levels
    ids = random_ids_all_levels(ref, 1_000_000, 19)
    want = ref.edge_neighbors(ids)
    assert (cellid.edge_neighbors(ids) == want).all()
    d = torch.from_numpy(ids.view(np.int64)).cuda()
    assert (cellid.edge_neighbors(d).cpu().numpy().view(np.uint64) == want).all()


def test_synthetic_streams_device_equals_host(torch_cuda):
    """The counter-based generators (bench.py's inputs) give the same bits on the device as the NumPy restatement, for
    arbitrary slices of the stream: what makes the multi-GPU results comparable across N and with the reference arm."""
    from s2geometry_b200 import synth
    c = np.array([0.3, -0.5, 0.8])
    for seed, first, n in ((2, 0, 300_000), (1000, 999_999_000, 100_000), (2**63 + 11, 2**45, 4097)):
        assert np.array_equal(synth.stream_sphere_points_cuda(seed, first, n, "cuda").cpu().numpy(), synth.stream_sphere_points(seed, first, n))
        assert np.array_equal(synth.stream_sphere_latlng_cuda(seed, first, n, "cuda").cpu().numpy(), synth.stream_sphere_latlng(seed, first, n))
        assert np.array_equal(synth.stream_cap_points_cuda(seed, first, n, c, 0.2125, "cuda").cpu().numpy(), synth.stream_cap_points(seed, first, n, c, 0.2125))
    assert synth.stream_sphere_points_cuda(1, 0, 0, "cuda").shape == (0, 3)
