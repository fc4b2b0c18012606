"""CPU-side checks for the encode path (no GPU needed):
 - the reference oracle (oracle/_ref) reproduces the reference's own known answers  -> the oracle is pinned;
 - the committed golden fixture equals what the oracle produces now                  -> the fixture is current;
 - the device headers, compiled for the host (tests/hostsim), reproduce both         -> kernel logic is right before
   any GPU time is spent.
"""
import ctypes

import numpy as np
import pytest

from tests import util
from tests.util import vp

KA = util.known_answers()


def test_oracle_known_answers(ref):
    for row in KA["point_to_id"]:
        assert ref.encode_points(np.array([row["xyz"]]), 30, 1)[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_id"]:
        assert ref.encode_latlng_deg(np.array([[row["lat"], row["lng"]]]), 30, 1)[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_face"]:
        assert int(ref.encode_latlng_deg(np.array([[row["lat"], row["lng"]]], float), 30, 1)[0]) >> 61 == row["face"]
    for row in KA["latlng_deg_in_cells"]:
        leaf = ref.encode_latlng_deg(np.array([[row["lat"], row["lng"]]]), 30, 1)
        lo, hi = ref.id_range(np.array([int(c, 16) for c in row["cells"]], np.uint64))
        assert ((lo <= leaf[0]) & (leaf[0] <= hi)).any()
    for row in KA["face_pos_level"]:
        assert ref.from_face_pos_level(row["face"], int(row["pos"], 16), row["level"]) == int(row["id"], 16)
    ids = np.array([int(r["id"], 16) for r in KA["tokens"]], np.uint64)
    assert util.tok_str(*ref.to_token(ids)) == [r["token"] for r in KA["tokens"]]


def test_fixture_is_current(ref):
    fx = util.fixture("encode_fixture.npz")
    for lvl in (30, 20, 12, 0):
        assert (ref.encode_points(fx["xyz"], lvl, 1) == fx["xyz_ids_L%d" % lvl]).all()
        assert (ref.encode_latlng(fx["latlng"], lvl, 1) == fx["latlng_ids_L%d" % lvl]).all()
    assert (ref.encode_latlng_e7(fx["e7"], 30, 1) == fx["e7_ids_L30"]).all()


def test_hostsim_known_answers(hostsim):
    for row in KA["point_to_id"]:
        assert util.hs_encode_points(hostsim, np.array([row["xyz"]]))[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_id"]:
        ll = np.radians(np.array([[row["lat"], row["lng"]]]))
        assert util.hs_encode_latlng(hostsim, ll)[0][0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_face"]:
        ll = np.array([[(np.pi / 180) * row["lat"], (np.pi / 180) * row["lng"]]])
        assert int(util.hs_encode_latlng(hostsim, ll)[0][0]) >> 61 == row["face"]
    ids = np.array([int(r["id"], 16) for r in KA["tokens"]], np.uint64)
    tok = np.zeros((len(ids), 16), np.uint8); ln = np.zeros(len(ids), np.uint8)
    hostsim.hs_cellid_to_token(vp(ids), ctypes.c_size_t(len(ids)), vp(tok), vp(ln))
    assert util.tok_str(tok, ln) == [r["token"] for r in KA["tokens"]]


def test_hostsim_fixture(hostsim):
    fx = util.fixture("encode_fixture.npz")
    for lvl in (30, 20, 12, 0):
        assert (util.hs_encode_points(hostsim, fx["xyz"], lvl) == fx["xyz_ids_L%d" % lvl]).all()
        got, _ = util.hs_encode_latlng(hostsim, fx["latlng"], lvl)
        assert (got == fx["latlng_ids_L%d" % lvl]).all()
    ids = fx["tok_ids"]  # keep a reference: NpzFile returns a fresh array per access
    tok = np.zeros((len(ids), 16), np.uint8); ln = np.zeros(len(ids), np.uint8)
    hostsim.hs_cellid_to_token(vp(ids), ctypes.c_size_t(len(ln)), vp(tok), vp(ln))
    assert (tok == fx["tok"]).all() and (ln == fx["tok_len"]).all()


def test_hostsim_vs_reference_random(hostsim, ref):
    from s2geometry_b200 import synth
    p = synth.sphere_points(1_000_000, 7)
    assert (util.hs_encode_points(hostsim, p) == ref.encode_points(p, 30)).all()
    ll = synth.sphere_latlng(1_000_000, 8)
    got, n_slow = util.hs_encode_latlng(hostsim, ll)
    want = ref.encode_latlng(ll, 30)
    # lat/lng ids depend on libm: ours is correctly rounded, glibc is not; expected ~1e-3 mismatches per 1M.
    assert (got != want).sum() <= 1
    assert n_slow < 2000  # the double-double stage is rare (~2e-4 per sincos)


def test_sincos_correctly_rounded(hostsim):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 300
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 20000), rng.uniform(-50, 50, 3000), 10.0 ** rng.uniform(-30, 6, 1000),
                        [np.pi / 2, -np.pi / 2, np.pi, -np.pi, np.pi / 4, 0.0, -0.0, 1 / 128, 3 / 256, 5e-324, 1048575.9]])
    for mode in (0, 1):  # production (Ziv) and forced double-double stage
        s = np.empty_like(x); c = np.empty_like(x)
        hostsim.hs_sincos(vp(x), ctypes.c_size_t(len(x)), mode, vp(s), vp(c), None)
        for i in range(len(x)):
            assert float(mp.sin(mp.mpf(x[i]))) == s[i] and float(mp.cos(mp.mpf(x[i]))) == c[i], (mode, x[i].hex())


def test_fast_sincos_error_bound(hostsim):
    """The filter in front of the correctly rounded sin/cos assumes |error| <= 4.5e-16 (kFastSinCosAbsErr)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    rng = np.random.default_rng(12)
    x = np.concatenate([rng.uniform(-np.pi, np.pi, 40000), rng.uniform(-2000, 2000, 5000), rng.uniform(-1e6, 1e6, 3000),
                        np.pi / 2 * np.arange(-8, 9), np.pi / 4 * np.arange(-9, 10) * (1 + 1e-16), [0.0, 1e-300, -1e-10]])
    # arguments next to the table nodes k*pi/256 and half way between them (largest |d|), over the whole supported range
    k = rng.integers(-600, 600, 6000)
    x = np.concatenate([x, k * (np.pi / 256) * (1 + rng.uniform(-2e-16, 2e-16, len(k))), (k + 0.5) * (np.pi / 256),
                        (rng.integers(-2**27, 2**27, 4000) + rng.choice([0.0, 0.5, 0.4999999], 4000)) * (np.pi / 256)])
    x = x[np.abs(x) < 2.0 ** 20]
    for fn, limit in ((hostsim.hs_sincos_fast, 2.3e-16), (hostsim.hs_sincos_tab, 1.3e-16)):
        s = np.empty_like(x); c = np.empty_like(x)
        fn(vp(x), ctypes.c_size_t(len(x)), vp(s), vp(c))
        worst = 0.0
        for i in range(len(x)):
            worst = max(worst, abs(float(mp.sin(mp.mpf(x[i])) - mp.mpf(s[i]))), abs(float(mp.cos(mp.mpf(x[i])) - mp.mpf(c[i]))))
        assert worst < limit, worst     # the encoder budgets 4.5e-16 (kFastSinCosAbsErr); the table routine is the one in use


def boundary_latlngs(ref, n, seed):
    """lat/lng pairs whose points sit on or within a few ulps of leaf-cell boundaries, face boundaries and cube corners."""
    rng = np.random.default_rng(seed)
    # vertices / edge points of cells at several levels: one or both of (i, j) integral at that level
    out = []
    for lvl in (30, 30, 24, 12, 3):
        step = 1 << (30 - lvl)
        i = rng.integers(0, (1 << lvl) + 1, n) * step
        j = rng.integers(0, (1 << lvl) + 1, n) * step
        frac = rng.uniform(0, step, n)
        which = rng.integers(0, 3, n)                 # 0: vertex, 1: i on a boundary, 2: j on a boundary
        fi = np.where(which == 2, np.minimum(i + frac, 2.0 ** 30), i).astype(np.float64)
        fj = np.where(which == 1, np.minimum(j + frac, 2.0 ** 30), j).astype(np.float64)
        s_, t_ = fi / 2.0 ** 30, fj / 2.0 ** 30
        st2uv = lambda s: np.where(s >= 0.5, (1 / 3.) * (4 * s * s - 1), (1 / 3.) * (1 - 4 * (1 - s) * (1 - s)))
        u, v = st2uv(s_), st2uv(t_)
        face = rng.integers(0, 6, n)
        one = np.ones(n)
        xyz = np.select([(face == k)[:, None] for k in range(6)],
                        [np.stack(c, 1) for c in ((one, u, v), (-u, one, v), (-u, -v, one), (-one, -v, -u), (v, -one, -u), (v, u, -one))])
        xyz /= np.sqrt((xyz * xyz).sum(1))[:, None]
        out.append(np.stack([np.arcsin(np.clip(xyz[:, 2], -1, 1)), np.arctan2(xyz[:, 1], xyz[:, 0])], axis=1))
    ll = np.vstack(out)
    # nudge by a few ulps in both coordinates
    k = rng.integers(-3, 4, ll.shape)
    ll2 = ll.copy()
    for _ in range(3):
        ll2 = np.where(k > 0, np.nextafter(ll2, np.inf), np.where(k < 0, np.nextafter(ll2, -np.inf), ll2))
        k = k - np.sign(k)
    # face boundaries: |x| = |y| (lng = +-45, +-135 deg) and cube-corner latitudes
    t = rng.uniform(-1.2, 1.2, 4000)
    fb = np.stack([t, rng.choice([np.pi / 4, -np.pi / 4, 3 * np.pi / 4, -3 * np.pi / 4], 4000)], axis=1)
    c35 = np.arctan(1 / np.sqrt(2))
    fc = np.stack([rng.choice([c35, -c35], 2000), rng.choice([np.pi / 4, -np.pi / 4, 3 * np.pi / 4], 2000)], axis=1)
    axis = np.array([[0, 0], [0, np.pi / 2], [np.pi / 2, 0], [0, np.pi], [0, -np.pi / 2], [-np.pi / 2, 0], [0, -np.pi], [np.pi / 4, 0], [0, 1e-300]])
    return np.ascontiguousarray(np.vstack([ll, ll2, fb, fc, axis]))


def hs_filtered(hs, ll, force_cr, level=30):
    out = np.empty(len(ll), np.uint64)
    ns = ctypes.c_uint64(0)
    hs.hs_encode_latlng_filtered(vp(ll), ctypes.c_size_t(len(ll)), level, force_cr, vp(out), ctypes.byref(ns))
    return out, ns.value


def test_filtered_ids_equal_correctly_rounded_ids(hostsim, ref):
    from s2geometry_b200 import synth
    ll = synth.sphere_latlng(2_000_000, 9)
    fast, n_sens = hs_filtered(hostsim, ll, 0)
    exact, n_all = hs_filtered(hostsim, ll, 1)
    assert n_all == len(ll) and n_sens < 1e-3 * len(ll)       # ~6e-5 of points pay for the correctly rounded path
    assert (fast == exact).all()
    old, _ = util.hs_encode_latlng(hostsim, ll)               # the unfiltered implementation
    assert (exact == old).all()
    bl = boundary_latlngs(ref, 20000, 10)
    fast, n_sens = hs_filtered(hostsim, bl, 0)
    exact, _ = hs_filtered(hostsim, bl, 1)
    assert n_sens > 0.2 * len(bl)                              # these inputs really are boundary cases
    assert (fast == exact).all()
    # and against the reference (glibc): any difference must be a boundary-sensitive point
    want = ref.encode_latlng(bl, 30, 1)
    assert (fast != want).mean() < 0.01


def test_wide_hilbert_table_equals_reference_layout(hostsim):
    """The kernels use a 6-bit (16384-entry) Hilbert table, 5 lookups; it must agree with the reference's 4-bit layout."""
    rng = np.random.default_rng(3)
    ij = rng.integers(0, 1 << 30, (500000, 2), dtype=np.uint32)
    ij[:8] = [[0, 0], [(1 << 30) - 1, (1 << 30) - 1], [0, (1 << 30) - 1], [(1 << 30) - 1, 0], [1, 2], [63, 64], [4095, 4096], [1 << 29, 1 << 29]]
    hostsim.hs_check_wide_hilbert.restype = ctypes.c_uint64
    assert hostsim.hs_check_wide_hilbert(vp(ij), ctypes.c_size_t(len(ij))) == 0


def test_from_token_hostsim(hostsim, ref):
    fx = util.fixture("encode_fixture.npz")
    tok, ln, ids = fx["tok"], fx["tok_len"], fx["tok_ids"]
    out = np.empty(len(ids), np.uint64)
    hostsim.hs_cellid_from_token(vp(tok), vp(ln), ctypes.c_size_t(len(ids)), vp(out))
    assert (out == ref.from_token(tok, ln)).all()
    assert (out[ids != 0] == ids[ids != 0]).all()          # round trip (the token "X" of id 0 decodes to 0 too)
    bad = np.zeros((4, 16), np.uint8); bl = np.array([8, 9, 8, 9], np.uint8)
    for k, t in enumerate([b"876b e99", b"876bee99\n", b"876[ee99", b" 876bee99"]):   # s2cell_id_test.cc:362-368
        bad[k, : len(t)] = np.frombuffer(t, np.uint8)
    out = np.empty(4, np.uint64)
    hostsim.hs_cellid_from_token(vp(bad), vp(bl), ctypes.c_size_t(4), vp(out))
    assert (out == 0).all() and (ref.from_token(bad, bl) == 0).all()
    up = np.zeros((1, 16), np.uint8); up[0, :3] = np.frombuffer(b"4AF", np.uint8)
    o = np.empty(1, np.uint64)
    hostsim.hs_cellid_from_token(vp(up), vp(np.array([3], np.uint8)), ctypes.c_size_t(1), vp(o))
    assert o[0] == 0x4AF0000000000000 == ref.from_token(up, np.array([3], np.uint8))[0]


def random_ids_all_levels(ref, n, seed):
    from s2geometry_b200 import synth
    rng = np.random.default_rng(seed)
    leaf = ref.encode_points(synth.sphere_points(n, seed), 30, 1)
    lv = rng.integers(0, 31, n)
    lsb = np.uint64(1) << (2 * (30 - lv)).astype(np.uint64)
    ids = (leaf & (~lsb + np.uint64(1))) | lsb
    # cells along face edges and corners (their neighbours wrap to other faces)
    edge = np.array([ref.from_face_pos_level(f, 0, l) for f in range(6) for l in (0, 1, 5, 30)] +
                    [ref.from_face_pos_level(f, (1 << 60) - 1, l) for f in range(6) for l in (1, 7, 30)], np.uint64)
    return np.concatenate([ids, edge])


def test_edge_neighbors_hostsim(hostsim, ref):
    ids = random_ids_all_levels(ref, 200000, 17)
    out = np.empty((len(ids), 4), np.uint64)
    hostsim.hs_edge_neighbors(vp(ids), ctypes.c_size_t(len(ids)), vp(out))
    assert (out == ref.edge_neighbors(ids)).all()


def test_traversal_helpers_hostsim(hostsim, ref):
    """child / advance / advance_wrap / GetCommonAncestorLevel as the kernels compute them (host build of s2b_core.cuh)
    against the reference, incl. the first and last cell of every level and steps far beyond a level's size."""
    from oracle.reflib import _p, _sz
    rng = np.random.default_rng(23)
    ids = random_ids_all_levels(ref, 100000, 29)
    ends = []
    for level in range(31):
        lsb = 1 << (2 * (30 - level))
        ends += [lsb, (6 << 61) - lsb, 3 * lsb, (3 << 61) + lsb]
    ids = np.concatenate([ids, np.array(ends, np.uint64)])
    n = len(ids)
    others = np.concatenate([ids[rng.permutation(n)[: n // 2]], ids[n // 2:]])
    steps = (rng.choice([-1, 1], n) * 10.0 ** rng.uniform(0, 18.9, n)).astype(np.int64)
    steps[::7] = rng.integers(-3, 4, len(steps[::7]))
    ch = np.empty((n, 4), np.uint64); adv = np.empty(n, np.uint64); advw = np.empty(n, np.uint64); anc = np.empty(n, np.int8)
    hostsim.hs_cellid_traverse(vp(ids), vp(others), vp(steps), ctypes.c_size_t(n), vp(ch), vp(adv), vp(advw), vp(anc))
    want = np.empty(n, np.uint64)
    ref.lib.ref_cellid_advance(_p(ids), _p(steps), _sz(n), 0, _p(want)); assert (adv == want).all()
    ref.lib.ref_cellid_advance(_p(ids), _p(steps), _sz(n), 1, _p(want)); assert (advw == want).all()
    wa = np.empty(n, np.int8)
    ref.lib.ref_cellid_common_ancestor_level(_p(ids), _p(others), _sz(n), _p(wa)); assert (anc == wa).all()
    par = (ids & np.uint64(1)) == 0
    wc = np.empty((int(par.sum()), 4), np.uint64)
    pids = np.ascontiguousarray(ids[par])
    ref.lib.ref_cellid_children(_p(pids), _sz(len(pids)), _p(wc)); assert (ch[par] == wc).all() and (ch[~par] == 0).all()


def test_strict_host_libm_mode_logic_equals_reference(hostsim, ref):
    """The strict lat/lng mode (kernel filter -> correctly rounded ids -> host libm for the points whose leaf cell a 1-ulp
    change of sin/cos could move), restated on the host from the same headers: ids == the reference's S2CellId(S2LatLng)
    for every point, while only ~1e-5 of uniform points are re-evaluated. 16M uniform points, points on / next to cell
    boundaries, and angles outside the device routines' domain."""
    import ctypes
    from s2geometry_b200 import synth
    rng = np.random.default_rng(23)
    sets = [synth.stream_sphere_latlng(9, 0, 16_000_000), boundary_latlngs(ref, 20000, 12),
            np.concatenate([rng.uniform(-4e6, 4e6, (50000, 2)), 10.0 ** rng.uniform(6, 300, (5000, 2)),
                            np.array([[np.inf, 1.0], [np.nan, 0.5], [0.5, np.nan], [0.0, 0.0], [np.pi / 2, 0.0], [0.0, np.pi / 4]])])]
    for k, ll in enumerate(sets):
        ll = np.ascontiguousarray(ll)
        out = np.empty(len(ll), np.uint64)
        nre = ctypes.c_uint64(0)
        hostsim.hs_encode_latlng_strict(util.vp(ll), ctypes.c_size_t(len(ll)), util.vp(out), ctypes.byref(nre))
        assert (out != ref.encode_latlng(ll, 30)).sum() == 0
        if k == 0:
            assert 0 < nre.value < 1e-4 * len(ll)
        else:
            assert nre.value > 0.1 * len(ll)
