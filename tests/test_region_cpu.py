"""S2ShapeIndexRegion cell tests (s2b_region.cuh compiled for the host by tests/hostsim) against the reference's
S2ShapeIndexRegion, S2::ClipToPaddedFace and S2::IntersectsRect."""
import ctypes

import pytest

import numpy as np

from s2geometry_b200 import synth
from s2geometry_b200.index import FlatIndex
from tests import region_scenes, scenes
from tests.util import vp

KMAX = 9 * np.sqrt(0.5) * np.finfo(float).eps + 3 * np.sqrt(2.0) * np.finfo(float).eps


def hs_region_cells(hs, flat, ids, mode):
    out = np.zeros(len(ids), np.uint8)
    bad = ctypes.c_uint64(0)
    st = flat.as_struct()
    assert hs.hs_region_cells(ctypes.byref(st), vp(ids), ctypes.c_size_t(len(ids)), mode, vp(out), ctypes.byref(bad)) == 0
    return out, bad.value


def hs_region_shapes(hs, flat, ids):
    hs.hs_region_intersecting_shapes.restype = ctypes.c_size_t
    st = flat.as_struct()
    off = np.zeros(len(ids) + 1, np.uint32)
    total = hs.hs_region_intersecting_shapes(ctypes.byref(st), vp(ids), ctypes.c_size_t(len(ids)), vp(off), None, None, ctypes.c_size_t(0))
    shp = np.zeros(max(total, 1), np.int32); con = np.zeros(max(total, 1), np.uint8)
    hs.hs_region_intersecting_shapes(ctypes.byref(st), vp(ids), ctypes.c_size_t(len(ids)), vp(off), vp(shp), vp(con), ctypes.c_size_t(total))
    return off, shp[:total], con[:total]


def test_clip_to_padded_face_and_intersects_rect_bit_exact(hostsim, ref):
    rng = np.random.default_rng(21)
    n = 60000
    a = synth.sphere_points(n, 31)
    # edges of every length: long random ones, short ones, some exactly on face boundaries / cube diagonals
    step = rng.choice([1.5, 0.3, 1e-2, 1e-5, 1e-9], n)[:, None]
    b = scenes._norm(a + step * rng.standard_normal((n, 3)))
    special = scenes._norm(np.array([[1, 1, 0], [1, 1, 1], [1, 0, 0], [0, 1, 1], [1, -1, 1], [1, 1e-300, 0], [-1, -1, -1]], float))
    a[: len(special)] = special
    b[len(special): 2 * len(special)] = special
    a[-50:] = b[-50:]                       # degenerate edges (RobustCrossProd returns Ortho)
    face = rng.integers(0, 6, n).astype(np.int32)
    for padding in (0.0, KMAX, 1e-3):
        uv = np.zeros((n, 4)); ok = np.zeros(n, np.uint8); sup = np.zeros(n, np.uint8)
        hostsim.hs_clip_to_padded_face(vp(a), vp(b), vp(face), ctypes.c_size_t(n), ctypes.c_double(padding), vp(uv), vp(ok), vp(sup))
        want_uv, want_ok = ref.clip_to_padded_face(a, b, face, padding)
        s = sup.astype(bool)
        assert s.mean() > 0.99
        assert (ok[s] == want_ok[s]).all()
        hit = s & (ok == 1)
        assert hit.sum() > 5000
        assert (uv[hit].view(np.uint64) == want_uv[hit].view(np.uint64)).all()      # bit for bit
        # IntersectsRect on the clipped edges against cell-sized and face-sized rectangles
        c = rng.uniform(-1, 1, (n, 2)); h = rng.choice([1e-6, 1e-3, 0.1, 1.0], n)
        rect = np.stack([c[:, 0] - h, c[:, 0] + h, c[:, 1] - h, c[:, 1] + h], 1)
        rect[::7, 0] = uv[::7, 0]            # rectangle sides through an endpoint
        rect[::11, 3] = uv[::11, 3]
        got = np.zeros(n, np.uint8)
        hostsim.hs_intersects_rect(vp(want_uv), vp(np.ascontiguousarray(rect)), ctypes.c_size_t(n), vp(got))
        want = ref.intersects_rect(want_uv, rect)
        assert (got == want).all()
        assert 0.02 < want.mean() < 0.98


def test_region_cell_tests_match_reference(hostsim, ref):
    for name, soup in region_scenes.region_scene_list():
        ri = ref.index_create(soup)
        flat = FlatIndex(**ri.flatten())
        targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 5)
        rel, _ = ri.locate_cells(targets)
        assert all((rel == k).sum() > 50 for k in (0, 1, 2)), name
        for mode in (0, 1):
            got, bad = hs_region_cells(hostsim, flat, targets, mode)
            assert bad == 0, name
            want = ri.region_cells(targets, mode)
            assert (got == want).all(), (name, mode, int((got != want).sum()))
            inside = rel == 0
            assert 0 < want[inside].sum() < inside.sum(), (name, mode)      # both answers occur below the index cells
        off, shp, con = hs_region_shapes(hostsim, flat, targets)
        woff, wshp, wcon = ri.region_intersecting_shapes(targets)
        assert (off == woff).all() and (shp == wshp).all() and (con == wcon).all(), name
        assert 0 < wcon.sum() < len(wcon), name


def _nudge(p, k):
    """p with every coordinate moved by k[i] ulps."""
    q = p.copy()
    for _ in range(int(np.abs(k).max())):
        q = np.where(k > 0, np.nextafter(q, np.inf), np.where(k < 0, np.nextafter(q, -np.inf), q))
        k = k - np.sign(k)
    return q


def robust_cross_prod_cases(seed=5, n=20000):
    """Pairs that reach every stage of S2::RobustCrossProd: endpoints 0..40 ulps apart (long double stage), further apart
    (double stage), identical, nearly and exactly antipodal, exactly proportional (symbolic), with zero components,
    and on the coordinate axes."""
    rng = np.random.default_rng(seed)
    a = scenes._norm(rng.standard_normal((n, 3)))
    a[::7, rng.integers(0, 3)] = 0.0
    a = scenes._norm(a)
    k = rng.integers(-40, 41, a.shape); k[::3] = rng.integers(-2, 3, (len(k[::3]), 3))
    near = _nudge(a, k)
    anti = -_nudge(a, rng.integers(-30, 31, a.shape))
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], float)
    ax_a = np.repeat(axes, 6, 0); ax_b = np.tile(axes, (6, 1))
    prop = a * (1 + 2.0 ** -rng.integers(20, 52, (n, 1)))                      # (1 + eps) * a: often exactly proportional
    pow2 = np.array([[1, 0.5, 0.25], [0.5, 0.5, 0], [0, 0.75, 0.75], [0.375, 0, 0]])
    pa = np.concatenate([pow2, pow2, -pow2, pow2]); pb = np.concatenate([2 * pow2, -pow2, -4 * pow2, pow2 * 1.5])   # exactly proportional
    m = n // 8
    A = np.concatenate([a, a, a, ax_a, a, pa, a[:2000], a[:m], a[:m], a[:m], a[:m], a[:m], a[:m] * 2.0 ** -200])
    B = np.concatenate([near, a, anti, ax_b, prop, pb, scenes._norm(a[:2000] + 1e-9 * rng.standard_normal((2000, 3))),
                        -a[:m], 2 * a[:m], -0.25 * a[:m], a[:m] * 2.0 ** -600,       # exactly (anti)parallel: exact + symbolic stage
                        near[:m] * 2.0 ** -600, near[:m] * 2.0 ** -300])             # tiny non-zero exact products: NormalizableFromExact's scaling
    return np.ascontiguousarray(A), np.ascontiguousarray(B)


def hs_robust_cross_prod(hostsim, a, b):
    out = np.empty_like(a); stage = np.empty(len(a), np.uint8)
    hostsim.hs_robust_cross_prod(vp(a), vp(b), ctypes.c_size_t(len(a)), vp(out), vp(stage))
    return out, stage


def test_robust_cross_prod_every_stage_matches_reference(hostsim, ref):
    """S2::RobustCrossProd restated with all of its stages (double; a == b; x87 long double, emulated bit for bit; exact
    arithmetic; symbolic perturbation) returns the reference's vector, bit for bit, on pairs that reach each stage."""
    a, b = robust_cross_prod_cases()
    got, stage = hs_robust_cross_prod(hostsim, a, b)
    want = ref.robust_cross_prod(a, b)
    counts = np.bincount(stage, minlength=4)
    assert (counts > [1000, 1000, 1000, 1000]).all(), counts          # every stage is exercised
    bad = (got.view(np.uint64) != want.view(np.uint64)).any(axis=1) & ~((got == 0) & (want == 0)).all(axis=1)
    assert not bad.any(), (int(bad.sum()), np.bincount(stage[bad], minlength=4), a[bad][:3], b[bad][:3], got[bad][:3], want[bad][:3])
    assert (got.view(np.uint64) == want.view(np.uint64)).all()          # signed zeros included (clip_exit_axis reads sign bits)


def test_long_double_threshold_constant():
    """kMinNorm^2 of GetStableCrossProd<long double> (s2edge_crossings.cc:128-135) as embedded in s2b_region.cuh:
    recomputed here in the x87's format (numpy.longdouble on x86-64)."""
    ld = np.longdouble
    if np.finfo(ld).nmant != 63:
        pytest.skip("numpy.longdouble is not the x87 80-bit format here")
    dbl_err, t_err, sqrt3 = 2.0 ** -53, ld(2.0) ** -64, 1.7320508075688772935274463415058
    k_min = ld(32 * sqrt3 * dbl_err) / (ld(6 * dbl_err) / t_err - ld(1 + 2 * sqrt3))
    sq = k_min * k_min
    mant, exp = np.frexp(sq)
    assert int(mant * ld(2.0) ** 64) == 0xAACA6DBF0CAB2C09 and int(exp) - 64 == -185


def test_clipping_edges_that_need_the_extended_stages(hostsim, ref):
    """ClipToPaddedFace on edges whose RobustCrossProd leaves the double-precision stage (endpoints a few ulps apart,
    antipodal): clipped (u, v) bit-identical to the reference."""
    a, b = robust_cross_prod_cases(seed=9, n=4000)
    for face in range(6):
        f = np.full(len(a), face, np.int32)
        uv = np.zeros((len(a), 4)); ok = np.zeros(len(a), np.uint8); plain = np.ones(len(a), np.uint8)
        hostsim.hs_clip_to_padded_face(vp(a), vp(b), vp(f), ctypes.c_size_t(len(a)), ctypes.c_double(1e-9), vp(uv), vp(ok), vp(plain))
        want_uv, want_ok = ref.clip_to_padded_face(a, b, f, 1e-9)
        assert (ok == want_ok).all()
        hit = ok.astype(bool)
        assert (uv[hit] == want_uv[hit]).all()
        assert (plain == 0).sum() > 1000


def test_region_cell_tests_over_the_librarys_own_builder(hostsim, ref):
    """The library's host builder produces the reference's cells at the same granularity, so the region answers over
    its index equal the reference's too (they depend on the cell structure: SUBDIVIDED targets "may intersect")."""
    from s2geometry_b200.index import build_flat_index
    for name, soup in region_scenes.region_scene_list():
        ri = ref.index_create(soup)
        flat = build_flat_index(soup, 0)
        if len(flat.cell_ids) != len(ri.flatten()["cell_ids"]) or (flat.cell_ids != ri.flatten()["cell_ids"]).any():
            continue        # identical cells are a property of these scenes, not a guarantee of the builder
        targets = region_scenes.region_targets(ref, flat.cell_ids, soup.vertices, 9, 500)
        for mode in (0, 1):
            got, bad = hs_region_cells(hostsim, flat, targets, mode)
            assert bad == 0 and (got == ri.region_cells(targets, mode)).all(), (name, mode)


def test_cell_union_bound_matches_reference(hostsim, ref):
    """GetCellUnionBound: single-face indexes (children of the smallest enclosing cell), multi-face ones, one cell, one
    cell per face."""
    from s2geometry_b200.shapes import ShapeSoup
    from s2geometry_b200 import synth
    cases = [soup for _, soup in region_scenes.region_scene_list()]
    for centre, radius, nv in (([1, 0.01, 0.02], 1e-3, 12), ([1, 1, 1], 0.3, 30), ([0.2, -0.9, 0.1], 1e-6, 5), ([0, 0, 1], 1.4, 40)):
        s = ShapeSoup(); s.add_polygon([synth.wobbly_loop(np.array(centre, float), radius, nv)]); cases.append(s.freeze())
    s = ShapeSoup(); s.add_points(scenes._norm(np.array([[1, 0.3, 0.2]], float))); cases.append(s.freeze())
    s = ShapeSoup(); s.add_points(scenes._norm(np.array([[1, .1, .1], [-1, .1, .1], [.1, 1, .1], [.1, -1, .1], [.1, .1, 1], [.1, .1, -1]], float))); cases.append(s.freeze())
    seen = set()
    for soup in cases:
        ri = ref.index_create(soup)
        ids = np.ascontiguousarray(ri.flatten()["cell_ids"], np.uint64)
        out = np.zeros(6, np.uint64)
        n = hostsim.hs_region_cell_union_bound(vp(ids), ctypes.c_size_t(len(ids)), vp(out))
        want = ri.region_cell_union_bound()
        assert n == len(want) and (out[:n] == want).all()
        seen.add(n)
    assert {1, 4, 6} & seen and len(seen) >= 3
