"""Differential fuzzing of the containment path against the reference: random scenes of touching / nested / shared-vertex
/ sliver / pole-and-face-crossing polygons, polylines and points, queried at vertices, edge points and random points
under all three vertex models, through the library's own index builder AND the reference-built index."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _norm(p):
    p = np.asarray(p, float)
    return p / np.sqrt((p * p).sum(axis=-1))[..., None]


def random_scene(rng):
    from s2geometry_b200 import synth
    from s2geometry_b200.shapes import ShapeSoup
    soup = ShapeSoup()
    kind = int(rng.integers(0, 6))
    special = []
    if kind == 0:      # grid of squares sharing edges and vertices (lat/lng aligned, some on the equator / prime meridian)
        step = np.radians(float(rng.choice([0.5, 2.0, 10.0])))
        lat0 = float(rng.choice([0.0, -step, 0.3])); lng0 = float(rng.choice([0.0, -step, 1.0, np.pi - step]))
        for a in range(3):
            for b in range(3):
                ll = [(lat0 + a * step, lng0 + b * step), (lat0 + a * step, lng0 + (b + 1) * step),
                      (lat0 + (a + 1) * step, lng0 + (b + 1) * step), (lat0 + (a + 1) * step, lng0 + b * step)]
                v = np.array([[np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)] for la, lo in ll])
                soup.add_polygon([v])
                special.append(v)
    elif kind == 1:    # nested rings (shell, hole, island) around a random centre, incl. poles and cube corners
        c = _norm([rng.choice([0, 1, -1]) + 1e-3 * rng.standard_normal() for _ in range(3)]) if rng.random() < 0.5 else synth.sphere_points(1, int(rng.integers(1 << 30)))[0]
        if not np.isfinite(c).all() or np.linalg.norm(c) == 0:
            c = np.array([0, 0, 1.0])
        r = float(rng.uniform(0.01, 0.8))
        nv = int(rng.integers(3, 40))
        shell = synth.wobbly_loop(c, r, nv, 0.1, 3)
        hole = synth.wobbly_loop(c, r * 0.6, nv, 0.1, 3)[::-1].copy()
        island = synth.wobbly_loop(c, r * 0.3, nv, 0.1, 3)
        soup.add_polygon([shell, hole]); soup.add_polygon([island])
        special += [shell, hole, island]
    elif kind == 2:    # fan of thin slivers meeting at one vertex
        c = synth.sphere_points(1, int(rng.integers(1 << 30)))[0]
        x, y, z = synth.frame(c)
        k = int(rng.integers(3, 14))
        ang = np.sort(rng.uniform(0, 2 * np.pi, 2 * k))
        r = float(rng.uniform(0.001, 0.3))
        for t in range(k):
            a0, a1 = ang[2 * t], ang[2 * t + 1]
            v = _norm(np.array([z, z * np.cos(r) + np.sin(r) * (np.cos(a0) * x + np.sin(a0) * y), z * np.cos(r) + np.sin(r) * (np.cos(a1) * x + np.sin(a1) * y)]))
            soup.add_polygon([v]); special.append(v)
    elif kind == 3:    # complement of a small loop + the loop itself + a polyline and points on its vertices
        c = synth.sphere_points(1, int(rng.integers(1 << 30)))[0]
        v = synth.wobbly_loop(c, float(rng.uniform(0.01, 0.5)), int(rng.integers(3, 60)))
        soup.add_polygon([v[::-1].copy()]); soup.add_polygon([v]); soup.add_polyline(v[: max(2, len(v) // 2)]); soup.add_points(v[::3])
        special.append(v)
    elif kind == 4:    # many overlapping random polygons
        for loops in synth.many_polygons(int(rng.integers(5, 60)), int(rng.integers(1 << 30)), 3, 30, 300, 4000, 0.3):
            soup.add_polygon(loops); special.append(loops[0])
    else:              # big polygon covering several faces + its reflection
        v = synth.wobbly_loop(np.array([0.2, 0.1, 1.0]), float(rng.uniform(1.0, 1.5)), int(rng.integers(4, 200)), 0.05, 5)
        soup.add_polygon([v]); soup.add_polygon([-v[::-1]]); special.append(v)
    soup = soup.freeze()
    vs = np.vstack(special)
    w = np.roll(vs, -1, axis=0)
    q = [vs, _norm(vs + w), _norm(3 * vs + w), _norm(vs + 1e-9 * rng.standard_normal(vs.shape)),
         _norm(vs + 0.05 * rng.standard_normal(vs.shape)), synth.sphere_points(3000, int(rng.integers(1 << 30)))]
    return soup, np.ascontiguousarray(np.vstack(q))


def test_fuzz_scenes_vs_reference(ref):
    import torch
    from s2geometry_b200.index import ContainsPointQuery, FlatIndex, ShapeIndex, build_flat_index
    rng = np.random.default_rng(2026)
    checked = 0
    for scene in range(60):
        soup, q = random_scene(rng)
        ri = ref.index_create(soup, int(rng.choice([0, 0, 2, 50])))
        own = ShapeIndex(build_flat_index(soup, int(rng.choice([0, 0, 1, 100]))))
        via_ref = ShapeIndex(FlatIndex(**ri.flatten()))
        dq = torch.from_numpy(q).cuda()
        for m in (0, 1, 2):
            want_any = ri.contains(q, m, 1)
            want_hist = ri.shape_histogram(q, m, 1)
            roff, rids = ri.containing_shapes(q, m)
            for ix in (own, via_ref):
                query = ContainsPointQuery(ix, m)
                assert (query.contains(dq).cpu().numpy() == want_any).all(), (scene, m)
                assert (query.shape_histogram(dq).cpu().numpy().view(np.uint64) == want_hist).all(), (scene, m)
                off, ids = query.containing_shape_ids(q)
                assert (off == roff).all() and (ids == rids).all(), (scene, m)
            s = int(rng.integers(0, soup.num_shapes))
            assert (ContainsPointQuery(own, m).shape_contains(s, q) == ri.shape_contains(s, q, m, 1)).all(), (scene, m, s)
            checked += 3 * len(q)
    assert checked > 1_000_000
