"""oracle/restatement.py (the NumPy / exact-rational restatement) is pinned by the reference's own known answers and
validated against the compiled reference (oracle/_ref) — so the CUDA path is checked by two independent oracles."""
import json
import os

import numpy as np
import pytest

from oracle import restatement as rs
from tests import adversarial, scenes, util

KA = util.known_answers()
with open(os.path.join(util.GOLDEN, "predicates_known_answers.json")) as f:
    PKA = json.load(f)


def test_encode_known_answers_and_fixture():
    for row in KA["point_to_id"]:
        assert rs.encode_points(np.array([row["xyz"]]))[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_id"]:
        assert rs.encode_latlng(np.radians(np.array([[row["lat"], row["lng"]]])))[0] == int(row["id"], 16)
    for row in KA["latlng_deg_to_face"]:
        ll = np.array([[(np.pi / 180) * row["lat"], (np.pi / 180) * row["lng"]]])
        assert int(rs.encode_latlng(ll)[0]) >> 61 == row["face"]
    ids = np.array([int(r["id"], 16) for r in KA["tokens"]], np.uint64)
    assert rs.to_token(ids) == [r["token"] for r in KA["tokens"]]
    assert (rs.from_token([r["token"] for r in KA["tokens"]]) == np.where(ids == 0, 0, ids)).all()
    fx = util.fixture("encode_fixture.npz")
    for lvl in (30, 20, 12, 0):
        assert (rs.encode_points(fx["xyz"], lvl) == fx["xyz_ids_L%d" % lvl]).all()
    assert (rs.encode_latlng(fx["latlng"][:3000], 30) == fx["latlng_ids_L30"][:3000]).all()
    assert (rs.encode_latlng_e7(fx["e7"][:2000], 30) == fx["e7_ids_L30"][:2000]).all()
    assert rs.to_token(fx["tok_ids"]) == util.tok_str(fx["tok"], fx["tok_len"])


def test_encode_and_cell_centres_vs_reference(ref):
    from s2geometry_b200 import synth
    p = synth.sphere_points(300000, 71)
    assert (rs.encode_points(p) == ref.encode_points(p, 30)).all()
    rng = np.random.default_rng(5)
    leaf = ref.encode_points(p[:50000], 30)
    ids = np.array([rs.parent(leaf[k:k + 1], int(l))[0] for k, l in enumerate(rng.integers(0, 31, len(leaf)))], np.uint64)
    assert (rs.to_point(ids) == ref.to_point(ids)).all()
    cover = ref.cover_cap(np.array([0.4, 0.5, 0.76]), 0.05, 500)
    q = np.concatenate([synth.cap_points(np.array([0.4, 0.5, 0.76]), 0.08, 20000, 3), ref.to_point(cover)])
    assert (rs.cellunion_contains(cover, q) == ref.cellunion_contains(cover, q)).all()


def _pt(v):
    return np.array(PKA["origin"] if v == "origin" else [float(x) for x in v], float)


def test_predicates_known_answers(ref):
    for case in PKA["crossings"]:
        a, b, c, d = (ref.normalize(_pt(case[k]))[0] for k in "abcd")
        cs, sg = case["crossing_sign"], case["signed"]
        want = 0 if any((x == y).all() for x in (a, b) for y in (c, d)) else cs
        assert rs.crossing_sign(a, b, c, d) == want
        assert rs.edge_or_vertex_crossing(a, b, c, d) == (sg != 0)
        shared = lambda p, q, r, t: any((x == y).all() for x in (p, q) for y in (r, t))
        assert rs.crossing_sign(a, a, c, d) == (0 if shared(a, a, c, d) else -1)
        assert rs.crossing_sign(a, b, c, c) == (0 if shared(a, b, c, c) else -1)
    for s in PKA["symbolic"]:
        a, b, c = _pt(s["a"]), _pt(s["b"]), _pt(s["c"])
        e = s["expected"]
        assert [rs.sign(*t) for t in ((a, b, c), (b, c, a), (c, a, b), (c, b, a), (b, a, c), (a, c, b))] == [e, e, e, -e, -e, -e]
    u = PKA["stable_underflow"]
    assert rs.sign(_pt(u["a"]), _pt(u["b"]), _pt(u["c"])) == 1
    for s in PKA["collinear"]:
        a, b, c = _pt(s["a"]), _pt(s["b"]), _pt(s["c"])
        assert rs.sign(a, b, c) != 0 and rs.sign(a, b, c) == rs.sign(b, c, a) == -rs.sign(c, b, a)


def test_predicates_vs_reference_on_unit_vectors(ref):
    tri = adversarial.sign_triples(300, 91)
    unit = np.abs(np.sqrt((tri.reshape(-1, 3, 3) ** 2).sum(2)) - 1).max(1) < 1e-12
    tri = tri[unit]
    got = np.array([rs.sign(t[0:3], t[3:6], t[6:9]) for t in tri])
    assert (got == ref.sign(tri, 0)).all()
    quads = adversarial.crossing_quads(200, 92)
    unit = np.abs(np.sqrt((quads.reshape(-1, 4, 3) ** 2).sum(2)) - 1).max(1) < 1e-12
    quads = quads[unit]
    got = np.array([rs.crossing_sign(q[0:3], q[3:6], q[6:9], q[9:12]) for q in quads])
    assert (got == ref.crossing(quads, 0)).all()
    got = np.array([rs.edge_or_vertex_crossing(q[0:3], q[3:6], q[6:9], q[9:12]) for q in quads])
    assert (got == ref.crossing(quads, 3)).all()


def test_containment_vs_reference_small_cases(ref):
    ka = scenes.known_answers()
    q = scenes.latlng_deg_to_point(ka["queries"])
    spec = ka["index"]
    shapes = [(0, [scenes.latlng_deg_to_point(spec["points"])])] + [(1, [scenes.latlng_deg_to_point(l)]) for l in spec["polylines"]] + \
             [(2, [scenes.latlng_deg_to_point(loop) for loop in poly]) for poly in spec["polygons"]]
    for name, m in scenes.MODELS.items():
        got = [int(any(rs.shape_contains(dim, chains, p, m) for dim, chains in shapes)) for p in q]
        assert got == ka["contains"][name], name
    from s2geometry_b200 import synth
    from s2geometry_b200.shapes import ShapeSoup
    loop = synth.wobbly_loop(np.array([0.3, -0.5, 0.8]), 0.17, 60)
    soup = ShapeSoup(); soup.add_polygon([loop]); soup = soup.freeze()
    ri = ref.index_create(soup)
    pts = np.concatenate([synth.cap_points(np.array([0.3, -0.5, 0.8]), 0.22, 150, 4), synth.degenerate_points(loop)])
    for m in (0, 1, 2):
        got = np.array([rs.polygon_contains([loop], p, m) for p in pts], np.uint8)
        assert (got == ri.contains(pts, m, 1)).all(), m
    assert (np.array([rs.polygon_contains_brute_force([loop], p) for p in pts], np.uint8) == ri.brute_force_contains(0, pts, 1)).all()
