"""Sign / CrossingSign chain: the reference oracle reproduces the reference's own known answers, and the device
headers compiled for the host (tests/hostsim) agree with the oracle on those and on adversarial inputs."""
import ctypes
import json
import os

import numpy as np
import pytest

from tests import adversarial, util
from tests.util import vp


def _ka():
    with open(os.path.join(util.GOLDEN, "predicates_known_answers.json")) as f:
        return json.load(f)


KA = _ka()


def _pt(v):
    if v == "origin":
        return np.array(KA["origin"], float)
    return np.array([float(x) for x in v], float)


def hs_sign(hs, abc, kind=0):
    abc = np.ascontiguousarray(abc, np.float64).reshape(-1, 9)
    out = np.empty(len(abc), np.int8)
    ne = ctypes.c_uint64(0)
    hs.hs_sign(vp(abc), ctypes.c_size_t(len(abc)), kind, vp(out), ctypes.byref(ne))
    return out, ne.value


def hs_crossing(hs, abcd, kind=0):
    abcd = np.ascontiguousarray(abcd, np.float64).reshape(-1, 12)
    out = np.empty(len(abcd), np.int8)
    hs.hs_crossing(vp(abcd), ctypes.c_size_t(len(abcd)), kind, vp(out))
    return out


def crossing_cases(ref):
    """Expands the 12 cases through TestCrossings' permutations -> list of (a,b,c,d, crossing_sign, edge_or_vertex)."""
    rows = []

    def one(a, b, c, d, cs, signed):
        if (a == c).all() or (a == d).all() or (b == c).all() or (b == d).all():
            cs = 0
        rows.append((np.concatenate([a, b, c, d]), cs, int(signed != 0)))

    for case in KA["crossings"]:
        a, b, c, d = (ref.normalize(_pt(case[k]))[0] for k in "abcd")
        cs, sg = case["crossing_sign"], case["signed"]
        one(a, b, c, d, cs, sg); one(b, a, c, d, cs, -sg); one(a, b, d, c, cs, -sg); one(b, a, d, c, cs, sg)
        one(a, a, c, d, -1, 0); one(a, b, c, c, -1, 0); one(a, a, c, c, -1, 0); one(a, b, a, b, 0, 1)
        one(c, d, a, b, cs, 0 if cs == 0 else -sg)
    return rows


def run_known_answers(sign_fn, crossing_fn, ref):
    rows = crossing_cases(ref)
    quads = np.array([r[0] for r in rows])
    assert (crossing_fn(quads, 0) == np.array([r[1] for r in rows])).all()
    assert (crossing_fn(quads, 3) == np.array([r[2] for r in rows])).all()
    for inv in KA["invalid"]:
        p = _pt(inv["point"])
        q = np.concatenate([p, p, p, p])[None]
        assert crossing_fn(q, 0)[0] == inv["crossing_sign"]
        assert crossing_fn(q, 3)[0] == inv["edge_or_vertex"]
    for s in KA["symbolic"]:
        a, b, c = _pt(s["a"]), _pt(s["b"]), _pt(s["c"])
        e = s["expected"]
        tri = np.array([np.concatenate(t) for t in ((a, b, c), (b, c, a), (c, a, b), (c, b, a), (b, a, c), (a, c, b))])
        assert (sign_fn(tri, 4) == np.array([e, e, e, -e, -e, -e])).all(), s
    for s in KA["collinear"]:
        a, b, c = _pt(s["a"]), _pt(s["b"]), _pt(s["c"])
        r = sign_fn(np.array([np.concatenate(t) for t in ((a, b, c), (b, c, a), (c, b, a))]), 0)
        assert r[0] != 0 and r[0] == r[1] == -r[2]
    u = KA["stable_underflow"]
    tri = np.concatenate([_pt(u["a"]), _pt(u["b"]), _pt(u["c"])])[None]
    assert sign_fn(tri, 2)[0] == 0 and sign_fn(tri, 3)[0] == 1 and sign_fn(tri, 0)[0] == 1


def test_oracle_known_answers(ref):
    run_known_answers(lambda t, k: ref.sign(t, k), lambda q, k: ref.crossing(q, k), ref)
    # the stateful S2CopyingEdgeCrosser agrees with S2::CrossingSign on the same cases
    quads = np.array([r[0] for r in crossing_cases(ref)])
    assert (ref.crossing(quads, 1) == ref.crossing(quads, 0)).all()


def test_hostsim_known_answers(hostsim, ref):
    run_known_answers(lambda t, k: hs_sign(hostsim, t, k)[0], lambda q, k: hs_crossing(hostsim, q, k), ref)


def test_hostsim_sign_adversarial(hostsim, ref):
    tri = adversarial.sign_triples(20000, 21)
    for kind in (0, 1, 2, 4):
        got, n_exact = hs_sign(hostsim, tri, kind)
        want = ref.sign(tri, kind)
        assert (got == want).all(), (kind, np.where(got != want)[0][:5])
        if kind == 0:
            assert n_exact > 1000  # the exact stage is really exercised
    # exact stage alone (distinct points only: ExactSign requires it)
    a, b, c = tri[:, 0:3], tri[:, 3:6], tri[:, 6:9]
    distinct = ~((a == b).all(1) | (b == c).all(1) | (a == c).all(1)) & np.isfinite(tri).all(1)
    for kind in (3, 5):
        got, _ = hs_sign(hostsim, tri[distinct], kind)
        assert (got == ref.sign(tri[distinct], kind)).all()


def test_hostsim_crossing_adversarial(hostsim, ref):
    quads = adversarial.crossing_quads(20000, 22)
    for kind in (0, 3):
        got = hs_crossing(hostsim, quads, kind)
        want = ref.crossing(quads, kind)
        assert (got == want).all(), (kind, np.where(got != want)[0][:5])
    # VertexCrossing is only defined when a vertex is shared
    a, b, c, d = (quads[:, 3 * k:3 * k + 3] for k in range(4))
    shared = (a == c).all(1) | (a == d).all(1) | (b == c).all(1) | (b == d).all(1)
    assert (hs_crossing(hostsim, quads[shared], 2) == ref.crossing(quads[shared], 2)).all()
    # the stateful copying crosser (what S2ContainsPointQuery uses) is the same function
    assert (ref.crossing(quads, 1) == ref.crossing(quads, 0)).all()
    pts = np.vstack([a[:2000], quads[60000:62000, 0:3]])
    ho = np.empty_like(pts)
    hostsim.hs_ortho(vp(np.ascontiguousarray(pts)), ctypes.c_size_t(len(pts)), vp(ho))
    assert (ho == ref.ortho(pts)).all()
