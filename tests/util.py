import ctypes
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def known_answers():
    with open(os.path.join(GOLDEN, "cellid_known_answers.json")) as f:
        return json.load(f)


def fixture(name):
    return np.load(os.path.join(GOLDEN, name))


def hs_encode_points(hs, xyz, level=30):
    xyz = np.ascontiguousarray(xyz, np.float64)
    out = np.empty(len(xyz), np.uint64)
    hs.hs_encode_points(vp(xyz), ctypes.c_size_t(len(xyz)), level, vp(out))
    return out


def hs_encode_latlng(hs, ll, level=30):
    ll = np.ascontiguousarray(ll, np.float64)
    out = np.empty(len(ll), np.uint64)
    ns = ctypes.c_uint64(0)
    hs.hs_encode_latlng(vp(ll), ctypes.c_size_t(len(ll)), level, vp(out), ctypes.byref(ns))
    return out, ns.value


def tok_str(tok, ln):
    return [bytes(tok[i, : ln[i]]).decode("ascii") for i in range(len(ln))]


def leaf_boundary_points(n, seed):
    """S2Points whose (i, j) coordinates sit on, or within 1e-8 .. 1e-4 leaf cells of, integer values (leaf-cell
    boundaries, incl. face edges and centre lines), so that floor(fi) / floor(fj) is as sensitive as it gets."""
    import numpy as np
    rng = np.random.default_rng(seed)
    i = rng.integers(0, (1 << 30) + 1, n).astype(np.float64)
    j = rng.uniform(0, 2.0 ** 30, n)
    eps = rng.choice([0.0, 1e-8, -1e-8, 1e-7, -1e-7, 1e-6, -1e-6, 4e-6, -4e-6, 1e-5, -1e-5, 1e-4, -1e-4], n)
    i = np.clip(i + eps, 0, 2.0 ** 30)
    both = rng.integers(0, 4, n) == 0
    j = np.where(both, np.clip(np.round(j) + rng.choice([0.0, 1e-7, -1e-6], n), 0, 2.0 ** 30), j)
    swap = rng.integers(0, 2, n).astype(bool)
    fi, fj = np.where(swap, j, i), np.where(swap, i, j)
    s_, t_ = fi / 2.0 ** 30, fj / 2.0 ** 30
    st2uv = lambda s: np.where(s >= 0.5, (1 / 3.) * (4 * s * s - 1), (1 / 3.) * (1 - 4 * (1 - s) * (1 - s)))
    u, v = st2uv(s_), st2uv(t_)
    face = rng.integers(0, 6, n)
    one = np.ones(n)
    xyz = np.select([(face == k)[:, None] for k in range(6)],
                    [np.stack(c, 1) for c in ((one, u, v), (-u, one, v), (-u, -v, one), (-one, -v, -u), (v, -one, -u), (v, u, -one))])
    scale = np.where(rng.integers(0, 2, n)[:, None] == 0, 1.0 / np.sqrt((xyz * xyz).sum(1))[:, None], 1.0)   # half normalised, half on the cube
    return np.ascontiguousarray(xyz * scale)
