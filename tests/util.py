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
