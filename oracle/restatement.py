"""oracle/restatement.py — TEST INFRASTRUCTURE: an independent CPU restatement of the reference's algorithm for the
hot path, in NumPy (vectorised IEEE double arithmetic, no FMA) and pure Python (exact rational arithmetic for the
robust predicates; small cases only). It shares no code with the product (s2geometry_b200/csrc) nor with the compiled
reference (oracle/_ref), so the three implementations check each other:

    reference's own known answers  -> pin this file and oracle/_ref      (tests/test_restatement.py, tests/golden/*.json)
    oracle/_ref (real reference)   -> validates this file on random and adversarial inputs
    this file + oracle/_ref        -> check the CUDA path                 (tests/, smoke())

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it. Citations are reference
file:line under src/s2/.
"""
import math
from fractions import Fraction

import numpy as np

# ---------------------------------------------------------------------------------------------------------------
# Hilbert curve tables (s2coords_internal.h:44-82; lookup construction s2cell_id.cc:75-115)
K_IJ_TO_POS = ((0, 1, 3, 2), (0, 3, 1, 2), (2, 3, 1, 0), (2, 1, 3, 0))
K_POS_TO_IJ = ((0, 1, 3, 2), (0, 2, 3, 1), (3, 2, 0, 1), (3, 1, 0, 2))
K_POS_TO_ORIENTATION = (1, 0, 0, 3)   # kSwapMask, 0, 0, kInvertMask + kSwapMask
SWAP_MASK, INVERT_MASK = 1, 2


def _make_lookup():
    """lookup_pos[iiiijjjjoo] = ppppppppoo and lookup_ij[ppppppppoo] = iiiijjjjoo by the reference's recursion
    (InitLookupCell, s2cell_id.cc:82-103)."""
    pos_tab = np.zeros(1024, np.uint16)
    ij_tab = np.zeros(1024, np.uint16)

    def rec(level, i, j, orig, pos, orient):
        if level == 4:
            ij = (i << 4) + j
            pos_tab[(ij << 2) + orig] = (pos << 2) + orient
            ij_tab[(pos << 2) + orig] = (ij << 2) + orient
            return
        r = K_POS_TO_IJ[orient]
        for k in range(4):
            rec(level + 1, (i << 1) + (r[k] >> 1), (j << 1) + (r[k] & 1), orig, (pos << 2) + k, orient ^ K_POS_TO_ORIENTATION[k])

    for o in (0, SWAP_MASK, INVERT_MASK, SWAP_MASK | INVERT_MASK):
        rec(0, 0, 0, o, 0, o)
    return pos_tab, ij_tab


LOOKUP_POS, LOOKUP_IJ = _make_lookup()


# ---------------------------------------------------------------------------------------------------------------
# Encode: S2CellId(S2Point)  (s2cell_id.cc:309-315)
def xyz_to_face_uv(p):
    """S2::XYZtoFaceUV (s2coords.h:389-419): GetFace = LargestAbsComponent (vector.h:508-513, strict '>') + 3 if negative."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    ax, ay, az = np.abs(x), np.abs(y), np.abs(z)
    face = np.where(ax > ay, np.where(ax > az, 0, 2), np.where(ay > az, 1, 2))
    major = np.choose(face, [x, y, z])
    face = face + np.where(major < 0, 3, 0)
    with np.errstate(all="ignore"):
        u = np.choose(face, [y / x, -x / y, -x / z, z / x, z / y, -y / z])
        v = np.choose(face, [z / x, z / y, -y / z, y / x, -x / y, -x / z])
    return face, u, v


def uv_to_st(u):
    """Quadratic projection (s2coords.h:329-332)."""
    with np.errstate(all="ignore"):
        return np.where(u >= 0, 0.5 * np.sqrt(1 + 3 * u), 1 - 0.5 * np.sqrt(1 - 3 * u))


def st_to_uv(s):
    """s2coords.h:324-327."""
    return np.where(s >= 0.5, (1 / 3.) * (4 * s * s - 1), (1 / 3.) * (1 - 4 * (1 - s) * (1 - s)))


def st_to_ij(s):
    """s2coords.h:345-356: !(s > 0) -> 0 (also NaN); else min((int)(2^30 * s), 2^30 - 1)."""
    with np.errstate(all="ignore"):
        v = np.where(s > 0, 1073741824.0 * s, 0.0)
        v = np.where(np.isnan(v), 0.0, v)
    return np.minimum(v.astype(np.int64), 1073741823)


def from_face_ij(face, i, j):
    """S2CellId::FromFaceIJ (s2cell_id.cc:267-307)."""
    face = face.astype(np.uint64); i = i.astype(np.uint64); j = j.astype(np.uint64)
    n = face << np.uint64(60)
    bits = face & np.uint64(1)
    for k in range(7, -1, -1):
        bits = bits + (((i >> np.uint64(4 * k)) & np.uint64(15)) << np.uint64(6))
        bits = bits + (((j >> np.uint64(4 * k)) & np.uint64(15)) << np.uint64(2))
        bits = LOOKUP_POS[bits.astype(np.int64)].astype(np.uint64)
        n = n | ((bits >> np.uint64(2)) << np.uint64(8 * k))
        bits = bits & np.uint64(3)
    return n * np.uint64(2) + np.uint64(1)


def parent(ids, level):
    """S2CellId::parent(level) (s2cell_id.h:662-668)."""
    ids = np.asarray(ids, np.uint64)
    lsb = np.uint64(1) << np.uint64(2 * (30 - level))
    return (ids & (~lsb + np.uint64(1))) | lsb


def encode_points(xyz, level=30):
    xyz = np.asarray(xyz, np.float64).reshape(-1, 3)
    face, u, v = xyz_to_face_uv(xyz)
    ids = from_face_ij(face, st_to_ij(uv_to_st(u)), st_to_ij(uv_to_st(v)))
    return ids if level >= 30 else parent(ids, level)


def latlng_to_point(latlng_rad):
    """S2LatLng::ToPoint (s2latlng.cc:68-77) with the platform libm, one call per value (s1angle.h:343-357)."""
    ll = np.asarray(latlng_rad, np.float64).reshape(-1, 2)
    out = np.empty((len(ll), 3))
    for k, (lat, lng) in enumerate(ll):
        sp, cp, st, ct = math.sin(lat), math.cos(lat), math.sin(lng), math.cos(lng)
        out[k] = (ct * cp, st * cp, sp)
    return out


def encode_latlng(latlng_rad, level=30):
    """S2CellId(S2LatLng) (s2cell_id.cc:317)."""
    return encode_points(latlng_to_point(latlng_rad), level)


def encode_latlng_e7(e7, level=30):
    """S2LatLng::FromE7 -> S1Angle::E7 = Degrees(1e-7 * v) = (M_PI / 180) * (1e-7 * v)  (s1angle.h:363-377)."""
    e7 = np.asarray(e7, np.int32).reshape(-1, 2).astype(np.float64)
    return encode_latlng((math.pi / 180) * (1e-7 * e7), level)


def to_token(ids):
    """S2CellId::ToToken (s2cell_id.cc:217-233)."""
    out = []
    for v in np.asarray(ids, np.uint64).tolist():
        out.append("X" if v == 0 else ("%016x" % v).rstrip("0"))
    return out


def from_token(tokens):
    """S2CellId::FromToken (s2cell_id.cc:235-254)."""
    out = []
    for t in tokens:
        if len(t) > 16 or any(c not in "0123456789abcdefABCDEF" for c in t):
            out.append(0)
        else:
            out.append(int(t.ljust(16, "0"), 16) if t else 0)
    return np.array(out, np.uint64)


def to_point(ids):
    """S2CellId::ToPoint: ToFaceIJOrientation (s2cell_id.cc:319-355), GetCenterSiTi (s2cell_id.h:555-581),
    FaceSiTitoXYZ (s2coords.cc:68-72), Normalize (vector.h:191-198)."""
    ids = np.asarray(ids, np.uint64)
    face = (ids >> np.uint64(61)).astype(np.int64)
    bits = face & 1
    i = np.zeros(len(ids), np.int64); j = np.zeros(len(ids), np.int64)
    for k in range(7, -1, -1):
        nbits = 2 if k == 7 else 4
        bits = bits + ((((ids >> np.uint64(k * 8 + 1)).astype(np.int64)) & ((1 << (2 * nbits)) - 1)) << 2)
        bits = LOOKUP_IJ[bits].astype(np.int64)
        i += (bits >> 6) << (k * 4)
        j += ((bits >> 2) & 15) << (k * 4)
        bits &= 3
    is_leaf = (ids & np.uint64(1)) != 0
    low = (ids & np.uint64(0xFFFFFFFF)).astype(np.int64)
    low = np.where(low >= 2 ** 31, low - 2 ** 32, low)          # static_cast<int>(id_)
    delta = np.where(is_leaf, 1, np.where(((i ^ (low >> 2)) & 1) != 0, 2, 0))
    si, ti = 2 * i + delta, 2 * j + delta
    u = st_to_uv((1.0 / 2147483648.0) * si); v = st_to_uv((1.0 / 2147483648.0) * ti)
    one = np.ones(len(ids))
    raw = np.select([(face == f)[:, None] for f in range(6)],
                    [np.stack(c, 1) for c in ((one, u, v), (-u, one, v), (-u, -v, one), (-one, -v, -u), (v, -one, -u), (v, u, -one))])
    n = np.sqrt((0.0 + raw[:, 0] * raw[:, 0]) + raw[:, 1] * raw[:, 1] + raw[:, 2] * raw[:, 2])
    inv = np.where(n != 0, 1.0 / n, n)
    return raw * inv[:, None]


def cellunion_contains(sorted_ids, xyz):
    """S2CellUnion::Contains(S2Point) (s2cell_union.cc:285-300,564-566)."""
    ids = np.asarray(sorted_ids, np.uint64)
    lsb = ids & (~ids + np.uint64(1))
    lo, hi = ids - (lsb - np.uint64(1)), ids + (lsb - np.uint64(1))
    leaf = encode_points(xyz)
    k = np.searchsorted(hi, leaf, side="left")        # first cell whose range_max >= leaf
    ok = k < len(ids)
    kk = np.minimum(k, max(len(ids) - 1, 0))
    if len(ids) == 0:
        return np.zeros(len(leaf), np.uint8)
    return (ok & (lo[kk] <= leaf) & (leaf <= hi[kk])).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------
# Robust predicates, exactly: the reference's filters (TriageSign, StableSign) are certified, so for unit-length
# inputs s2pred::Sign(a,b,c) is the exact sign of det[a;b;c], or the symbolic perturbation when that is zero
# (s2predicates.cc:131-297). Python ints/Fractions make the exact part one line.
def _fr(p):
    return tuple(Fraction(float(x)) for x in p)


def _sgn(x):
    return (x > 0) - (x < 0)


def exact_sign(a, b, c, perturb=True):
    """ExactSign (s2predicates.cc:226-262) + SymbolicallyPerturbedSign (:131-222)."""
    pts = [tuple(float(x) for x in a), tuple(float(x) for x in b), tuple(float(x) for x in c)]
    perm = 1
    for x, y in ((0, 1), (1, 2), (0, 1)):
        if pts[x] > pts[y]:
            pts[x], pts[y] = pts[y], pts[x]
            perm = -perm
    xa, xb, xc = _fr(pts[0]), _fr(pts[1]), _fr(pts[2])
    bxc = (xb[1] * xc[2] - xb[2] * xc[1], xb[2] * xc[0] - xb[0] * xc[2], xb[0] * xc[1] - xb[1] * xc[0])
    det = xa[0] * bxc[0] + xa[1] * bxc[1] + xa[2] * bxc[2]
    s = _sgn(det)
    if s == 0 and perturb:
        chain = (bxc[2], bxc[1], bxc[0], xc[0] * xa[1] - xc[1] * xa[0], xc[0], -xc[1], xc[2] * xa[0] - xc[0] * xa[2], xc[2],
                 xa[0] * xb[1] - xa[1] * xb[0], -xb[0], xb[1], xa[0], Fraction(1))
        s = next(_sgn(t) for t in chain if t != 0)
    return perm * s


def _eq(a, b):
    return float(a[0]) == float(b[0]) and float(a[1]) == float(b[1]) and float(a[2]) == float(b[2])


def sign(a, b, c):
    """s2pred::Sign (s2predicates.h:330-335, s2predicates.cc:275-297): 0 iff two points are equal."""
    if _eq(a, b) or _eq(b, c) or _eq(c, a):
        return 0
    return exact_sign(a, b, c, True)


def crossing_sign(a, b, c, d):
    """S2::CrossingSign as a pure function (s2edge_crosser.h:365-388, s2edge_crosser.cc:40-96)."""
    # For unit-length inputs a shared vertex makes the corresponding triage determinant uncertain, so the crosser's
    # fast path (s2edge_crosser.h:376) cannot fire and the shared-vertex test (s2edge_crosser.cc:76) answers 0.
    if _eq(a, c) or _eq(a, d) or _eq(b, c) or _eq(b, d):
        return 0
    if _eq(a, b) or _eq(c, d):
        return -1
    acb, bda = -sign(a, b, c), sign(a, b, d)
    if bda != acb:
        return -1
    cbd = -sign(c, d, b)
    if cbd != acb:
        return -1
    return 1 if sign(c, d, a) == acb else -1


def ortho(a):
    """S2::Ortho (s2pointutil.cc:48-59)."""
    a = [float(x) for x in a]
    ab = [abs(x) for x in a]
    k = (0 if ab[0] > ab[2] else 2) if ab[0] > ab[1] else (1 if ab[1] > ab[2] else 2)
    k = k - 1 if k > 0 else 2
    t = [0.012, 0.0053, 0.00457]
    t[k] = 1.0
    cx = (a[1] * t[2] - a[2] * t[1], a[2] * t[0] - a[0] * t[2], a[0] * t[1] - a[1] * t[0])
    n = math.sqrt((0.0 + cx[0] * cx[0]) + cx[1] * cx[1] + cx[2] * cx[2])
    n = 1.0 / n if n != 0 else n
    return (cx[0] * n, cx[1] * n, cx[2] * n)


def ordered_ccw(a, b, c, o):
    """s2pred::OrderedCCW (s2predicates.cc:299-312)."""
    return (sign(b, o, a) >= 0) + (sign(c, o, b) >= 0) + (sign(a, o, c) > 0) >= 2


def vertex_crossing(a, b, c, d):
    """S2::VertexCrossing (s2edge_crossings.cc:371-391)."""
    if _eq(a, b) or _eq(c, d):
        return False
    if _eq(a, c):
        return _eq(b, d) or ordered_ccw(ortho(a), d, b, a)
    if _eq(b, d):
        return ordered_ccw(ortho(b), c, a, b)
    if _eq(a, d):
        return _eq(b, c) or ordered_ccw(ortho(a), c, b, a)
    if _eq(b, c):
        return ordered_ccw(ortho(b), d, a, b)
    return False


def edge_or_vertex_crossing(a, b, c, d):
    s = crossing_sign(a, b, c, d)
    return s > 0 if s != 0 else vertex_crossing(a, b, c, d)


# ---------------------------------------------------------------------------------------------------------------
# Containment ground truth (small cases): s2shapeutil::ContainsBruteForce (s2shapeutil_contains_brute_force.cc:26-40)
ORIGIN = (-0.0099994664350250197, 0.0025924542609324121, 0.99994664350250195)   # S2::Origin(), s2pointutil.h:98-116


def polygon_edges(loops):
    edges = []
    for loop in loops:
        n = len(loop)
        edges += [(tuple(loop[k]), tuple(loop[(k + 1) % n])) for k in range(n)]
    return edges


def contains_vertex_sign(edges, v):
    """S2ContainsVertexQuery::ContainsSign (s2contains_vertex_query.cc:30-49)."""
    dirs = {}
    for e0, e1 in edges:
        if _eq(e0, v):
            dirs[e1] = dirs.get(e1, 0) + 1
        if _eq(e1, v):
            dirs[e0] = dirs.get(e0, 0) - 1
    ref_dir = ortho(v)
    best, best_sign = ref_dir, 0
    for p in sorted(dirs):
        if dirs[p] != 0 and ordered_ccw(ref_dir, best, p, v):
            best, best_sign = p, dirs[p]
    return best_sign


def reference_point(edges, has_empty_chain=False):
    """s2shapeutil::GetReferencePoint (s2shapeutil_get_reference_point.cc:59-108)."""
    if not edges:
        return ORIGIN, has_empty_chain
    s = contains_vertex_sign(edges, edges[0][0])
    if s != 0:
        return edges[0][0], s > 0
    fwd = sorted(edges); rev = sorted((b, a) for a, b in edges)
    for f, r in zip(fwd, rev):
        if f < r:
            return f[0], contains_vertex_sign(edges, f[0]) > 0
        if r < f:
            return r[0], contains_vertex_sign(edges, r[0]) > 0
    return ORIGIN, has_empty_chain


def polygon_contains_brute_force(loops, p):
    """ContainsBruteForce for a polygon given as loops of unit vectors (semi-open boundary model)."""
    edges = polygon_edges(loops)
    ref, inside = reference_point(edges)
    p = tuple(float(x) for x in p)
    if _eq(ref, p):
        return inside
    for c, d in edges:
        inside ^= edge_or_vertex_crossing(ref, p, c, d)
    return inside


def polygon_contains(loops, p, model=1):
    """S2ContainsPointQuery semantics for one polygon under a vertex model (s2contains_point_query.h:352-389):
    OPEN / CLOSED override the answer when p is one of the polygon's vertices."""
    p = tuple(float(x) for x in p)
    if model != 1 and any(_eq(tuple(v), p) for loop in loops for v in loop):
        return model == 2
    return polygon_contains_brute_force(loops, p)


def shape_contains(dim, chains, p, model=1):
    """One shape of the soup (dimension 0/1/2) under a vertex model: dimension < 2 shapes contain only their vertices and
    only under CLOSED (s2contains_point_query.h:359-369)."""
    if dim == 2:
        return polygon_contains(chains, p, model)
    p = tuple(float(x) for x in p)
    return model == 2 and any(_eq(tuple(v), p) for c in chains for v in c)
