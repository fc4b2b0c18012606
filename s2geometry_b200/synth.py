"""Synthetic inputs for tests and bench.py: uniform-on-sphere points, lat/lng pairs, a
10k-vertex wobbly loop, many small polygons, degenerate (on-vertex / on-edge / cell-centre) point sets.

Two families. (1) Small host generators on numpy's PCG64 with fixed seeds (test scenes). (2) The COUNTER-BASED streams
(`stream_*`): point i is a pure function of (seed, i) — integer hashing and + - * / sqrt in IEEE double only — restated
here in NumPy and in CUDA (csrc/s2b_synth.cuh), bit-identical on every host and GPU (tests/test_synth_*.py). bench.py
gives each rank its slice of one global stream, so results are the same at any number of GPUs and on the reference arm.
"""
import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def sphere_points(n, seed):
    """n points uniform on the unit sphere, (n,3) float64 (normalised Gaussians)."""
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((n, 3))
    p /= np.sqrt((p * p).sum(axis=1))[:, None]
    return p


def sphere_latlng(n, seed):
    """n (lat, lng) pairs in radians, uniform on the sphere: lat = asin(U(-1,1)), lng = U(-pi,pi)."""
    rng = np.random.default_rng(seed)
    ll = np.empty((n, 2))
    ll[:, 0] = np.arcsin(rng.uniform(-1.0, 1.0, n))
    ll[:, 1] = rng.uniform(-np.pi, np.pi, n)
    return ll


def sphere_points_cuda(n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    p = torch.randn((n, 3), dtype=torch.float64, device=device, generator=g)
    p /= p.norm(dim=1, keepdim=True)
    return p


def sphere_latlng_cuda(n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.rand((n, 2), dtype=torch.float64, device=device, generator=g)
    ll = torch.empty((n, 2), dtype=torch.float64, device=device)
    ll[:, 0] = torch.asin(2 * u[:, 0] - 1)
    ll[:, 1] = (2 * u[:, 1] - 1) * np.pi
    return ll


def frame(center):
    """An orthonormal frame (x, y, z=center)."""
    z = np.asarray(center, float)
    z = z / np.linalg.norm(z)
    a = np.array([1.0, 0, 0]) if abs(z[0]) < 0.9 else np.array([0, 1.0, 0])
    x = np.cross(z, a); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return x, y, z


def wobbly_loop(center, radius_rad, num_vertices, wobble=0.15, lobes=7, phase=0.0):
    """CCW loop around `center`: the construction of S2Loop::MakeRegularLoop (s2loop.cc:1466-1491) with a
    sinusoidal radius wobble. Returns (num_vertices,3) unit vectors."""
    x, y, z = frame(center)
    t = 2 * np.pi * np.arange(num_vertices) / num_vertices
    r = radius_rad * (1 + wobble * np.sin(lobes * t + phase))
    v = (np.cos(r)[:, None] * z + np.sin(r)[:, None] * (np.cos(t)[:, None] * x + np.sin(t)[:, None] * y))
    v /= np.sqrt((v * v).sum(axis=1))[:, None]
    return v


def cap_points(center, radius_rad, n, seed):
    """n points uniform in the spherical cap around center."""
    rng = np.random.default_rng(seed)
    x, y, z = frame(center)
    h = rng.uniform(0, 1 - np.cos(radius_rad), n)
    cz = 1 - h
    sz = np.sqrt(np.maximum(0, 1 - cz * cz))
    t = rng.uniform(0, 2 * np.pi, n)
    p = cz[:, None] * z + sz[:, None] * (np.cos(t)[:, None] * x + np.sin(t)[:, None] * y)
    p /= np.sqrt((p * p).sum(axis=1))[:, None]
    return p


def many_polygons(num, seed, min_vertices=16, max_vertices=256, min_radius_km=1.0, max_radius_km=100.0, hole_fraction=0.2):
    """BASELINE config 4: `num` polygons with random centres, 16-256 vertices, radii log-uniform 1-100 km, a
    fraction with one hole. Returns a list of polygons, each a list of loops ((k,3) arrays; holes are CW)."""
    rng = np.random.default_rng(seed)
    centers = sphere_points(num, seed + 1)
    polys = []
    for i in range(num):
        nv = int(rng.integers(min_vertices, max_vertices + 1))
        rad = np.exp(rng.uniform(np.log(min_radius_km), np.log(max_radius_km))) / 6371.0
        shell = wobbly_loop(centers[i], rad, nv, wobble=0.2, lobes=int(rng.integers(3, 9)), phase=rng.uniform(0, 6.28))
        loops = [shell]
        if rng.uniform() < hole_fraction:
            hv = max(8, nv // 4)
            hole = wobbly_loop(centers[i], rad * 0.4, hv, wobble=0.1, lobes=3, phase=rng.uniform(0, 6.28))[::-1].copy()
            loops.append(hole)
        polys.append(loops)
    return polys


def degenerate_points(loop_vertices):
    """On-vertex, edge-midpoint and quarter-point queries for a loop (the 'D' set of SURVEY.md §8d)."""
    v = np.asarray(loop_vertices)
    w = np.roll(v, -1, axis=0)
    mid = v + w; mid /= np.sqrt((mid * mid).sum(axis=1))[:, None]
    q = 3 * v + w; q /= np.sqrt((q * q).sum(axis=1))[:, None]
    return np.concatenate([v, mid, q])


# ------------------------------------------------------------------------------------------------------------------
# Counter-based streams (csrc/s2b_synth.cuh restated in NumPy; uint64 arithmetic wraps like the C code).
_U = np.uint64
_GOLD = _U(0x9E3779B97F4A7C15)


def _mix(z):
    z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
    return z ^ (z >> _U(31))


def _state(seed, index, attempt):
    with np.errstate(over="ignore"):
        base = _mix(np.array([seed], _U) * _GOLD + _U(0x632BE59BD9B4E019))[0]
        return base ^ (index * _GOLD) ^ (_U(attempt) * _U(0xD1B54A32D192ED03))


def _signed(r):
    return ((r >> _U(12)).astype(np.float64) + 0.5) * 2.0 ** -51 - 1.0


def _unit(r):
    return ((r >> _U(12)).astype(np.float64) + 0.5) * 2.0 ** -52


def _draw2(seed, index, attempt):
    with np.errstate(over="ignore"):
        st = _state(seed, index, attempt)
        return _mix(st), _mix(st + _GOLD)


def stream_sphere_points(seed, first, n):
    """Points [first, first + n) of the uniform-on-sphere stream `seed`: (n, 3) float64."""
    index = np.arange(first, first + n, dtype=_U)
    out = np.empty((n, 3))
    todo = np.arange(n)
    for k in range(64):
        r1, r2 = _draw2(seed, index[todo], k)
        x1, x2 = _signed(r1), _signed(r2)
        s = x1 * x1 + x2 * x2
        ok = s < 1.0
        r = np.sqrt(1.0 - s[ok])
        sel = todo[ok]
        out[sel, 0] = (2.0 * x1[ok]) * r
        out[sel, 1] = (2.0 * x2[ok]) * r
        out[sel, 2] = 1.0 - 2.0 * s[ok]
        todo = todo[~ok]
        if len(todo) == 0:
            return out
    raise AssertionError("rejection sampling did not terminate")   # probability 0.215^64


def _asin(z):
    a = np.abs(z)
    small = a < 0.5
    w = np.where(small, z * z, (1.0 - a) * 0.5)
    p = w * (1.66666666666666657415e-01 + w * (-3.25565818622400915405e-01 + w * (2.01212532134862925881e-01 +
             w * (-4.00555345006794114027e-02 + w * (7.91534994289814532176e-04 + w * 3.47933107596021167570e-05)))))
    q = 1.0 + w * (-2.40339491173441421878e+00 + w * (2.02094576023350569471e+00 + w * (-6.88283971605453293030e-01 +
                   w * 7.70381505559019352791e-02)))
    r = p / q
    s = np.sqrt(w)
    big = 1.57079632679489655800e+00 - 2.0 * (s + s * r)
    return np.where(small, z + z * r, np.where(z < 0, -big, big))


def stream_sphere_latlng(seed, first, n):
    """(lat, lng) radians [first, first + n) of the uniform-on-sphere lat/lng stream `seed`: (n, 2) float64."""
    out = np.empty((n, 2))
    step = 1 << 24
    for lo in range(0, n, step):
        m = min(step, n - lo)
        r1, r2 = _draw2(seed, np.arange(first + lo, first + lo + m, dtype=_U), 0)
        out[lo:lo + m, 0] = _asin(_signed(r1))
        out[lo:lo + m, 1] = _signed(r2) * 3.14159265358979323846
    return out


def stream_cap_points(seed, first, n, center, radius_rad):
    """Points [first, first + n) of the stream `seed` of points uniform in the cap (center, radius): (n, 3) float64."""
    f = cap_frame(center)
    height = cap_height(radius_rad)
    index = np.arange(first, first + n, dtype=_U)
    r0, _ = _draw2(seed, index, 0)
    cz = 1.0 - _unit(r0) * height
    t = 1.0 - cz * cz
    sz = np.sqrt(np.where(t > 0, t, 0.0))
    a = np.ones(n); b = np.zeros(n)
    todo = np.arange(n)
    for k in range(1, 64):
        r1, r2 = _draw2(seed, index[todo], k)
        x1, x2 = _signed(r1), _signed(r2)
        s = x1 * x1 + x2 * x2
        ok = (s < 1.0) & (s > 2.0 ** -40)
        inv = 1.0 / np.sqrt(s[ok])
        a[todo[ok]] = x1[ok] * inv
        b[todo[ok]] = x2[ok] * inv
        todo = todo[~ok]
        if len(todo) == 0:
            break
    p = np.empty((n, 3))
    for c in range(3):
        p[:, c] = cz * f[6 + c] + sz * (a * f[c] + b * f[3 + c])
    inv = 1.0 / np.sqrt((p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) + p[:, 2] * p[:, 2])
    return p * inv[:, None]


def cap_frame(center):
    """The 9 doubles s2b_synth_cap_points_dev takes: orthonormal x, y, z rows with z = the cap axis."""
    x, y, z = frame(center)
    return np.ascontiguousarray(np.concatenate([x, y, z]), dtype=np.float64)


def cap_height(radius_rad):
    return float(1 - np.cos(radius_rad))


def _stream_cuda(fn_name, seed, first, n, cols, device, *extra):
    import ctypes
    from . import _lib
    lib = _lib.load()
    out = torch.empty((n, cols), dtype=torch.float64, device=device)
    with torch.cuda.device(out.device):
        _lib.check(getattr(lib, fn_name)(int(seed), int(first), int(n), *extra, ctypes.c_void_p(out.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def stream_sphere_points_cuda(seed, first, n, device):
    """Device generator of stream_sphere_points (same bits), (n, 3) float64 CUDA tensor."""
    return _stream_cuda("s2b_synth_sphere_points_dev", seed, first, n, 3, device)


def stream_sphere_latlng_cuda(seed, first, n, device):
    return _stream_cuda("s2b_synth_sphere_latlng_dev", seed, first, n, 2, device)


def stream_cap_points_cuda(seed, first, n, center, radius_rad, device):
    import ctypes
    f = cap_frame(center)
    return _stream_cuda("s2b_synth_cap_points_dev", seed, first, n, 3, device, ctypes.c_void_p(f.ctypes.data), ctypes.c_double(cap_height(radius_rad)))
