"""Synthetic inputs for tests and bench.py: uniform-on-sphere points, lat/lng pairs, a
10k-vertex wobbly loop, many small polygons, degenerate (on-vertex / on-edge / cell-centre) point sets.

Host generators use numpy's PCG64 with fixed seeds (deterministic for this image); device generators use torch's
CUDA RNG and exist only to fill HBM quickly for the large benchmark configurations.
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
