"""Test scenes for the point-in-shape path: geometry (ShapeSoup) + query point sets, shared by CPU and GPU tests."""
import json
import os

import numpy as np

from s2geometry_b200 import synth
from s2geometry_b200.shapes import ShapeSoup

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = {"OPEN": 0, "SEMI_OPEN": 1, "CLOSED": 2}


def latlng_deg_to_point(deg):
    """S2LatLng::FromDegrees(lat, lng).ToPoint() with the platform libm (text-format fixtures; s2text_format.cc)."""
    deg = np.asarray(deg, np.float64).reshape(-1, 2)
    rad = (np.pi / 180) * deg
    lat, lng = rad[:, 0], rad[:, 1]
    return np.stack([np.cos(lng) * np.cos(lat), np.sin(lng) * np.cos(lat), np.sin(lat)], axis=1)


def text_index(spec):
    soup = ShapeSoup()
    if spec["points"]:
        soup.add_points(latlng_deg_to_point(spec["points"]))
    for line in spec["polylines"]:
        soup.add_polyline(latlng_deg_to_point(line))
    for poly in spec["polygons"]:
        soup.add_polygon([latlng_deg_to_point(loop) for loop in poly])
    return soup.freeze()


def known_answers():
    with open(os.path.join(GOLDEN, "contains_known_answers.json")) as f:
        return json.load(f)


LOOP_CENTER = np.array([0.3, -0.5, 0.8])
LOOP_RADIUS = 0.17


def loop_scene(num_vertices=10000):
    """BASELINE config 3: one wobbly loop (default 10k vertices, radius ~10 degrees)."""
    soup = ShapeSoup()
    v = synth.wobbly_loop(LOOP_CENTER, LOOP_RADIUS, num_vertices)
    soup.add_polygon([v])
    return soup.freeze(), v


def loop_points(v, n_uniform, n_near, seed):
    """U (uniform on the sphere), B (uniform in the loop's bounding cap), D (every vertex, edge midpoint, quarter point)."""
    u = synth.sphere_points(n_uniform, seed)
    b = synth.cap_points(LOOP_CENTER, LOOP_RADIUS * 1.25, n_near, seed + 1)
    d = synth.degenerate_points(v)
    return u, b, d


def axis_aligned_scene():
    """Geometry that forces exact-zero determinants: loops whose vertices lie on coordinate planes and on S2 cell
    corners, a polyline on the equator, and a point set; queries share those zero coordinates."""
    soup = ShapeSoup()
    # a 'diamond' around +x with vertices in the z = 0 and y = 0 planes
    a = 0.2
    soup.add_polygon([_norm(np.array([[1, -a, 0], [1, 0, -a], [1, a, 0], [1, 0, a]], float))])
    # a square on face 2 with axis-parallel edges through cell-corner-like points
    b = 0.25
    soup.add_polygon([_norm(np.array([[-b, -b, 1], [b, -b, 1], [b, b, 1], [-b, b, 1]], float))])
    # shell with a hole, both on the x = 0 / y = 0 cross
    c, h = 0.3, 0.1
    soup.add_polygon([_norm(np.array([[0, -1, -c], [c, -1, 0], [0, -1, c], [-c, -1, 0]], float)),
                      _norm(np.array([[0, -1, -h], [-h, -1, 0], [0, -1, h], [h, -1, 0]], float))])
    # equatorial polyline and some points (dimension 1 and 0 shapes)
    t = np.linspace(-0.5, 0.5, 9)
    soup.add_polyline(np.stack([np.cos(t), np.sin(t), np.zeros_like(t)], axis=1))
    soup.add_points(_norm(np.array([[1, 0, 0], [1, 0.1, 0], [0, 0, 1], [0.25, 0.25, 1]], float)))
    return soup.freeze()


def axis_aligned_queries(seed=5):
    rng = np.random.default_rng(seed)
    t = rng.uniform(-0.35, 0.35, 400)
    z = np.zeros_like(t)
    o = np.ones_like(t)
    q = [np.stack([o, t, z], 1), np.stack([o, z, t], 1), np.stack([t, z, o], 1), np.stack([z, t, o], 1),
         np.stack([t, t, o], 1), np.stack([t, -o, z], 1), np.stack([z, -o, t], 1),
         np.stack([np.cos(t * 2), np.sin(t * 2), z], 1)]
    # exact vertices and edge midpoints of the axis-aligned shapes
    a, b = 0.2, 0.25
    verts = np.array([[1, -a, 0], [1, 0, -a], [1, a, 0], [1, 0, a], [-b, -b, 1], [b, -b, 1], [b, b, 1], [-b, b, 1],
                      [1, 0, 0], [0, 0, 1], [0, -1, 0], [1, 0.1, 0], [0.25, 0.25, 1], [b, 0, 1], [0, b, 1], [1, a / 2, a / 2]], float)
    q.append(verts)
    return _norm(np.vstack(q))


def _norm(p):
    p = np.asarray(p, float)
    return p / np.sqrt((p * p).sum(axis=-1))[..., None]


def polygons_scene(num, seed, **kw):
    """BASELINE config 4 in miniature: many small polygons (some with holes), scaled-up radii so that random points hit."""
    soup = ShapeSoup()
    for loops in synth.many_polygons(num, seed, **kw):
        soup.add_polygon(loops)
    return soup.freeze()
