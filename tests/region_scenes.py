"""Target cells for the S2ShapeIndexRegion cell tests, shared by the CPU (hostsim) and GPU tests."""
import numpy as np

from s2geometry_b200 import synth
from s2geometry_b200.shapes import ShapeSoup
from tests import scenes


def big_loop_scene():
    """Two loops whose long edges run across several cube faces (so ClipToPaddedFace leaves its same-face fast path),
    a polyline and a few points."""
    soup = ShapeSoup()
    soup.add_polygon([synth.wobbly_loop(np.array([0.6, 0.6, 0.5]), 0.9, 48)])
    soup.add_polygon([synth.wobbly_loop(np.array([-0.7, 0.1, -0.7]), 0.5, 24)])
    t = np.linspace(-2.0, 2.0, 12)
    soup.add_polyline(np.stack([np.cos(t), np.sin(t), 0.3 * np.sin(3 * t)], 1) / np.sqrt(1 + (0.3 * np.sin(3 * t)) ** 2)[:, None])
    soup.add_points(scenes._norm(np.array([[1, 1, 1], [1, 0, 0], [-1, -1, 0.5]], float)))
    return soup.freeze()


def region_targets(ref, flat_cell_ids, vertices, seed, n_random=3000):
    """Cells in every relation to the index: the index cells, their descendants (INDEXED, with and without edges
    nearby), their ancestors (SUBDIVIDED), cells around the shapes' vertices at levels 3..30 and random cells."""
    rng = np.random.default_rng(seed)
    cell_ids = np.asarray(flat_cell_ids, np.uint64)
    out = [cell_ids[:: max(1, len(cell_ids) // 1500)]]
    # descendants / ancestors of index cells: re-encode a point of the cell at another level
    pick = cell_ids[rng.integers(0, len(cell_ids), 3000)]
    centres = ref.to_point(pick)
    jitter = scenes._norm(centres + 1e-4 * rng.standard_normal(centres.shape))
    for pts in (centres, jitter):
        leaf = ref.encode_points(pts, 30, 1)
        lv = rng.integers(0, 31, len(leaf))
        out.append(np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(lv)], np.uint64))
    # around the vertices (the cells coverers ask about when they refine along a boundary)
    v = np.asarray(vertices, np.float64)
    v = v[rng.integers(0, len(v), 3000)]
    near = scenes._norm(v + rng.choice([0.0, 1e-9, 1e-6, 1e-3], len(v))[:, None] * rng.standard_normal(v.shape))
    leaf = ref.encode_points(near, 30, 1)
    lv = rng.integers(3, 31, len(leaf))
    out.append(np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(lv)], np.uint64))
    # random cells
    leaf = ref.encode_points(synth.sphere_points(n_random, seed + 1), 30, 1)
    lv = rng.integers(0, 31, len(leaf))
    out.append(np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(lv)], np.uint64))
    out.append(np.array([ref.parent(leaf[:1], 0)[0]], np.uint64))
    return np.ascontiguousarray(np.concatenate(out))


def region_scene_list():
    loop, _ = scenes.loop_scene(1500)
    return [("loop", loop), ("polygons", scenes.polygons_scene(150, 11)), ("axis", scenes.axis_aligned_scene()), ("big", big_loop_scene())]
