"""Reader / writer for the reference's serialized shape index (s2b_index_decode / s2b_index_encode) against the
reference's own MutableS2ShapeIndex::Encode and ::Init (mutable_s2shape_index.cc:1988-2040). Host-only: no GPU needed."""
import os

import numpy as np
import pytest

from s2geometry_b200 import _lib, synth
from s2geometry_b200.index import FlatIndex, build_flat_index, decode_index, encode_index
from s2geometry_b200.shapes import FrozenSoup, ShapeSoup
from tests import scenes
from tests.util import fixture

FLAT_KEYS = ("cell_ids", "cell_center", "cell_clip_begin", "clip_shape_id", "clip_flags", "clip_edge_begin", "edge_id", "edge_v0", "edge_v1")


def mixed_soup(seed, polygons=12):
    """Polygons (some with holes), polylines and point sets in one index: exercises every clipped-shape record kind."""
    rng = np.random.default_rng(seed)
    soup = ShapeSoup()
    for loops in synth.many_polygons(polygons, seed, min_vertices=4, max_vertices=90, min_radius_km=50.0, max_radius_km=900.0):
        soup.add_polygon(loops)
        c = loops[0].mean(axis=0)
        if rng.random() < 0.5:    # a polyline wandering across the polygon
            k = int(rng.integers(2, 60))
            w = c + 0.05 * np.cumsum(rng.standard_normal((k, 3)), axis=0) / np.sqrt(k)
            soup.add_polyline(w / np.linalg.norm(w, axis=1)[:, None])
        if rng.random() < 0.5:    # points, some coinciding with polygon vertices
            pts = np.concatenate([loops[0][:: max(1, len(loops[0]) // 5)], synth.cap_points(c / np.linalg.norm(c), 0.05, int(rng.integers(1, 40)), seed + 5)])
            soup.add_points(pts)
    soup.add_polygon([])          # the empty polygon and the full polygon (one empty loop) are legal lax polygons
    return soup.freeze()


def wire_scenes(small_only=False):
    ka = scenes.known_answers()
    yield "text_index", scenes.text_index(ka["index"]).freeze(), 0
    yield "loop500_m1", scenes.loop_scene(500)[0].freeze(), 1         # 0 or 1 edge per cell
    yield "mixed7_m3", mixed_soup(7, 6), 3
    if small_only:
        return
    yield "loop10k", scenes.loop_scene(10000)[0].freeze(), 0          # contiguous edge runs (the 1-byte records)
    yield "loop3000_m40", scenes.loop_scene(3000)[0].freeze(), 40     # n > 17: general records with run-length edge lists
    yield "axis_aligned", scenes.axis_aligned_scene().freeze(), 0
    yield "polygons300", scenes.polygons_scene(300, 11), 0
    yield "polygons80_m2", scenes.polygons_scene(80, 12), 2
    yield "mixed21", mixed_soup(21, 40), 0
    yield "mixed22_m64", mixed_soup(22, 40), 64
    empty = ShapeSoup()
    yield "empty", empty.freeze(), 0
    pts = ShapeSoup(); pts.add_points(synth.sphere_points(300, 9))
    yield "points_only", pts.freeze(), 0


def assert_flat_equal(got, want, name):
    for k in FLAT_KEYS:
        g = np.asarray(getattr(got, k)).reshape(-1)
        w = np.asarray(want[k]).reshape(-1)
        assert g.shape == w.shape, (name, k, g.shape, w.shape)
        assert (g.view(np.uint8) == w.astype(g.dtype).view(np.uint8)).all(), (name, k)   # bit-exact, incl. doubles


@pytest.mark.parametrize("case", list(wire_scenes()), ids=lambda c: c[0])
def test_decode_matches_reference_and_encode_is_byte_exact(ref, case):
    name, soup, mepc = case
    ri = ref.index_create(soup, mepc)
    data = ri.encode()
    flat, got_mepc, used = decode_index(data, soup)
    assert used == len(data) and got_mepc == (mepc if mepc > 0 else 10)
    assert flat.num_shape_ids == soup.num_shapes
    assert_flat_equal(flat, ri.flatten(), name)
    again = encode_index(flat, mepc)
    assert again == data, name                      # same bytes as MutableS2ShapeIndex::Encode
    assert ri.init_matches(again)                   # and the reference's own Init reads them back to the same cells
    # trailing bytes after the index are left unread (the reference's Decoder semantics)
    flat2, _, used2 = decode_index(data + b"\x01\x02\x03", soup)
    assert used2 == len(data) and (flat2.cell_ids == flat.cell_ids).all()


def test_reference_reads_indexes_written_from_our_builder(ref):
    """An index built by the library's own host builder, serialized, is accepted by MutableS2ShapeIndex::Init; on these
    scenes the builder produces the reference's cells, so Init's result equals the reference-built index."""
    for name, soup, mepc in [("loop", scenes.loop_scene(2000)[0].freeze(), 0), ("polygons", scenes.polygons_scene(120, 5), 0),
                             ("mixed", mixed_soup(3, 20), 0)]:
        own = build_flat_index(soup, mepc)
        data = encode_index(own, mepc)
        ri = ref.index_create(soup, mepc)
        assert ri.init_matches(data), name
        back, _, _ = decode_index(data, soup)
        assert (back.cell_ids == own.cell_ids).all() and (back.edge_id == own.edge_id).all() and (back.clip_flags == own.clip_flags).all()


def test_golden_wire_fixture():
    """Committed reference output (tools/gen_golden.py::wire_fixture): runs without the reference library."""
    z = fixture("index_wire.npz")
    names = sorted({k.split("__")[0] for k in z.files})
    assert names
    for name in names:
        class S:   # FrozenSoup-like
            pass
        soup = S()
        for k in ("shape_dim", "shape_chain_begin", "chain_vertex_begin", "vertices"):
            setattr(soup, k, z[name + "__soup_" + k])
        soup.num_shapes = len(soup.shape_dim)
        soup.freeze = lambda s=soup: s
        data = z[name + "__bytes"].tobytes()
        flat, mepc, used = decode_index(data, soup)
        assert used == len(data) and mepc == int(z[name + "__max_edges_per_cell"])
        assert_flat_equal(flat, {k: z[name + "__" + k] for k in FLAT_KEYS}, name)
        assert encode_index(flat, mepc) == data


def test_malformed_input_is_rejected_not_crashed(ref):
    soup = mixed_soup(4, 10)
    data = ref.index_create(soup).encode()
    for cut in list(range(0, min(len(data), 200))) + list(range(200, len(data), 97)):
        with pytest.raises(_lib.S2BError):
            decode_index(data[:cut], soup)          # every proper prefix is truncated somewhere
    rng = np.random.default_rng(1)
    ok = bad = 0
    for _ in range(3000):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        try:
            flat, _, _ = decode_index(bytes(b), soup)
            ok += 1                                 # still a well-formed index: every offset array must be consistent
            assert flat.cell_clip_begin[-1] == flat.num_clips and flat.clip_edge_begin[-1] == flat.num_edges
            assert (np.diff(flat.cell_ids.astype(np.uint64)) > 0).all()
        except _lib.S2BError:
            bad += 1
    assert bad > 0 and ok + bad == 3000
    with pytest.raises(_lib.S2BError):              # wrong geometry: edge ids beyond the soup's shapes
        small = ShapeSoup(); small.add_polygon([synth.wobbly_loop(np.array([0, 0, 1.0]), 0.1, 5)])
        decode_index(ref.index_create(scenes.loop_scene(500)[0].freeze()).encode(), small)
    v1 = bytearray(data); v1[0] = (v1[0] & ~3) | 1   # version 1
    with pytest.raises(_lib.S2BError):
        decode_index(bytes(v1), soup)
