#!/usr/bin/env python3
"""Generates the reference-derived golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libs2ref.so, built from /root/reference/src). /root/reference does not exist on the GPU box, so the
outputs are committed and this script documents how they were made.  Run from the repo root:
    python tools/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402
from s2geometry_b200 import synth  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def encode_fixture(ref):
    n = 20000
    pts = synth.sphere_points(n, 101)
    # add adversarial rows: axis points, face ties, zeros, negative zero, NaN, tiny and huge magnitudes
    extra = np.array([
        [1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1],
        [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1], [-1, -1, -1], [1, -1, 1], [-1, 1, 1],
        [0, 0, 0], [-0.0, 0, 0], [0, -0.0, 1], [np.nan, 1, 1], [1, np.nan, 0], [1e-300, 1e-300, 1e-300],
        [1e300, 1e-300, 1], [5e-324, 0, 0], [1, 1 - 2 ** -53, 0], [1 - 2 ** -53, 1, 1], [0.5, 0.5, 0.5 + 2 ** -54],
        [1, 2, 3], [3, 2, 1], [-3, 2, -1], [np.inf, 1, 1], [1, 1 / 3.0, -1 / 3.0], [1, -1 / 3.0, 1 / 3.0],
    ], dtype=np.float64)
    pts = np.concatenate([pts, extra])
    ll = synth.sphere_latlng(n, 102)
    ll_extra = np.array([[0, 0], [0, np.pi / 2], [np.pi / 2, 0], [0, np.pi], [0, -np.pi / 2], [-np.pi / 2, 0], [0, -np.pi],
                         [np.pi / 4, np.pi / 4], [-np.pi / 4, 3 * np.pi / 4], [1e-300, -1e-300], [0.6154797086703873, np.pi / 4],
                         [np.radians(39.0), np.radians(-120.0)], [np.radians(63.431052), np.radians(10.395083)]])
    ll = np.concatenate([ll, ll_extra])
    e7 = np.random.default_rng(103).integers([-900000000, -1800000000], [900000001, 1800000001], size=(5000, 2)).astype(np.int32)
    out = {"xyz": pts, "latlng": ll, "e7": e7}
    for lvl in (30, 20, 12, 0):
        out["xyz_ids_L%d" % lvl] = ref.encode_points(pts, lvl, 1)
        out["latlng_ids_L%d" % lvl] = ref.encode_latlng(ll, lvl, 1)
    out["e7_ids_L30"] = ref.encode_latlng_e7(e7, 30, 1)
    # random ids at all levels + tokens
    rng = np.random.default_rng(104)
    leaf = ref.encode_points(synth.sphere_points(4000, 105), 30, 1)
    ids = np.array([ref.parent(leaf[i:i + 1], int(l))[0] for i, l in enumerate(rng.integers(0, 31, len(leaf)))], dtype=np.uint64)
    ids = np.concatenate([ids, np.array([0, 0xFFFFFFFFFFFFFFFF, 0xF000000000000000, 1, 0x1000000000000000], dtype=np.uint64)])
    tok, ln = ref.to_token(ids)
    out["tok_ids"] = ids; out["tok"] = tok; out["tok_len"] = ln
    np.savez_compressed(os.path.join(G, "encode_fixture.npz"), **out)
    print("encode_fixture.npz", {k: v.shape for k, v in out.items()})


def index_fixture(ref):
    """Flat index of the 10k-vertex loop as built and iterated by the reference + reference answers for a fixed query set."""
    sys.path.insert(0, os.path.join(ROOT))
    from tests import scenes
    soup, v = scenes.loop_scene(10000)
    ri = ref.index_create(soup)
    f = ri.flatten()
    u, b, d = scenes.loop_points(v, 20000, 40000, 77)
    q = np.concatenate([u, b, d])
    out = {k: f[k] for k in ("cell_ids", "cell_center", "cell_clip_begin", "clip_shape_id", "clip_flags", "clip_edge_begin", "edge_v0", "edge_v1")}
    out["num_shape_ids"] = np.int32(f["num_shape_ids"])
    out["loop_vertices"] = v
    out["queries"] = q
    for name, m in scenes.MODELS.items():
        out["contains_" + name] = ri.contains(q, m, 1)
    out["brute_force"] = ri.brute_force_contains(0, q, 0)
    np.savez_compressed(os.path.join(G, "index_loop10k.npz"), **out)
    print("index_loop10k.npz cells=%d clips=%d edges=%d queries=%d inside=%d" % (
        len(f["cell_ids"]), len(f["clip_shape_id"]), len(f["edge_v0"]), len(q), out["contains_SEMI_OPEN"].sum()))


def wire_fixture(ref):
    """Bytes of MutableS2ShapeIndex::Encode for two small scenes (single-shape and multi-shape with points and polylines)
    + the reference's own flattening of the same indexes: pins s2b_index_decode / s2b_index_encode."""
    from tests import scenes
    from tests.test_index_wire import wire_scenes
    out = {}
    for name, soup, mepc in wire_scenes(small_only=True):
        ri = ref.index_create(soup, mepc)
        f = ri.flatten()
        out[name + "__bytes"] = np.frombuffer(ri.encode(), np.uint8)
        out[name + "__max_edges_per_cell"] = np.int32(mepc if mepc > 0 else 10)
        for k in ("shape_dim", "shape_chain_begin", "chain_vertex_begin", "vertices"):
            out[name + "__soup_" + k] = getattr(soup, k)
        for k in ("cell_ids", "cell_center", "cell_clip_begin", "clip_shape_id", "clip_flags", "clip_edge_begin", "edge_id", "edge_v0", "edge_v1"):
            out[name + "__" + k] = f[k]
        print("index_wire", name, "bytes=%d cells=%d" % (len(out[name + "__bytes"]), len(f["cell_ids"])))
    np.savez_compressed(os.path.join(G, "index_wire.npz"), **out)


def shapes_wire_fixture(ref):
    """Tagged-shape bytes (FAST and COMPACT hints) from the reference's own lax shape encoders for a small snapped scene +
    the edges the reference's own decoders return: pins s2b_shapes_decode / s2b_shapes_encode without the library."""
    from tests.test_shapes_wire import wire_soups
    out = {}
    for name, soup in wire_soups(ref):
        if name not in ("snapped18", "degenerate"):
            continue
        for compact in (False, True):
            data = ref.tagged_shapes_encode(soup, compact)
            dims, ne, edges = ref.tagged_shapes_decode_edges(data)
            key = "%s_%s" % (name, "compact" if compact else "fast")
            out[key + "__bytes"] = np.frombuffer(data, np.uint8)
            out[key + "__dims"] = dims; out[key + "__num_edges"] = ne; out[key + "__edges"] = edges
            print("shapes_wire", key, "bytes=%d edges=%d" % (len(data), len(edges)))
    # S2Polygon::Shape / S2Polyline::Shape (tags 1, 2) in their lossless and compressed encodings
    from tests.test_shapes_wire import classic_soups
    for name, soup in classic_soups(ref):
        if name != "snapped20":
            continue
        for compact in (False, True):
            data = ref.tagged_shapes_encode(soup, compact, classic=True)
            dims, ne, edges = ref.tagged_shapes_decode_edges(data)
            key = "classic%s_%s" % (name, "compact" if compact else "fast")
            out[key + "__bytes"] = np.frombuffer(data, np.uint8)
            out[key + "__dims"] = dims; out[key + "__num_edges"] = ne; out[key + "__edges"] = edges
            print("shapes_wire", key, "bytes=%d edges=%d" % (len(data), len(edges)))
    np.savez_compressed(os.path.join(G, "shapes_wire.npz"), **out)


def covering_fixture(ref):
    """BASELINE configs[4]: the S2RegionCoverer covering (max_cells = 1000, default levels) of the 10k-vertex loop, computed
    by the reference's own coverer over S2ShapeIndexRegion(MutableS2ShapeIndex{loop}); + the reference's answers for a fixed
    query set (S2CellUnion::Contains(S2Point))."""
    from tests import scenes
    soup, v = scenes.loop_scene(10000)
    ri = ref.index_create(soup)
    ids = ri.covering(1000)
    u, b, d = scenes.loop_points(v, 20000, 40000, 78)
    q = np.concatenate([u, b, d])
    np.savez_compressed(os.path.join(G, "covering_loop10k.npz"), cell_ids=ids, queries=q, contains=ref.cellunion_contains(ids, q, 1))
    print("covering_loop10k.npz cells=%d queries=%d inside=%d" % (len(ids), len(q), ref.cellunion_contains(ids, q, 1).sum()))


if __name__ == "__main__":
    ref = reflib.load()
    if len(sys.argv) > 1 and sys.argv[1] == "covering":
        covering_fixture(ref)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "shapes":
        shapes_wire_fixture(ref)
        sys.exit(0)
    encode_fixture(ref)
    index_fixture(ref)
    wire_fixture(ref)
    shapes_wire_fixture(ref)
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        pass
