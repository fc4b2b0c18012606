"""Reader / writer for the reference's serialized shapes (s2b_shapes_decode / s2b_shapes_encode) against the reference's
own S2PointVectorShape / S2LaxPolylineShape / S2LaxPolygonShape encoders and decoders, in both EncodedS2PointVector
formats (UNCOMPRESSED = FAST hint, CELL_IDS = COMPACT hint). Host-only: no GPU needed."""
import numpy as np
import pytest

from s2geometry_b200 import _lib, synth
from s2geometry_b200.index import build_flat_index, decode_index
from s2geometry_b200.shapes import ShapeSoup, decode_shapes, encode_shapes
from tests import scenes
from tests.test_index_wire import mixed_soup


def snap(ref, pts, level):
    """S2CellId(p).parent(level).ToPoint(): vertices at cell centres are what the CELL_IDS format stores compactly."""
    ids = ref.parent(ref.encode_points(pts, 30, 1), level)
    return ref.to_point(ids)


def edges_of(soup):
    """All edges of a soup in (shape, edge id) order as an (E,6) array + per-shape counts (same rule as the C++ side)."""
    out, counts = [], []
    for s in range(soup.num_shapes):
        n0 = len(out)
        for c in range(soup.shape_chain_begin[s], soup.shape_chain_begin[s + 1]):
            v = soup.vertices[soup.chain_vertex_begin[c]: soup.chain_vertex_begin[c + 1]]
            n = len(v)
            d = soup.shape_dim[s]
            for k in range(n if d != 1 else max(0, n - 1)):
                out.append(np.concatenate([v[k], v[k] if d == 0 else v[(k + 1) % n]]))
        counts.append(len(out) - n0)
    return (np.array(out).reshape(-1, 6), np.array(counts, np.uint32))


def single_chain_polylines(soup):
    return all(soup.shape_chain_begin[s + 1] - soup.shape_chain_begin[s] <= 1 for s in range(soup.num_shapes) if soup.shape_dim[s] == 1)


def wire_soups(ref):
    yield "mixed", mixed_soup(31, 25)
    yield "loop10k", scenes.loop_scene(10000)[0]
    yield "polygons", scenes.polygons_scene(200, 17)
    # vertices snapped to cell centres (one level, so COMPACT stores cell ids) + a few unsnapped ones (exceptions)
    for level in (30, 18, 7):
        s = ShapeSoup()
        rng = np.random.default_rng(level)
        for loops in synth.many_polygons(30, 40 + level, min_vertices=3, max_vertices=70, min_radius_km=200.0, max_radius_km=900.0):
            s.add_polygon([snap(ref, l, level) for l in loops])
        line = snap(ref, synth.cap_points(np.array([0.0, 1.0, 0.0]), 0.2, 45, 3), level)
        line[7] = synth.sphere_points(1, 5)[0]                                  # an exception inside a block
        s.add_polyline(line)
        s.add_points(snap(ref, synth.sphere_points(33, 9), level))
        s.add_points(np.concatenate([snap(ref, synth.sphere_points(16, 10), level), synth.sphere_points(5, 11)]))
        yield "snapped%d" % level, s.freeze()
    e = ShapeSoup(); e.add_polygon([]); e.add_polygon([np.zeros((0, 3))]); e.add_points(np.zeros((0, 3))); e.add_polyline(np.zeros((0, 3)))
    e.add_polyline(synth.sphere_points(1, 2)); e.add_polygon([synth.sphere_points(1, 3)])
    yield "degenerate", e.freeze()
    yield "empty", ShapeSoup().freeze()


def test_decode_and_encode_match_reference(ref):
    for name, soup in wire_soups(ref):
        want_edges, want_counts = edges_of(soup)
        for compact in (False, True):
            data = ref.tagged_shapes_encode(soup, compact)
            got, used = decode_shapes(data)
            assert used == len(data), name
            rd = ref.tagged_shapes_decode_edges(data)
            assert rd is not None, name
            dims, ne, redges = rd
            ge, gc = edges_of(got)
            assert got.num_shapes == soup.num_shapes and (got.shape_dim == dims).all(), (name, compact)
            assert (gc == ne).all() and ge.shape == redges.shape, (name, compact)
            assert (ge.view(np.uint64) == redges.view(np.uint64)).all(), (name, compact)     # bit-identical to the reference's decoder
            if not compact:   # lossless: exactly the input geometry
                assert (ge.view(np.uint64) == want_edges.view(np.uint64)).all() and (gc == want_counts).all(), name
                assert (got.shape_chain_begin == soup.shape_chain_begin).all() and (got.chain_vertex_begin == soup.chain_vertex_begin).all()
        if single_chain_polylines(soup):
            assert encode_shapes(soup) == ref.tagged_shapes_encode(soup, False), name           # byte-identical FAST encoding
            assert encode_shapes(soup, compact=True) == ref.tagged_shapes_encode(soup, True), name   # ... and COMPACT encoding
    # the compact form really exercised the CELL_IDS format: much smaller than 24 bytes per vertex
    soup = [s for n, s in wire_soups(ref) if n == "snapped18"][0]
    assert len(ref.tagged_shapes_encode(soup, True)) < 0.4 * len(ref.tagged_shapes_encode(soup, False))


def test_serialized_index_plus_shapes_round_trip(ref):
    """What a reference user ships: shapes (compact) followed by the index. Decoding both gives a device-ready flat
    index identical to flattening the reference's own index over the decoded geometry."""
    from tests.test_index_wire import FLAT_KEYS, assert_flat_equal
    soup = mixed_soup(41, 30)
    blob = ref.tagged_shapes_encode(soup, False) + ref.index_create(soup).encode()
    shapes, used = decode_shapes(blob)
    flat, mepc, used2 = decode_index(blob[used:], shapes)
    assert used + used2 == len(blob) and mepc == 10
    assert_flat_equal(flat, ref.index_create(soup).flatten(), "mixed41")
    own = build_flat_index(shapes)
    assert (own.cell_ids == flat.cell_ids).all()


def test_malformed_shapes_are_rejected_not_crashed(ref):
    soup = [s for n, s in wire_soups(ref) if n == "snapped18"][0]
    csoup = [s for n, s in classic_soups(ref) if n == "snapped20"][0]
    for compact, classic in ((False, False), (True, False), (False, True), (True, True)):
        data = ref.tagged_shapes_encode(csoup if classic else soup, compact, classic=classic)
        for cut in list(range(0, 64)) + list(range(64, len(data), 131)):
            with pytest.raises(_lib.S2BError):
                decode_shapes(data[:cut])
        rng = np.random.default_rng(3 + compact + 2 * classic)
        bad = 0
        for _ in range(1500):
            b = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, min(len(b), 4000)))] = int(rng.integers(0, 256))
            try:
                got, _ = decode_shapes(bytes(b))
                assert got.chain_vertex_begin[-1] == len(got.vertices)
            except _lib.S2BError:
                bad += 1
        assert bad > 0
    with pytest.raises(_lib.S2BError):          # an unknown type tag
        decode_shapes(bytes([8, 2, 9, 0]))


def classic_soups(ref):
    """Scenes for S2Polygon::Shape / S2Polyline::Shape: valid nested polygons (holes, several shells), polylines, snapped
    (compressed encodings: S2EncodePointsCompressed) and unsnapped (lossless), with a few off-centre vertices."""
    yield "polygons", scenes.polygons_scene(60, 23)
    for level in (30, 20, 9):
        s = ShapeSoup()
        for loops in synth.many_polygons(25, 70 + level, min_vertices=4, max_vertices=90, min_radius_km=300.0, max_radius_km=1200.0):
            s.add_polygon([snap(ref, l, level) for l in loops])
        line = snap(ref, synth.cap_points(np.array([0.3, 0.2, 0.9]), 0.3, 120, 4), level)
        s.add_polyline(line)
        mixed = line.copy(); mixed[5] = synth.sphere_points(1, 6)[0]; mixed[77] = synth.sphere_points(1, 8)[0]     # off-centre points
        s.add_polyline(mixed)
        big = snap(ref, synth.wobbly_loop(np.array([-0.2, 0.7, 0.1]), 0.4, 200), level)                              # >= 64 vertices: bound encoded
        s.add_polygon([big])
        yield "snapped%d" % level, s.freeze()
    e = ShapeSoup(); e.add_polygon([]); e.add_polyline(np.zeros((0, 3))); e.add_polygon([np.zeros((0, 3))]); e.add_polyline(synth.sphere_points(2, 2))
    yield "empty_and_full", e.freeze()


def test_s2polygon_and_s2polyline_shapes_decode_like_the_reference(ref):
    """Tags 1 and 2 (S2Polygon::Shape, S2Polyline::Shape), both coding hints: the decoded vertex chains give exactly the
    edges the reference's own OwningShape::Init gives (oriented: holes reversed), and the lossless form reproduces the
    input rings up to the reference's own loop reordering (compared as edge sets per shape)."""
    saw_compressed = 0
    for name, soup in classic_soups(ref):
        for compact in (False, True):
            data = ref.tagged_shapes_encode(soup, compact, classic=True)
            lax = ref.tagged_shapes_encode(soup, False)
            got, used = decode_shapes(data)
            assert used == len(data), name
            dims, ne, redges = ref.tagged_shapes_decode_edges(data)
            ge, gc = edges_of(got)
            assert got.num_shapes == soup.num_shapes and (got.shape_dim == dims).all(), (name, compact)
            assert (gc == ne).all(), (name, compact, gc, ne)
            assert (ge.view(np.uint64) == redges.view(np.uint64)).all(), (name, compact)     # bit-identical, same edge order
            if compact and len(data) < 0.5 * len(lax):
                saw_compressed += 1
            if not compact:      # lossless: the same set of directed edges as the input soup, shape by shape
                we, wc = edges_of(soup)
                assert (gc == wc).all(), name
                o = 0
                for n in gc:
                    a = sorted(map(bytes, ge[o:o + n].view(np.uint8).reshape(n, 48))) if n else []
                    b = sorted(map(bytes, we[o:o + n].view(np.uint8).reshape(n, 48))) if n else []
                    assert a == b, name
                    o += n
            # and the decoded geometry indexes to the same answers as the reference's index over the original shapes
        flat = build_flat_index(decode_shapes(ref.tagged_shapes_encode(soup, False, classic=True))[0])
        assert flat.num_shape_ids == soup.num_shapes
    assert saw_compressed >= 2      # the snapped scenes really used S2EncodePointsCompressed


def test_golden_shapes_fixture():
    """Committed reference output (tools/gen_golden.py::shapes_wire_fixture): runs without the reference library."""
    from tests.util import fixture
    z = fixture("shapes_wire.npz")
    keys = sorted({k.split("__")[0] for k in z.files})
    assert len(keys) == 6 and sum(k.startswith("classic") for k in keys) == 2
    for key in keys:
        data = z[key + "__bytes"].tobytes()
        got, used = decode_shapes(data)
        assert used == len(data)
        ge, gc = edges_of(got)
        assert (got.shape_dim == z[key + "__dims"]).all() and (gc == z[key + "__num_edges"]).all()
        assert (ge.view(np.uint64) == z[key + "__edges"].reshape(-1, 6).view(np.uint64)).all()
        if key.endswith("_fast") and not key.startswith("classic"):
            assert encode_shapes(got) == data


def test_compact_point_vector_encoder_is_byte_identical_on_adversarial_vectors(ref):
    """EncodeS2PointVectorCompact's parameter choices (level, base, per-block delta / offset / overlap widths, exceptions)
    on point vectors built to reach its corners: single points, one block / many blocks / a last block of one value,
    blocks made only of exceptions, values spread over several faces (wide deltas, long offsets), tight clusters (short
    deltas, long base), mixed levels (the majority level wins, the rest become exceptions), and too few cell centres
    (falls back to the uncompressed form). Every vector must come out byte for byte as the reference writes it."""
    rng = np.random.default_rng(2024)
    vectors = []
    for level in (0, 1, 2, 5, 11, 15, 23, 29, 30):
        for n in (1, 2, 15, 16, 17, 31, 33, 100):
            for spread in (3.0, 0.3, 1e-2, 1e-5, 1e-8):
                c = synth.sphere_points(1, int(rng.integers(1 << 30)))[0]
                raw = c + spread * rng.standard_normal((n, 3))
                raw /= np.linalg.norm(raw, axis=1)[:, None]
                pts = snap(ref, raw, level)
                kind = int(rng.integers(5))
                if kind == 1 and n > 2:                       # a few exceptions
                    k = rng.choice(n, max(1, n // 7), replace=False)
                    pts[k] = synth.sphere_points(len(k), int(rng.integers(1 << 30)))
                elif kind == 2 and n >= 32:                   # one block of exceptions only
                    pts[16:32] = synth.sphere_points(16, int(rng.integers(1 << 30)))
                elif kind == 3 and n > 4:                     # a minority at another level
                    k = rng.choice(n, n // 3, replace=False)
                    pts[k] = snap(ref, raw[k], max(0, level - 3))
                elif kind == 4 and n >= 40:                   # hardly any cell centres: < 5 % -> uncompressed
                    pts[2:] = synth.sphere_points(n - 2, int(rng.integers(1 << 30)))
                vectors.append(pts)
    # block minima / maxima placed around the nibble and byte boundaries GetBlockCode's comment works through
    from oracle import restatement as rs

    def deinterleave(v):          # inverse of InterleaveUint32BitPairs: bit pairs alternate between sj and tj
        sj = tj = 0
        for k in range(16):
            sj |= ((v >> (4 * k)) & 3) << (2 * k)
            tj |= ((v >> (4 * k + 2)) & 3) << (2 * k)
        return sj, tj

    level = 24
    for lo, hi in ((0x72, 0x7e), (0x78, 0x84), (0x08, 0x104), (0xf08, 0x1004), (0, 0xffff), (0xfff0, 0x1000f), (0x3, 0x3)):
        for base in (0, 0x5a5a5a5a5a5 << 8):
            vals = np.unique(np.linspace(lo, hi, 16).astype(np.int64)) + base
            st = np.array([deinterleave(int(v)) for v in vals], np.int64)      # face 0: sj = i, tj = j at `level`
            leaf = rs.from_face_ij(np.zeros(len(vals), np.int64), st[:, 0] << (30 - level), st[:, 1] << (30 - level))
            vectors.append(ref.to_point(ref.parent(leaf, level)))
    checked = compressed = 0
    for pts in vectors:
        s = ShapeSoup(); s.add_points(pts); s = s.freeze()
        want = ref.tagged_shapes_encode(s, True)
        got = encode_shapes(s, compact=True)
        assert got == want, (len(pts), checked)
        compressed += len(want) < 20 * len(pts)
        checked += 1
    assert checked >= 370 and compressed > 250
    # polylines and polygons take the same path
    for level in (30, 12):
        s = ShapeSoup()
        for loops in synth.many_polygons(20, 5 + level, min_vertices=3, max_vertices=90, min_radius_km=0.01, max_radius_km=3000.0):
            s.add_polygon([snap(ref, l, level) for l in loops])
        s.add_polyline(snap(ref, synth.cap_points(np.array([0.0, 0.6, 0.8]), 1e-4, 77, 3), level))
        s = s.freeze()
        assert encode_shapes(s, compact=True) == ref.tagged_shapes_encode(s, True)
