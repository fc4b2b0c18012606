#!/bin/bash
# Builds oracle/_ref/libs2ref.so: the UNMODIFIED reference sources (compiled where they lie under
# /root/reference/src) against the header-only abseil stand-in in oracle/absl_shim, plus oracle/ref_capi.cc.
# Test infrastructure only — never linked into the product. Outputs go to oracle/_ref/ (git-ignored).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${S2_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
[ -d "$REF/s2" ] || { echo "reference sources not found at $REF (expected on the GPU box: use prebuilt $OUT/libs2ref.so)"; exit 0; }
mkdir -p "$OUT/obj"
TUS="s2cell_id s2coords s2latlng s1angle s1chord_angle s1interval r2rect s2error s2debug s2metrics s2pointutil s2predicates
s2edge_crosser s2edge_crossings s2edge_clipping s2edge_distances s2measures s2padded_cell s2cell s2cap s2latlng_rect
s2cell_union s2shape_index mutable_s2shape_index s2memory_tracker s2shapeutil_contains_brute_force
s2shapeutil_get_reference_point s2contains_vertex_query encoded_s2cell_id_vector encoded_string_vector
util/coding/coder util/coding/varint util/math/mathutil util/math/exactfloat/exactfloat util/math/exactfloat/bignum
base/malloc_extension s2region_coverer s2closest_point_query s2min_distance_targets s2closest_edge_query
s2closest_cell_query s2cell_index s2region_sharder s2lax_polygon_shape s2lax_polyline_shape encoded_s2point_vector
s2loop s2polygon s2polyline s2loop_measures s2polyline_measures s2latlng_rect_bounder s2crossing_edge_query
s2wedge_relations s2point_compression s2centroids s2shapeutil_edge_wrap s2shapeutil_visit_crossing_edge_pairs
internal/s2index_cell_data internal/s2incident_edge_tracker util/bits/bit-interleave"
# -fvisibility=hidden + --gc-sections: only what ref_capi.cc's exported functions reach is linked, so S2Polygon / S2Loop /
# S2Polyline (serialization, containment) come in without S2Builder and the boolean operations they can also call.
CXXFLAGS="-std=c++20 -O2 -DNDEBUG -w -fPIC -ffunction-sections -fdata-sections -fvisibility=hidden -I$HERE/absl_shim -I$REF"
pids=()
for t in $TUS; do
  o="$OUT/obj/$(echo $t | tr / _).o"
  if [ ! -f "$o" ] || [ "$REF/s2/$t.cc" -nt "$o" ]; then
    ( g++ $CXXFLAGS -c "$REF/s2/$t.cc" -o "$o" || { echo "FAILED $t"; rm -f "$o"; } ) &
    pids+=($!)
    if [ ${#pids[@]} -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  fi
done
wait
if [ -f "$HERE/ref_capi.cc" ]; then
  g++ $CXXFLAGS -pthread -c "$HERE/ref_capi.cc" -o "$OUT/obj/ref_capi.o"
  g++ -shared -pthread -Wl,--gc-sections -Wl,--no-undefined -o "$OUT/libs2ref.so" "$OUT"/obj/*.o
  echo "built $OUT/libs2ref.so"
fi
