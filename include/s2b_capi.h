/* s2b_capi.h — C ABI of the B200-native S2 hot path (libs2b.so).
 *
 * The reference (google/s2geometry) has no FFI for this path; its seams are C++ value types and
 * templates. Each entry point below is the BATCH form of one reference call and cites
 * the reference interface it replaces (paths relative to the reference's src/s2/). INTEGRATION.md shows
 * the C++ sugar a maintainer adds on the reference side (S2CellId::FromPoints, ...ContainsBatch).
 *
 * Conventions
 *  - plain pointers and sizes only; caller owns every buffer; opaque handles own device memory.
 *  - functions return 0 on success or a negative S2B_E* code and never throw; s2b_last_error() returns a
 *    thread-local message for the last failure on the calling thread.
 *  - un-suffixed functions take HOST pointers (copies are made inside the call, pipelined over PCIe);
 *    `_dev` functions take DEVICE pointers on the current CUDA device plus a cudaStream_t passed as
 *    void* (NULL = legacy default stream) and only enqueue work — no synchronisation.
 *  - results are bit-exact with the reference for cell ids, containment booleans and counts. There is
 *    NO CPU fallback: if no CUDA device / kernel image is usable the call fails with S2B_ECUDA.
 *  - points are S2Point = 3 doubles AoS, 24 bytes (s2point.h:35); lat/lng are 2 doubles AoS in radians
 *    (lat, lng) like S2LatLng's internal R2Point (s2latlng.h:47-60); cell ids are uint64 (S2CellId::id()).
 */
#ifndef S2B_CAPI_H_
#define S2B_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2B_OK 0
#define S2B_EINVAL (-1)  /* bad argument (null pointer, level out of range, ...) */
#define S2B_ECUDA (-2)   /* CUDA runtime / launch failure, or no usable device */
#define S2B_ENOMEM (-3)  /* host or device allocation failed */
#define S2B_ECAPACITY (-4) /* caller-provided output capacity too small; required size is reported */
#define S2B_EUNSUPPORTED (-5) /* the requested mode is not available in this process (e.g. NCCL missing); stated per function */

/* S2VertexModel (s2contains_point_query.h:54) */
#define S2B_VERTEX_MODEL_OPEN 0
#define S2B_VERTEX_MODEL_SEMI_OPEN 1
#define S2B_VERTEX_MODEL_CLOSED 2

const char* s2b_last_error(void);
const char* s2b_version(void);
/* Number of visible CUDA devices (0 if none / driver missing). Never fails. */
int s2b_device_count(void);
/* Kernel launches issued by this library in this process so far (all threads). */
uint64_t s2b_launch_count(void);

/* Diagnostic: measured FP64 FMA issue rate of the current device (DFMA/s; x2 for FLOP/s), from a
 * register-resident DFMA chain. It is the FP64 roofline denominator bench.py reports next to the HBM one. */
int s2b_diag_fp64_peak(double* dfma_per_second_out);

/* Diagnostic: the lat/lng encoders evaluate sin/cos with a fast bounded-error routine and re-evaluate with the
 * correctly rounded one only where the leaf cell could depend on the difference. on != 0 makes every
 * point take the correctly rounded routine, which must give identical ids (tests check it). Returns the old setting. */
int s2b_diag_force_correctly_rounded(int on);

/* Diagnostic: the encoders first evaluate (face, i, j) with bounded-error arithmetic and repeat the evaluation with the
 * exact sequence (IEEE division / square root; for lat/lng also correctly rounded sin/cos) only where the leaf cell
 * could depend on the difference. This call measures what that rests on: the largest distance, in leaf
 * cells, between the two evaluations over n device-resident inputs (input_kind 0: xyz triples, 1: lat/lng radian
 * pairs), and the number of points for which they pick different faces without the filter noticing (must be 0).
 * Synchronises the stream. */
int s2b_diag_filter_deviation_dev(int input_kind, const double* in_dev, size_t n, double* max_deviation_out,
                                  uint64_t* unflagged_face_flips_out, void* stream);

/* Benchmark / test support: counter-based synthetic inputs. Point i of the stream `seed` is a pure
 * function of (seed, first_index + i) built from integer hashing and + - * / sqrt only, so any host (NumPy restatement:
 * s2geometry_b200/synth.py) and any GPU produce the same bits; ranks generate their slice of one global stream.
 *   sphere_points: uniform on the unit sphere (Marsaglia), n x 3 doubles
 *   sphere_latlng: (lat, lng) radians, uniform on the sphere, n x 2 doubles
 *   cap_points:    uniform in the cap of height `height` = 1 - cos(radius) around frame9[6..8]; frame9 = orthonormal x, y, z rows */
int s2b_synth_sphere_points_dev(uint64_t seed, uint64_t first_index, size_t n, double* xyz_out_dev, void* stream);
int s2b_synth_sphere_latlng_dev(uint64_t seed, uint64_t first_index, size_t n, double* latlng_out_dev, void* stream);
int s2b_synth_cap_points_dev(uint64_t seed, uint64_t first_index, size_t n, const double* frame9, double height,
                             double* xyz_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cell-id encoding.  Replaces the per-point constructors
 *   S2CellId::S2CellId(const S2Point&)   s2cell_id.cc:309-315  (XYZtoFaceUV s2coords.h:389-419,
 *                                         UVtoST :329-332, STtoIJ :345-356, FromFaceIJ s2cell_id.cc:267-307)
 *   S2CellId::S2CellId(const S2LatLng&)  s2cell_id.cc:317      (S2LatLng::ToPoint s2latlng.cc:68-77)
 * followed by S2CellId::parent(level)    s2cell_id.h:662-668 when level < 30.
 * level must be in [0,30].
 * ---------------------------------------------------------------------------------------------- */
int s2b_encode_points(const double* xyz, size_t n, int level, uint64_t* ids_out);
int s2b_encode_points_dev(const double* xyz_dev, size_t n, int level, uint64_t* ids_out_dev, void* stream);

/* lat/lng in radians. S2LatLng::ToPoint calls the platform libm's sin and cos (s2latlng.cc:68-77, s1angle.h:343-357), so
 * the reference's id for a point within ~1e-6 leaf cells of a cell boundary depends on which libm ran (glibc is < 1 ulp
 * but not correctly rounded for ~0.13% of arguments). Two modes, process-wide (s2b_set_latlng_mode):
 *   S2B_LATLNG_HOST_LIBM (default)  ids equal the reference's evaluated in the CALLING PROCESS, bit for bit, for every
 *       input: the device evaluates correctly rounded sin/cos and hands back every point whose leaf cell could change
 *       under a <= 1 ulp change of a sin/cos value (~1.6e-5 of uniform input; plus angles with |x| >= 2^20 or non-finite);
 *       exactly those points are re-evaluated with this process's sin/cos before the call returns. Applies to the
 *       host-pointer calls and to s2b_encode_latlng_strict_dev.
 *   S2B_LATLNG_CORRECTLY_ROUNDED    ids are those of correctly rounded sin/cos: platform independent, device only.
 * The plain `_dev` calls only enqueue work and therefore always produce the correctly rounded ids; boundary_flag_out
 * (nullable, all forms): 1 where a ±1-ulp change of a sin/cos result could move the leaf cell, i.e. exactly the points
 * the strict mode re-evaluates. */
#define S2B_LATLNG_HOST_LIBM 0
#define S2B_LATLNG_CORRECTLY_ROUNDED 1
/* Returns the previous mode (or S2B_EINVAL). */
int s2b_set_latlng_mode(int mode);
/* Points re-evaluated on the host by the strict mode, and how many of them changed id, since the last reset. */
int s2b_latlng_strict_stats(uint64_t* reevaluated_out, uint64_t* changed_out, int reset);
/* Device-pointer encode with the strict guarantee: input_kind 1 = (lat, lng) radian pairs, 2 = E7 int32 pairs; outputs
 * as s2b_encode_latlng_levels_tokens_dev (boundary_flag_out_dev, tokens16_out_dev / len_out_dev nullable; no token when
 * token_level_index < 0). In S2B_LATLNG_HOST_LIBM mode it SYNCHRONISES `stream`: kernel -> list of sensitive points
 * (32 bytes each) to the host -> host sin/cos -> the ids that differ are patched on the device. */
int s2b_encode_latlng_strict_dev(int input_kind, const void* latlng_dev, size_t n, const int* levels, int num_levels,
                                 uint64_t* ids_out_dev, uint8_t* boundary_flag_out_dev, int token_level_index,
                                 char* tokens16_out_dev, uint8_t* len_out_dev, void* stream);
int s2b_encode_latlng(const double* latlng_rad, size_t n, int level, uint64_t* ids_out, uint8_t* boundary_flag_out);
int s2b_encode_latlng_dev(const double* latlng_rad_dev, size_t n, int level, uint64_t* ids_out_dev,
                          uint8_t* boundary_flag_out_dev, void* stream);
/* E7 fixed point: S2LatLng::FromE7 = S1Angle::E7 = Degrees(1e-7 * v) (s1angle.h:367-377). */
int s2b_encode_latlng_e7(const int32_t* latlng_e7, size_t n, int level, uint64_t* ids_out, uint8_t* boundary_flag_out);
int s2b_encode_latlng_e7_dev(const int32_t* latlng_e7_dev, size_t n, int level, uint64_t* ids_out_dev,
                             uint8_t* boundary_flag_out_dev, void* stream);

/* Fused multi-level encode (BASELINE config 2: ids at levels 12/20/30 in one pass): ids_out is
 * num_levels arrays of n ids, array k at ids_out + k*n, for levels[k]. num_levels in [1,4]. */
int s2b_encode_latlng_levels(const double* latlng_rad, size_t n, const int* levels, int num_levels, uint64_t* ids_out);
int s2b_encode_latlng_levels_dev(const double* latlng_rad_dev, size_t n, const int* levels, int num_levels,
                                 uint64_t* ids_out_dev, void* stream);
int s2b_encode_points_levels_dev(const double* xyz_dev, size_t n, const int* levels, int num_levels,
                                 uint64_t* ids_out_dev, void* stream);

/* As s2b_encode_latlng_levels, and additionally S2CellId::ToToken (s2cell_id.cc:217-233) of the id at
 * levels[token_level_index], in the same pass (no second read of the ids): tokens16_out is n 16-byte slots,
 * NUL padded; len_out (nullable) receives strlen. */
int s2b_encode_latlng_levels_tokens(const double* latlng_rad, size_t n, const int* levels, int num_levels, uint64_t* ids_out,
                                    int token_level_index, char* tokens16_out, uint8_t* len_out);
int s2b_encode_latlng_levels_tokens_dev(const double* latlng_rad_dev, size_t n, const int* levels, int num_levels,
                                        uint64_t* ids_out_dev, int token_level_index, char* tokens16_out_dev,
                                        uint8_t* len_out_dev, void* stream);

/* S2CellId::parent(level) (s2cell_id.h:662-668). Ids whose level is < `level` are the caller's error
 * (reference: DCHECK only); they are transformed by the same bit formula. */
int s2b_cellid_parent(const uint64_t* ids, size_t n, int level, uint64_t* out);
int s2b_cellid_parent_dev(const uint64_t* ids_dev, size_t n, int level, uint64_t* out_dev, void* stream);

/* S2CellId::ToToken (s2cell_id.cc:209-233): tokens16_out is n fixed 16-byte slots, NUL padded;
 * len_out (nullable) receives strlen. id 0 -> "X". */
int s2b_cellid_to_token(const uint64_t* ids, size_t n, char* tokens16_out, uint8_t* len_out);
int s2b_cellid_to_token_dev(const uint64_t* ids_dev, size_t n, char* tokens16_out_dev, uint8_t* len_out_dev, void* stream);

/* The step after encode in most pipelines: batched inverses.
 * S2CellId::ToPoint (s2cell_id.h:184; GetCenterSiTi :555-581, FaceSiTitoXYZ s2coords.cc:68-72, Normalize): the unit
 * vector of each cell's centre, bit-exact with the reference. xyz_out = n x 3 doubles. */
int s2b_cellid_to_point(const uint64_t* ids, size_t n, double* xyz_out);
int s2b_cellid_to_point_dev(const uint64_t* ids_dev, size_t n, double* xyz_out_dev, void* stream);
/* S2CellId::ToLatLng (s2cell_id.cc:381 = S2LatLng(ToPointRaw()), s2latlng.h:235-250): (lat, lng) in radians, 2 doubles
 * per id. Floating-point output: atan2 comes from the CUDA math library (<= 2 ulp), the reference's from the host libm,
 * so results agree to a few ulp rather than bit for bit (ids and booleans elsewhere in this API are exact). */
int s2b_cellid_to_latlng(const uint64_t* ids, size_t n, double* latlng_rad_out);
int s2b_cellid_to_latlng_dev(const uint64_t* ids_dev, size_t n, double* latlng_rad_out_dev, void* stream);
/* S2CellId::GetEdgeNeighbors (s2cell_id.cc:499-512): the four edge neighbours (down, right, up, left) of every VALID
 * cell id, at the cell's own level; neighbors4_out = n x 4 ids. Neighbours across a cube-face edge are found with
 * FromFaceIJWrap (:458-489). */
int s2b_cellid_edge_neighbors(const uint64_t* ids, size_t n, uint64_t* neighbors4_out);
int s2b_cellid_edge_neighbors_dev(const uint64_t* ids_dev, size_t n, uint64_t* neighbors4_out_dev, void* stream);
/* S2CellId::is_valid / level / range_min / range_max (s2cell_id.h:583-603, 642-652) element-wise; every output is
 * optional (NULL). level_out is -1 for invalid ids. */
int s2b_cellid_info(const uint64_t* ids, size_t n, uint8_t* valid_out, int8_t* level_out, uint64_t* range_min_out, uint64_t* range_max_out);
int s2b_cellid_info_dev(const uint64_t* ids_dev, size_t n, uint8_t* valid_out_dev, int8_t* level_out_dev, uint64_t* range_min_out_dev,
                        uint64_t* range_max_out_dev, void* stream);
/* Traversal helpers of S2CellId, element-wise (s2cell_id.h:670-677, s2cell_id.cc:119-168, :193-207):
 *   children:  children4_out[4 i + k] = ids[i].child(k), k = 0..3 (Hilbert order); 0 for leaf cells and invalid ids
 *   advance:   out[i] = ids[i].advance(steps[i]) (clamped to [Begin(level), End(level)]) or, with wrap != 0,
 *              ids[i].advance_wrap(steps[i]); next() / prev() are steps of +1 / -1. 0 for invalid ids
 *   common_ancestor_level: level_out[i] = a[i].GetCommonAncestorLevel(b[i]) (-1: different faces, or an invalid id) */
int s2b_cellid_children(const uint64_t* ids, size_t n, uint64_t* children4_out);
int s2b_cellid_children_dev(const uint64_t* ids_dev, size_t n, uint64_t* children4_out_dev, void* stream);
int s2b_cellid_advance(const uint64_t* ids, const int64_t* steps, size_t n, int wrap, uint64_t* out);
int s2b_cellid_advance_dev(const uint64_t* ids_dev, const int64_t* steps_dev, size_t n, int wrap, uint64_t* out_dev, void* stream);
int s2b_cellid_common_ancestor_level(const uint64_t* a, const uint64_t* b, size_t n, int8_t* level_out);
int s2b_cellid_common_ancestor_level_dev(const uint64_t* a_dev, const uint64_t* b_dev, size_t n, int8_t* level_out_dev, void* stream);
/* S2CellId::AppendVertexNeighbors(level) (s2cell_id.cc:514-556): the cells of `level` around the vertex of
 * parent(level) closest to the cell, in the reference's order: neighbors4_out[4*i + k], count_out[i] = 3 (cube vertex)
 * or 4; unused slots are 0. Requires level < level(ids[i]) (as the reference does); otherwise count 0. */
int s2b_cellid_vertex_neighbors(const uint64_t* ids, size_t n, int level, uint64_t* neighbors4_out, uint8_t* count_out);
int s2b_cellid_vertex_neighbors_dev(const uint64_t* ids_dev, size_t n, int level, uint64_t* neighbors4_out_dev,
                                    uint8_t* count_out_dev, void* stream);
/* S2CellId::AppendAllNeighbors(nbr_level) (s2cell_id.cc:558-598): all cells of nbr_level (>= level(ids[i])) adjacent
 * to the cell, in the reference's order, at out[i*out_stride ...]; count_out[i] = 4 * 2^(nbr_level - level) + 4 (8 for
 * nbr_level == level). Entries beyond out_stride are dropped, unused slots are 0. */
int s2b_cellid_all_neighbors(const uint64_t* ids, size_t n, int nbr_level, uint64_t* out, size_t out_stride, uint32_t* count_out);
int s2b_cellid_all_neighbors_dev(const uint64_t* ids_dev, size_t n, int nbr_level, uint64_t* out_dev, size_t out_stride,
                                 uint32_t* count_out_dev, void* stream);

/* S2CellId::FromToken (s2cell_id.cc:235-254): tokens as produced by s2b_cellid_to_token (16-byte slots + lengths);
 * malformed tokens decode to 0 (S2CellId::None()). */
int s2b_cellid_from_token(const char* tokens16, const uint8_t* len, size_t n, uint64_t* ids_out);
int s2b_cellid_from_token_dev(const char* tokens16_dev, const uint8_t* len_dev, size_t n, uint64_t* ids_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Shape index: a flattened, device-resident MutableS2ShapeIndex.
 *
 * S2bFlatIndex is the index walked once through the reference's public iteration API
 * (S2ShapeIndex::Iterator BEGIN..done(), cell().clipped(i), shape(id)->edge(e), dimension();
 * s2shape_index.h:244-258,286-530) — INTEGRATION.md has the 30-line flattener. All pointers are HOST pointers; the
 * arrays are copied.
 *   cell_ids[C]            ascending, disjoint S2CellIds (index cells)
 *   cell_center[C*3]       S2CellId::ToPoint() of each cell (s2cell_id.h:184) — may be NULL: computed here
 *   cell_clip_begin[C+1]   clipped shapes of cell c are [cell_clip_begin[c], cell_clip_begin[c+1])
 *   clip_shape_id[K]       ascending within a cell (S2ShapeIndexCell order, s2shape_index.h:133-191)
 *   clip_flags[K]          bit0 = contains_center, bits1-2 = shape dimension (0,1,2)
 *   clip_edge_begin[K+1]   clipped edges of clip k are [clip_edge_begin[k], clip_edge_begin[k+1]), ascending edge id
 *   edge_v0/edge_v1[E*3]   endpoints of each clipped edge (S2Shape::edge(e), s2shape.h:185)
 * ---------------------------------------------------------------------------------------------- */
typedef struct S2bFlatIndex {
  uint64_t num_cells, num_clips, num_edges;
  int32_t num_shape_ids;
  const uint64_t* cell_ids;
  const double* cell_center;
  const uint32_t* cell_clip_begin;
  const int32_t* clip_shape_id;
  const uint8_t* clip_flags;
  const uint32_t* clip_edge_begin;
  const double* edge_v0;
  const double* edge_v1;
} S2bFlatIndex;

/* Builds the flat index on the host directly from shapes — the library's own counterpart of
 * MutableS2ShapeIndex::Add + ForceBuild (mutable_s2shape_index.h:372,423; build: mutable_s2shape_index.cc:643-1925)
 * for callers that do not have a reference-built index to flatten. Shapes are a "soup" of vertex chains
 * (S2LaxPolygonShape / S2LaxPolylineShape / S2PointVectorShape semantics):
 *   shape s has dimension shape_dim[s] (2 = closed loops, interior on the left; 1 = polylines; 0 = points) and
 *   chains [shape_chain_begin[s], shape_chain_begin[s+1]); chain c has vertices
 *   [chain_vertex_begin[c], chain_vertex_begin[c+1]) of `vertices` (x,y,z triples of unit vectors).
 * max_edges_per_cell <= 0 selects the reference default (10). The result owns host arrays; s2b_built_index_flat
 * exposes them as an S2bFlatIndex (valid until s2b_built_index_destroy) and, optionally, the shape-local edge id
 * of every clipped edge. The cells are not required to coincide with MutableS2ShapeIndex's; query results do.
 *
 * Granularity for device-resident indexes: query results do not depend on max_edges_per_cell, only the work split
 * does. On the GPU finer cells win (fewer points reach the edge loops, more are answered by Locate alone): measured on
 * B200, 10 -> 1 takes the 10k-polygon join from 43 to 66 Gpts/s and the in-cap single-loop query from 52 to 83 Gpts/s,
 * for a 3.5x larger index. S2B_MAX_EDGES_PER_CELL_DEVICE is the recommended value (the same option exists on the
 * reference: MutableS2ShapeIndex::Options::set_max_edges_per_cell). */
#define S2B_MAX_EDGES_PER_CELL_DEVICE 1
typedef struct S2bBuiltIndex S2bBuiltIndex;
int s2b_index_build(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                    const uint32_t* chain_vertex_begin, const double* vertices, int max_edges_per_cell,
                    S2bBuiltIndex** built_out);
int s2b_built_index_flat(const S2bBuiltIndex* built, S2bFlatIndex* flat_out, const int32_t** edge_id_out);
void s2b_built_index_destroy(S2bBuiltIndex* built);

/* Reads the reference's serialized index: the bytes written by MutableS2ShapeIndex::Encode
 * (mutable_s2shape_index.cc:1988-2009; cells: S2ShapeIndexCell::Encode s2shape_index.cc:69-195; cell ids:
 * EncodeS2CellIdVector encoded_s2cell_id_vector.cc:42-95), i.e. what MutableS2ShapeIndex::Init
 * (mutable_s2shape_index.cc:2011-2040) / EncodedS2ShapeIndex::Init consume. That format holds cells, clipped shapes
 * and edge ids only; the geometry is supplied as the same shape soup s2b_index_build takes (shape ids = positions in
 * the soup, edge ids numbered as S2LaxPolygonShape / S2LaxPolylineShape / S2PointVectorShape number them). The result
 * has exactly the reference index's cells. Malformed input fails with S2B_EINVAL (never crashes), like Init's `false`.
 * max_edges_per_cell_out / consumed_out (bytes read) are optional. */
int s2b_index_decode(const void* bytes, size_t len, int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                     const uint32_t* chain_vertex_begin, const double* vertices, S2bBuiltIndex** built_out,
                     int* max_edges_per_cell_out, size_t* consumed_out);
/* Writes a flat index in that format, byte-identical to MutableS2ShapeIndex::Encode for an index with the same cells;
 * edge_id[E] = shape-local id of every clipped edge (s2b_built_index_flat's edge_id_out). *size_out receives the
 * encoded size; if it exceeds `capacity` (or out is NULL) nothing is written and S2B_ECAPACITY is returned. */
int s2b_index_encode(const S2bFlatIndex* flat, const int32_t* edge_id, int max_edges_per_cell, void* out, size_t capacity,
                     size_t* size_out);

/* The geometry half of a serialized index: the bytes of s2shapeutil::FastEncodeTaggedShapes / CompactEncodeTaggedShapes
 * (s2shapeutil_coding.cc:134-156; read back there by TaggedShapeFactory / FullDecodeShape, :73-110, :158-172) for every
 * shape type the reference can decode: S2Polygon::Shape (tag 1; S2Polygon::Decode, lossless and compressed — loops become
 * oriented chains, holes reversed as S2Polygon::Shape orients its edges), S2Polyline::Shape (tag 2; lossless and
 * compressed), S2PointVectorShape (tag 3), S2LaxPolylineShape (tag 4) and S2LaxPolygonShape (tag 5), whose points are
 * EncodedS2PointVectors in either format (UNCOMPRESSED or CELL_IDS, encoded_s2point_vector.cc). Decoding yields the shape
 * soup s2b_index_build / s2b_index_decode take (arrays valid until s2b_shape_soup_destroy); vertices and edge order are
 * bit-identical to the reference's decoders. Unknown tags and malformed input fail with S2B_EINVAL.
 * s2b_shapes_encode writes the lax shapes as s2shapeutil::FastEncodeTaggedShapes does (CodingHint::FAST, uncompressed
 * point vectors); s2b_shapes_encode_hint with S2B_CODING_COMPACT writes what CompactEncodeTaggedShapes writes
 * (EncodeS2PointVectorCompact, encoded_s2point_vector.cc:323-594: the CELL_IDS block coding for vectors in which at
 * least 5 % of the points are cell centres, the uncompressed form otherwise). Both are byte-identical to the reference's
 * encoders; two-call capacity protocol as s2b_index_encode. */
#define S2B_CODING_FAST 0
#define S2B_CODING_COMPACT 1
typedef struct S2bShapeSoup S2bShapeSoup;
int s2b_shapes_decode(const void* bytes, size_t len, S2bShapeSoup** soup_out, size_t* consumed_out);
int s2b_shape_soup_arrays(const S2bShapeSoup* soup, int* num_shapes, const int8_t** shape_dim, const uint32_t** shape_chain_begin,
                          const uint32_t** chain_vertex_begin, const double** vertices, size_t* num_vertices);
void s2b_shape_soup_destroy(S2bShapeSoup* soup);
int s2b_shapes_encode(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin, const uint32_t* chain_vertex_begin,
                      const double* vertices, void* out, size_t capacity, size_t* size_out);
int s2b_shapes_encode_hint(int coding_hint, int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                           const uint32_t* chain_vertex_begin, const double* vertices, void* out, size_t capacity, size_t* size_out);

typedef struct S2bIndex S2bIndex;

/* Uploads the index to the CURRENT CUDA device (one replica per GPU: call once per process/device). */
int s2b_index_create(const S2bFlatIndex* flat, S2bIndex** index_out);
/* Wraps an S2CellUnion (S2CellUnion::cell_ids(): sorted, non-overlapping) as an index whose cells are all "interior"
 * cells of shape 0: s2b_contains over it is S2CellUnion::Contains(S2Point) (s2cell_union.cc:564-566) on the fast
 * Locate path. Preferable to s2b_cellunion_contains when the same union is queried with many batches. */
int s2b_index_create_from_cellunion(const uint64_t* sorted_ids, size_t m, S2bIndex** index_out);
void s2b_index_destroy(S2bIndex* index);
/* S2ContainsPointQuery::VisitIncidentEdges (s2contains_point_query.h:305-328): for every point, the edges of its index
 * cell that have the point as an endpoint, as (shape id, edge id) pairs in the reference's visiting order, CSR: the pairs
 * of point i are [offsets_out[i], offsets_out[i+1]). The flat index does not carry edge ids; attach them once with
 * s2b_index_set_edge_ids (edge_id[E] = shape-local id of every clipped edge, e.g. s2b_built_index_flat's edge_id_out).
 * Host form: S2B_ECAPACITY with *total_out set if `capacity` is too small. Device form: count, scan, fill. */
int s2b_index_set_edge_ids(S2bIndex* index, const int32_t* edge_id, size_t num_edges);
int s2b_incident_edges(const S2bIndex* index, const double* xyz, size_t n, uint32_t* offsets_out, int32_t* shape_ids_out,
                       int32_t* edge_ids_out, size_t capacity, size_t* total_out);
int s2b_incident_edges_count_dev(const S2bIndex* index, const double* xyz_dev, size_t n, uint32_t* counts_out_dev, void* stream);
int s2b_incident_edges_fill_dev(const S2bIndex* index, const double* xyz_dev, size_t n, const uint32_t* offsets_dev,
                                int32_t* shape_ids_out_dev, int32_t* edge_ids_out_dev, void* stream);
/* S2ShapeIndex::Iterator::Locate(const S2Point&) (S2CellIterator::LocateImpl, s2cell_iterator.h:137-157) as a batch query:
 * cell_out[i] = position (in the flat index's cell arrays) of the index cell containing point i, or -1. */
int s2b_index_locate_points(const S2bIndex* index, const double* xyz, size_t n, int32_t* cell_out);
int s2b_index_locate_points_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int32_t* cell_out_dev, void* stream);
/* S2ShapeIndex::Iterator::Locate(S2CellId) (S2CellIterator::LocateImpl, s2cell_iterator.h:159-189) for a batch of target
 * cells: relation_out[i] = S2CellRelation (0 INDEXED: the target is contained in index cell cell_out[i]; 1 SUBDIVIDED: it
 * contains index cells, the first is cell_out[i]; 2 DISJOINT, cell_out[i] = -1). cell_out (nullable) indexes the flat
 * index's cell arrays. */
int s2b_index_locate_cells(const S2bIndex* index, const uint64_t* cell_ids, size_t n, uint8_t* relation_out, int32_t* cell_out);
int s2b_index_locate_cells_dev(const S2bIndex* index, const uint64_t* cell_ids_dev, size_t n, uint8_t* relation_out_dev,
                               int32_t* cell_out_dev, void* stream);

/* S2ShapeIndexRegion's cell tests (s2shape_index_region.h) for a batch of target cells: what an S2RegionCoverer, or a
 * cell-vs-polygon classification join, asks of an index region. One target per thread: Locate(S2CellId), then for a
 * target inside an index cell AnyEdgeIntersects (:447-465: S2::ClipToPaddedFace + S2::IntersectsRect,
 * s2edge_clipping.cc:323-384, padded by kFaceClipErrorUVCoord + kIntersectsRectErrorUVDist) and ShapeContains of the
 * target's centre (SEMI_OPEN). Results depend on the index's cell structure exactly as the reference's do (a target that
 * is SUBDIVIDED "may intersect"), so they are identical to the reference's over an index with the same cells.
 *   s2b_region_may_intersect_cells  S2ShapeIndexRegion::MayIntersect(const S2Cell&)  :344-371
 *   s2b_region_contains_cells       S2ShapeIndexRegion::Contains(const S2Cell&)      :313-342
 * S2::RobustCrossProd (s2edge_crossings.cc:147-359) runs with all of its stages on the device: the double-precision
 * product, the x87 long double product the reference uses on x86-64 (emulated bit for bit), exact arithmetic and symbolic
 * perturbation — edges shorter than 9 * DBL_EPSILON, antipodal and exactly proportional endpoints included.
 * The _dev forms enqueue on `stream` but wait for it once per round of 2^27 targets (the number of targets that need edge
 * tests sizes the sort that precedes them), so they return with the results complete. */
int s2b_region_may_intersect_cells(const S2bIndex* index, const uint64_t* cell_ids, size_t n, uint8_t* out);
int s2b_region_may_intersect_cells_dev(const S2bIndex* index, const uint64_t* cell_ids_dev, size_t n, uint8_t* out_dev, void* stream);
int s2b_region_contains_cells(const S2bIndex* index, const uint64_t* cell_ids, size_t n, uint8_t* out);
int s2b_region_contains_cells_dev(const S2bIndex* index, const uint64_t* cell_ids_dev, size_t n, uint8_t* out_dev, void* stream);
/* S2ShapeIndexRegion::VisitIntersectingShapeIds (:373-424) as a CSR: the pairs of target i are
 * [offsets_out[i], offsets_out[i+1]) of (shape_ids_out, contains_out), ascending by shape id (the reference visits a
 * SUBDIVIDED target's shapes in hash-map order, i.e. unspecified). contains_out = 1 when the shape contains the whole
 * target cell. *total_out (nullable) = number of pairs; S2B_ECAPACITY after filling offsets_out if capacity is smaller. */
int s2b_region_intersecting_shapes(const S2bIndex* index, const uint64_t* cell_ids, size_t n, uint32_t* offsets_out,
                                   int32_t* shape_ids_out, uint8_t* contains_out, size_t capacity, size_t* total_out);
/* S2ShapeIndexRegion::GetCellUnionBound (:232-305): at most 6 cells covering the index (what a coverer starts from). Host
 * logic over the index's cell ids (one device-to-host copy of them per call). */
int s2b_region_cell_union_bound(const S2bIndex* index, uint64_t cell_ids_out[6], int* count_out);
/* The S2RegionCoverer loop (s2region_coverer.cc:GetCoveringInternal) for a fixed level, run on the device level by level:
 * starting from GetCellUnionBound, every frontier cell is tested with MayIntersect, survivors are subdivided until `level`.
 * The result is the covering S2RegionCoverer::GetCovering(S2ShapeIndexRegion) returns with min_level = max_level = level
 * (where max_cells does not bind): sorted level-`level` cell ids. *total_out = their number; S2B_ECAPACITY if it exceeds
 * capacity (nothing written). */
int s2b_region_covering_at_level(const S2bIndex* index, int level, uint64_t* cell_ids_out, size_t capacity, size_t* total_out);
/* Any output may be NULL. device_bytes = HBM held by the replica (ids + cells + clips + edges + lookup table). */
int s2b_index_info(const S2bIndex* index, uint64_t* num_cells, uint64_t* num_clips, uint64_t* num_edges,
                   int32_t* num_shape_ids, uint64_t* device_bytes);

/* S2ContainsPointQuery<S2ShapeIndex>(index, vertex_model).Contains(p)  (s2contains_point_query.h:255-265):
 * out[i] = 1 if any shape contains point i. Locate = s2cell_iterator.h:137-155.
 * Scratch: the point queries of this section take their work lists from the device's stream-ordered memory pool
 * (cudaMallocAsync on the call's stream, kept cached in the pool afterwards): at most 24 bytes per point of a round of up
 * to 2^28 points (6.4 GB), less when that much cannot be allocated. */
int s2b_contains(const S2bIndex* index, const double* xyz, size_t n, int vertex_model, uint8_t* out);
int s2b_contains_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int vertex_model, uint8_t* out_dev, void* stream);

/* Hit count of a batch: *count_dev += number of non-zero bytes of flags_dev[0..n) (the output of s2b_contains_dev /
 * s2b_cellunion_contains_dev). ADDS into the device counter (zero it first; all-reduce it across GPUs): the "how many of
 * these points does the region contain" half of BASELINE configs[4], at HBM speed (1 byte per point). */
int s2b_count_flags_dev(const uint8_t* flags_dev, size_t n, uint64_t* count_dev, void* stream);

/* ...ShapeContains(shape_id, p)  (s2contains_point_query.h:267-280). */
int s2b_shape_contains(const S2bIndex* index, int shape_id, const double* xyz, size_t n, int vertex_model, uint8_t* out);
int s2b_shape_contains_dev(const S2bIndex* index, int shape_id, const double* xyz_dev, size_t n, int vertex_model,
                           uint8_t* out_dev, void* stream);

/* Spatial join: counts[s] = number of points contained by shape s, s in [0, num_shape_ids)
 * (VisitContainingShapeIds, s2contains_point_query.h:282-295, accumulated per shape).
 * Host form overwrites counts; the _dev form ADDS into counts_dev (zero it first) so batches and, across GPUs,
 * an NCCL all-reduce of the uint64 array compose. */
int s2b_shape_histogram(const S2bIndex* index, const double* xyz, size_t n, int vertex_model, uint64_t* counts);
int s2b_shape_histogram_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int vertex_model,
                            uint64_t* counts_dev, void* stream);

/* GetContainingShapeIds (s2contains_point_query.h:342-350) for a batch, CSR form: shape ids of point i are
 * shape_ids_out[offsets_out[i] .. offsets_out[i+1]) in ascending order. total_out receives the number of
 * (point, shape) pairs; if it exceeds capacity the call returns S2B_ECAPACITY after filling offsets_out only. */
int s2b_containing_shapes(const S2bIndex* index, const double* xyz, size_t n, int vertex_model, uint32_t* offsets_out,
                          int32_t* shape_ids_out, size_t capacity, size_t* total_out);
/* Device building blocks of the same: per-point counts, then (after the caller's exclusive scan) the fill. */
int s2b_containing_shapes_count_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int vertex_model,
                                    uint32_t* counts_out_dev, void* stream);
int s2b_containing_shapes_fill_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int vertex_model,
                                   const uint32_t* offsets_dev, int32_t* shape_ids_out_dev, void* stream);

/* --------------------------------------------------------------------------------------------------
 * Multi-GPU driver (one process, every GPU of the box). The reference has no multi-device path; points
 * are independent, so batches are sharded by contiguous range (shard k of R gets [k*n/R, (k+1)*n/R), sizes differing by
 * at most one), the index is replicated, and the only exchange is the all-reduce of the uint64[num_shape_ids] hit
 * histogram over NVLink.
 *   s2b_index_replicate     copies the index GPU-to-GPU onto every listed device that does not hold it yet (the index's own
 *                           device is replica 0 whether or not it is listed), creates one stream per replica, an NCCL
 *                           communicator over the replica devices (ncclCommInitAll; NCCL is bound at run time) and enables
 *                           peer access between them. Replicas are owned by `index` and freed by s2b_index_destroy.
 *   s2b_index_set_reduce_mode  S2B_REDUCE_NCCL: ncclAllReduce(uint64, sum) on every replica's stream (the default when NCCL
 *                           loaded); S2B_REDUCE_PEER: the library's own reduction — one small kernel per GPU that sums the
 *                           peers' partial histograms with direct loads over NVLink peer memory, ordered behind the
 *                           peers' query kernels with events. Returns the previous mode, or S2B_EUNSUPPORTED.
 *   s2b_shape_histogram_multi / s2b_contains_multi   the host-pointer calls of the same name without _multi, sharded over
 *                           the replicas (one pipeline and worker thread per GPU); without replicas they are the
 *                           single-GPU calls.
 *   s2b_shape_histogram_multi_dev   device-resident shards: shard k = n[k] points at xyz_dev[k] on replica k's device
 *                           (s2b_index_replica_devices gives the order). Enqueues the query on each replica's own stream,
 *                           then the all-reduce; counts_out_dev[k] (array and entries nullable) receives the GLOBAL
 *                           histogram on device k. synchronize != 0 waits for all replicas' streams.
 *   s2b_encode_*_multi      encoding sharded over `devices` (no index, no exchange); lat/lng follows s2b_set_latlng_mode. */
#define S2B_REDUCE_NCCL 0
#define S2B_REDUCE_PEER 1
int s2b_index_replicate(S2bIndex* index, const int* devices, int ndev);
int s2b_index_num_replicas(const S2bIndex* index);
int s2b_index_replica_devices(const S2bIndex* index, int* devices_out, int capacity);
int s2b_index_set_reduce_mode(S2bIndex* index, int mode);
int s2b_shape_histogram_multi(const S2bIndex* index, const double* xyz, size_t n, int vertex_model, uint64_t* counts);
int s2b_shape_histogram_multi_dev(const S2bIndex* index, const double* const* xyz_dev, const size_t* n, int vertex_model,
                                  uint64_t* const* counts_out_dev, int synchronize);
int s2b_contains_multi(const S2bIndex* index, const double* xyz, size_t n, int vertex_model, uint8_t* out);
int s2b_encode_points_multi(const int* devices, int ndev, const double* xyz, size_t n, int level, uint64_t* ids_out);
int s2b_encode_latlng_levels_tokens_multi(const int* devices, int ndev, const double* latlng_rad, size_t n, const int* levels,
                                          int num_levels, uint64_t* ids_out, int token_level_index, char* tokens16_out,
                                          uint8_t* len_out);

/* Diagnostic: the index queries locate points with a single-precision pre-filter and fall back to the double-
 * precision Locate for every point the filter cannot decide with certainty (DESIGN.md K3 phase A0). on != 0 sends all
 * points through the double-precision path, which must give identical results (tests check it). Returns the old value. */
int s2b_diag_force_exact_locate(int on);

/* Number of predicate evaluations that needed the exact (arbitrary-precision) stage on the device since the
 * last reset, summed over all launches of this process on the current device. Resets the counter when reset != 0.
 * There is no CPU stage behind it: this is purely a statistic. */
int s2b_exact_stage_count(uint64_t* count_out, int reset);

/* --------------------------------------------------------------------------------------------------
 * S2PointIndex + S2ClosestPointQuery, batch form (s2point_index.h, s2closest_point_query.h; the radius join of
 * doc/examples/point_index.cc:25-52). The index holds m points with data = their position in `xyz`
 * (S2PointIndex<int>::Add(xyz[i], i)), sorted by leaf cell id on the current device.
 * Distances are S1ChordAngle values passed as their squared chord length: max_length2 =
 * S1ChordAngle(S1Angle::Radians(r)).length2() = (2 sin(r/2))^2 (s1chord_angle.h:339-350); INFINITY = no limit.
 * As in the reference, max_distance is exclusive (distance < limit).
 *   s2b_closest_points_count: counts_out[i] = FindClosestPoints(PointTarget(targets[i])).size() with
 *                             options.max_distance = the limit and unlimited max_results
 *   s2b_closest_point:        FindClosestPoint(PointTarget(targets[i])): data_out[i] = data of the nearest index point
 *                             within the limit or -1, length2_out[i] (nullable) = its S1ChordAngle length2 (INFINITY if
 *                             none). Among equidistant points the smallest data value is returned.
 * Results are exact: the distance of every candidate is S1ChordAngle(point, target) evaluated as the reference does.
 * -------------------------------------------------------------------------------------------------- */
typedef struct S2bPointIndex S2bPointIndex;
int s2b_point_index_create(const double* xyz, size_t m, S2bPointIndex** index_out);
void s2b_point_index_destroy(S2bPointIndex* index);
size_t s2b_point_index_size(const S2bPointIndex* index);
int s2b_closest_points_count(const S2bPointIndex* index, const double* targets_xyz, size_t n, double max_length2, uint32_t* counts_out);
int s2b_closest_points_count_dev(const S2bPointIndex* index, const double* targets_xyz_dev, size_t n, double max_length2,
                                 uint32_t* counts_out_dev, void* stream);
int s2b_closest_point(const S2bPointIndex* index, const double* targets_xyz, size_t n, double max_length2, int32_t* data_out,
                      double* length2_out);
/* FindClosestPoints with options.max_results = k (1..32) and max_distance: row i of data_out / length2_out (n x k, nullable
 * length2_out) holds the count_out[i] (nullable) nearest index points of target i within the limit, nearest first, padded
 * with -1 / INFINITY; among equidistant points the smaller data value comes first. */
int s2b_closest_points(const S2bPointIndex* index, const double* targets_xyz, size_t n, int max_results, double max_length2,
                       int32_t* data_out, double* length2_out, uint32_t* count_out);
int s2b_closest_points_dev(const S2bPointIndex* index, const double* targets_xyz_dev, size_t n, int max_results, double max_length2,
                           int32_t* data_out_dev, double* length2_out_dev, uint32_t* count_out_dev, void* stream);
int s2b_closest_point_dev(const S2bPointIndex* index, const double* targets_xyz_dev, size_t n, double max_length2,
                          int32_t* data_out_dev, double* length2_out_dev, void* stream);
/* The same query for the reference's other target types (s2closest_point_query.h:110-131) and its full option set
 * (s2closest_point_query_base.h:82-101):
 *   target_kind  S2B_TARGET_POINT  PointTarget: targets = n x 3 doubles
 *                S2B_TARGET_EDGE   EdgeTarget(a, b): targets = n x 6 doubles; distance = S2::UpdateMinDistance(p, a, b)
 *                                  (s2edge_distances.cc:93-223, with S2::RobustCrossProd in all of its stages)
 *                S2B_TARGET_CELL   CellTarget(S2Cell(id)): targets = n uint64 cell ids; distance = S2Cell::GetDistance(p)
 *                                  (s2cell.cc:382-436; 0 inside the cell)
 *   max_results  1..32: the nearest k index points within the limit per target (rows of data_out / length2_out as
 *                s2b_closest_points); 0: options.max_results unlimited — only count_out[i] = the number of index points within
 *                the limit is reported (data_out / length2_out may be NULL)
 *   max_length2  S1ChordAngle length2 of options.max_distance (exclusive); INFINITY = none
 *   max_error_radians  options.max_error: the reference MAY then substitute points up to that much further away; the results
 *                here are always the exact ones, which satisfies every max_error >= 0
 * (S2ClosestPointQuery::ShapeIndexTarget and options.region are not provided.) */
#define S2B_TARGET_POINT 0
#define S2B_TARGET_EDGE 1
#define S2B_TARGET_CELL 2
int s2b_closest_points_to_targets(const S2bPointIndex* index, int target_kind, const void* targets, size_t n, int max_results,
                                  double max_length2, double max_error_radians, int32_t* data_out, double* length2_out,
                                  uint32_t* count_out);
int s2b_closest_points_to_targets_dev(const S2bPointIndex* index, int target_kind, const void* targets_dev, size_t n, int max_results,
                                      double max_length2, double max_error_radians, int32_t* data_out_dev, double* length2_out_dev,
                                      uint32_t* count_out_dev, void* stream);
/* FindClosestPoints(ShapeIndexTarget(index)) (s2closest_point_query.h:132-165; S2MinDistanceShapeIndexTarget,
 * s2min_distance_targets.cc): ONE target — a whole shape index resident on the same device as the point index. The distance
 * of an index point p is what S2ClosestEdgeQuery(index).FindClosestEdge(PointTarget(p)) returns: 0 if include_interiors
 * != 0 and a polygon of the index contains p (SEMI_OPEN), else the minimum of S2::UpdateMinDistance(p, v0, v1) over the
 * index's edges, bit for bit. Results: the index points with distance < max_length2 (INFINITY = no limit), nearest first
 * (ties: smaller data first), at most max_results of them (0 = no limit). *total_out = the number of results; the first
 * min(total, capacity) are written to data_out / length2_out (nullable) and S2B_ECAPACITY is returned if total > capacity.
 * max_error_radians is accepted as in the calls above (the results are exact). */
int s2b_closest_points_to_shape_index(const S2bPointIndex* index, const S2bIndex* target, int include_interiors, int max_results,
                                      double max_length2, double max_error_radians, int32_t* data_out, double* length2_out,
                                      size_t capacity, size_t* total_out);
/* The same computation seen from the other side: S2ClosestEdgeQuery(index).GetDistance(PointTarget(points[i])) for a batch
 * (s2closest_edge_query.h; options include_interiors and max_distance). length2_out[i] = S1ChordAngle::length2() of the
 * distance from points[i] to the index's geometry — 0 inside a polygon when include_interiors != 0, the minimum of
 * S2::UpdateMinDistance over the index's edges otherwise, bit for bit — or INFINITY if that is not below max_length2
 * (exclusive). The work grows with the number of (edge, point) pairs within the limit: pass a finite limit for large
 * batches (INFINITY compares every edge with every point). Must be called with the index's device current; the _dev form
 * takes device pointers and returns after the default stream has finished. */
int s2b_index_point_distances(const S2bIndex* index, const double* xyz, size_t n, int include_interiors, double max_length2,
                              double* length2_out);
int s2b_index_point_distances_dev(const S2bIndex* index, const double* xyz_dev, size_t n, int include_interiors, double max_length2,
                                  double* length2_out_dev);

/* --------------------------------------------------------------------------------------------------
 * S2CellIndex joins, batch form (s2cell_index.h:113-441, s2cell_index.cc:71-163) and the S2RegionSharder built on it
 * (s2region_sharder.cc:35-140). The index holds m (cell_id, label) pairs — S2CellIndex::Add(cell_ids[i], labels[i]) followed
 * by Build() — on the current device; cells may overlap and repeat, labels are >= 0 (invalid ids / negative labels:
 * S2B_EINVAL, the reference DCHECKs them).
 * Targets are S2CellUnions in CSR form: target t = target_ids[target_offsets[t] .. target_offsets[t+1]), cells sorted and
 * non-overlapping as S2CellUnion requires; target_offsets == NULL means T single-cell targets. Invalid target ids report
 * nothing. Results are CSR too (offsets_out has T + 1 slots); S2B_ECAPACITY with *total_out set (offsets filled) if
 * `capacity` is too small.
 *   s2b_cell_index_intersecting_cells   VisitIntersectingCells(target): every indexed (cell_id, label) pair whose cell
 *                                       intersects the target, once per copy in the index, in unspecified order as in the
 *                                       reference; cell_ids_out / labels_out are each nullable
 *   s2b_cell_index_intersecting_labels  GetIntersectingLabels(target): the distinct labels, ascending
 *   s2b_cell_index_most_intersecting_label  S2RegionSharder(shards).GetMostIntersectingShard(region, default_label)
 *                                       (s2region_sharder.cc:35-140) over an index with label s on the cells of shard s (each
 *                                       shard a normalized S2CellUnion). Per region the caller passes `bound` = the normalized
 *                                       region.GetCellUnionBound() covering (CSR) and, optionally, the region itself as an
 *                                       S2CellUnion (CSR, region_ids NULL = the bound is the region): a shard's score is the
 *                                       sum of lsb() over the cells of (shard ∩ bound) that intersect the region
 *                                       (region.MayIntersect for an S2CellUnion region); a single candidate shard wins
 *                                       outright, otherwise the largest score, the smaller label on ties, default_label if none
 *   s2b_cell_index_point_histogram      for every point, the pairs VisitIntersectingCells reports for its leaf cell
 *                                       S2CellId(point): counts[label] += 1 per pair (num_labels = max label + 1 slots).
 *                                       Host form overwrites counts, the _dev form ADDS (zero it first; all-reduce across GPUs)
 * The _dev forms take device pointers, run on `stream`, and (except the histogram) synchronise it to size their scratch.
 * -------------------------------------------------------------------------------------------------- */
typedef struct S2bCellIndex S2bCellIndex;
int s2b_cell_index_create(const uint64_t* cell_ids, const int32_t* labels, size_t m, S2bCellIndex** index_out);
void s2b_cell_index_destroy(S2bCellIndex* index);
int s2b_cell_index_info(const S2bCellIndex* index, uint64_t* num_cells, int32_t* num_labels, uint64_t* device_bytes);
int s2b_cell_index_intersecting_cells(const S2bCellIndex* index, const uint64_t* target_ids, const uint32_t* target_offsets, size_t T,
                                      uint32_t* offsets_out, uint64_t* cell_ids_out, int32_t* labels_out, size_t capacity, size_t* total_out);
int s2b_cell_index_intersecting_cells_dev(const S2bCellIndex* index, const uint64_t* target_ids_dev, const uint32_t* target_offsets_dev, size_t T,
                                          uint32_t* offsets_out_dev, uint64_t* cell_ids_out_dev, int32_t* labels_out_dev, size_t capacity,
                                          size_t* total_out, void* stream);
int s2b_cell_index_intersecting_labels(const S2bCellIndex* index, const uint64_t* target_ids, const uint32_t* target_offsets, size_t T,
                                       uint32_t* offsets_out, int32_t* labels_out, size_t capacity, size_t* total_out);
int s2b_cell_index_intersecting_labels_dev(const S2bCellIndex* index, const uint64_t* target_ids_dev, const uint32_t* target_offsets_dev, size_t T,
                                           uint32_t* offsets_out_dev, int32_t* labels_out_dev, size_t capacity, size_t* total_out, void* stream);
int s2b_cell_index_most_intersecting_label(const S2bCellIndex* index, const uint64_t* bound_ids, const uint32_t* bound_offsets,
                                           const uint64_t* region_ids, const uint32_t* region_offsets, size_t T, int32_t default_label,
                                           int32_t* out);
int s2b_cell_index_most_intersecting_label_dev(const S2bCellIndex* index, const uint64_t* bound_ids_dev, const uint32_t* bound_offsets_dev,
                                               const uint64_t* region_ids_dev, const uint32_t* region_offsets_dev, size_t T,
                                               int32_t default_label, int32_t* out_dev, void* stream);
int s2b_cell_index_point_histogram(const S2bCellIndex* index, const double* xyz, size_t n, uint64_t* counts);
int s2b_cell_index_point_histogram_dev(const S2bCellIndex* index, const double* xyz_dev, size_t n, uint64_t* counts_dev, void* stream);

/* S2CellUnion::Normalize (s2cell_union.cc:171-199) on the GPU: sorts, drops duplicates and cells contained in other
 * cells, and replaces complete sibling quads by their parent until none is left. out has room for m ids;
 * *count_out (host) receives the number of cells. The _dev form synchronises the stream to report the count. */
int s2b_cellunion_normalize(const uint64_t* ids, size_t m, uint64_t* out, size_t* count_out);
int s2b_cellunion_normalize_dev(const uint64_t* ids_dev, size_t m, uint64_t* out_dev, size_t* count_out, void* stream);
/* S2CellUnion::Contains(S2CellId) and ::Intersects(S2CellId) (s2cell_union.cc:285-309) for a batch of cell ids against
 * one union (sorted, non-overlapping): out[i] bit0 = contains, bit1 = intersects. */
int s2b_cellunion_contains_cells(const uint64_t* sorted_ids, size_t m, const uint64_t* ids, size_t n, uint8_t* out);
int s2b_cellunion_contains_cells_dev(const uint64_t* sorted_ids_dev, size_t m, const uint64_t* ids_dev, size_t n,
                                     uint8_t* out_dev, void* stream);

/* S2CellUnion::Contains(const S2Point&) (s2cell_union.cc:285-300,564-566): sorted_ids are the union's cell ids
 * (S2CellUnion::cell_ids(): sorted, non-overlapping), out[i] = 1 if the union contains point i. */
int s2b_cellunion_contains(const uint64_t* sorted_ids, size_t m, const double* xyz, size_t n, uint8_t* out);
int s2b_cellunion_contains_dev(const uint64_t* sorted_ids_dev, size_t m, const double* xyz_dev, size_t n, uint8_t* out_dev,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2B_CAPI_H_ */
