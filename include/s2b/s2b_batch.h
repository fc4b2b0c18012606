// s2b_batch.h — header-only C++ sugar over the C ABI (s2b_capi.h) that keeps the reference's names for the hot path.
//
// The reference (google/s2geometry) exposes this path as per-point C++ calls:
//   S2CellId(const S2Point&), S2CellId(const S2LatLng&)        src/s2/s2cell_id.h:106-128
//   S2CellId::parent(level), S2CellId::ToToken()               src/s2/s2cell_id.h:662-668, s2cell_id.cc:217-233
//   S2ContainsPointQuery<Index>::Contains / ShapeContains / VisitContainingShapeIds / GetContainingShapeIds
//                                                              src/s2/s2contains_point_query.h:92-186
//   S2CellUnion::Contains(const S2Point&)                      src/s2/s2cell_union.cc:564-566
//   S2CellIndex::Add/Build/VisitIntersectingCells/GetIntersectingLabels   src/s2/s2cell_index.h:143-195
// The batch forms below take spans of the same value types (layout compatible: S2Point = 3 doubles, S2CellId = one
// uint64, S2LatLng = 2 doubles in radians) and run on the GPU. Errors: the reference's path has no error returns
// (DCHECK only); here every call returns false on failure and s2b::LastError() explains (no exceptions).
// INTEGRATION.md shows how these map onto static members added to the reference's own classes.
#ifndef S2B_BATCH_H_
#define S2B_BATCH_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <cmath>
#include <vector>

#include "../s2b_capi.h"

namespace s2b {

struct Point { double x, y, z; };          // == S2Point (Vector3<double>, s2point.h:35)
struct LatLng { double lat_rad, lng_rad; }; // == S2LatLng's R2Point of radians (s2latlng.h:47-60)
static_assert(sizeof(Point) == 24 && sizeof(LatLng) == 16, "layout must match the reference value types");

enum class S2VertexModel { OPEN = S2B_VERTEX_MODEL_OPEN, SEMI_OPEN = S2B_VERTEX_MODEL_SEMI_OPEN, CLOSED = S2B_VERTEX_MODEL_CLOSED };

inline const char* LastError() { return s2b_last_error(); }

// ---- S2CellId batch constructors and accessors -------------------------------------------------------------
struct S2CellIdBatch {
  // ids[i] = S2CellId(points[i]).parent(level).id()   (level 30 = leaf)
  static bool FromPoints(const Point* points, size_t n, uint64_t* ids, int level = 30) {
    return s2b_encode_points(&points->x, n, level, ids) == S2B_OK;
  }
  static bool FromPoints(const std::vector<Point>& points, std::vector<uint64_t>* ids, int level = 30) {
    ids->resize(points.size());
    return FromPoints(points.data(), points.size(), ids->data(), level);
  }
  // ids[i] = S2CellId(S2LatLng::FromRadians(lat, lng)).parent(level).id()
  static bool FromLatLngs(const LatLng* ll, size_t n, uint64_t* ids, int level = 30, uint8_t* boundary_flags = nullptr) {
    return s2b_encode_latlng(&ll->lat_rad, n, level, ids, boundary_flags) == S2B_OK;
  }
  // S2LatLng::FromE7 pairs
  static bool FromE7(const int32_t* lat_lng_e7, size_t n, uint64_t* ids, int level = 30) {
    return s2b_encode_latlng_e7(lat_lng_e7, n, level, ids, nullptr) == S2B_OK;
  }
  static bool Parent(const uint64_t* ids, size_t n, int level, uint64_t* out) { return s2b_cellid_parent(ids, n, level, out) == S2B_OK; }
  // tokens: n fixed 16-byte slots, NUL padded (S2CellId::ToToken)
  static bool ToTokens(const uint64_t* ids, size_t n, char* tokens16, uint8_t* lengths = nullptr) {
    return s2b_cellid_to_token(ids, n, tokens16, lengths) == S2B_OK;
  }
  static std::vector<std::string> ToTokens(const std::vector<uint64_t>& ids) {
    std::vector<char> buf(16 * ids.size());
    std::vector<uint8_t> len(ids.size());
    std::vector<std::string> out;
    if (!ToTokens(ids.data(), ids.size(), buf.data(), len.data())) return out;
    out.reserve(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) out.emplace_back(buf.data() + 16 * i, len[i]);
    return out;
  }
};

// ---- A shape index resident on the current CUDA device ------------------------------------------------------
// Owns the device replica. Construct from an S2bFlatIndex (a reference-built MutableS2ShapeIndex walked through its
// iterator — see INTEGRATION.md — or the result of s2b_index_build).
class DeviceShapeIndex {
 public:
  DeviceShapeIndex() = default;
  explicit DeviceShapeIndex(const S2bFlatIndex& flat) { Init(flat); }
  DeviceShapeIndex(const DeviceShapeIndex&) = delete;
  DeviceShapeIndex& operator=(const DeviceShapeIndex&) = delete;
  DeviceShapeIndex(DeviceShapeIndex&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ~DeviceShapeIndex() { s2b_index_destroy(h_); }
  bool Init(const S2bFlatIndex& flat) {
    s2b_index_destroy(h_);
    h_ = nullptr;
    return s2b_index_create(&flat, &h_) == S2B_OK;
  }
  // Builds the index from shapes with the library's own host builder (no reference library involved).
  bool InitFromShapes(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                      const uint32_t* chain_vertex_begin, const double* vertices, int max_edges_per_cell = 0) {
    S2bBuiltIndex* built = nullptr;
    if (s2b_index_build(num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices, max_edges_per_cell, &built) != S2B_OK) return false;
    S2bFlatIndex flat;
    bool ok = s2b_built_index_flat(built, &flat, nullptr) == S2B_OK && Init(flat);
    s2b_built_index_destroy(built);
    return ok;
  }
  // The counterpart of MutableS2ShapeIndex::Init(Decoder*, ShapeFactory) (mutable_s2shape_index.cc:2011-2040): `encoded`
  // is the output of MutableS2ShapeIndex::Encode, the shape arrays stand in for the ShapeFactory.
  bool InitFromEncoded(const void* encoded, size_t len, int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                       const uint32_t* chain_vertex_begin, const double* vertices) {
    S2bBuiltIndex* built = nullptr;
    if (s2b_index_decode(encoded, len, num_shapes, shape_dim, shape_chain_begin, chain_vertex_begin, vertices, &built, nullptr, nullptr) != S2B_OK) return false;
    S2bFlatIndex flat;
    bool ok = s2b_built_index_flat(built, &flat, nullptr) == S2B_OK && Init(flat);
    s2b_built_index_destroy(built);
    return ok;
  }
  // Multi-GPU (one process): copies the index onto every listed CUDA device; the batch queries below then shard their
  // points over all replicas and all-reduce the per-shape counts (NCCL, or the library's own peer-memory reduction).
  bool Replicate(const std::vector<int>& devices) { return h_ && s2b_index_replicate(h_, devices.data(), (int)devices.size()) == S2B_OK; }
  int num_replicas() const { return s2b_index_num_replicas(h_); }
  bool ok() const { return h_ != nullptr; }
  int num_shape_ids() const { int32_t s = 0; if (h_) s2b_index_info(h_, nullptr, nullptr, nullptr, &s, nullptr); return s; }
  const S2bIndex* handle() const { return h_; }

 private:
  S2bIndex* h_ = nullptr;
};

// ---- S2ContainsPointQuery, batch form -------------------------------------------------------------------------
class S2ContainsPointQueryBatch {
 public:
  explicit S2ContainsPointQueryBatch(const DeviceShapeIndex* index, S2VertexModel model = S2VertexModel::SEMI_OPEN)
      : index_(index), model_((int)model) {}
  // out[i] = Contains(points[i])
  bool Contains(const Point* points, size_t n, uint8_t* out) const { return s2b_contains_multi(index_->handle(), &points->x, n, model_, out) == S2B_OK; }
  // out[i] = ShapeContains(shape_id, points[i])
  bool ShapeContains(int shape_id, const Point* points, size_t n, uint8_t* out) const {
    return s2b_shape_contains(index_->handle(), shape_id, &points->x, n, model_, out) == S2B_OK;
  }
  // counts[s] = |{ i : shape s contains points[i] }|  — VisitContainingShapeIds accumulated per shape (spatial join)
  bool CountContainingShapes(const Point* points, size_t n, std::vector<uint64_t>* counts) const {
    counts->assign((size_t)index_->num_shape_ids(), 0);
    return s2b_shape_histogram_multi(index_->handle(), &points->x, n, model_, counts->data()) == S2B_OK;
  }
  // GetContainingShapeIds for every point, CSR: ids of point i are shape_ids[offsets[i] .. offsets[i+1])
  bool GetContainingShapeIds(const Point* points, size_t n, std::vector<uint32_t>* offsets, std::vector<int32_t>* shape_ids) const {
    offsets->resize(n + 1);
    shape_ids->resize(n / 8 + 1024);
    size_t total = 0;
    int rc = s2b_containing_shapes(index_->handle(), &points->x, n, model_, offsets->data(), shape_ids->data(), shape_ids->size(), &total);
    if (rc == S2B_ECAPACITY && total > shape_ids->size()) {
      shape_ids->resize(total);
      rc = s2b_containing_shapes(index_->handle(), &points->x, n, model_, offsets->data(), shape_ids->data(), shape_ids->size(), &total);
    }
    if (rc != S2B_OK) return false;
    shape_ids->resize(total);
    return true;
  }

 private:
  const DeviceShapeIndex* index_;
  int model_;
};

// ---- S2ShapeIndexRegion, batch form (s2shape_index_region.h) --------------------------------------------------------
// The S2Region interface of an indexed set of shapes, asked about many cells at once (what an S2RegionCoverer does one
// cell at a time): MayIntersect(S2Cell), Contains(S2Cell), VisitIntersectingShapeIds. Cells are passed as S2CellIds.
class S2ShapeIndexRegionBatch {
 public:
  explicit S2ShapeIndexRegionBatch(const DeviceShapeIndex* index) : index_(index) {}
  // out[i] = MayIntersect(S2Cell(cell_ids[i]))
  bool MayIntersect(const uint64_t* cell_ids, size_t n, uint8_t* out) const {
    return s2b_region_may_intersect_cells(index_->handle(), cell_ids, n, out) == S2B_OK;
  }
  // out[i] = Contains(S2Cell(cell_ids[i]))
  bool Contains(const uint64_t* cell_ids, size_t n, uint8_t* out) const {
    return s2b_region_contains_cells(index_->handle(), cell_ids, n, out) == S2B_OK;
  }
  // out[i] = Contains(points[i])  (SEMI_OPEN, dimension-2 shapes; s2shape_index_region.h:437-445)
  bool Contains(const Point* points, size_t n, uint8_t* out) const {
    return s2b_contains(index_->handle(), &points->x, n, S2B_VERTEX_MODEL_SEMI_OPEN, out) == S2B_OK;
  }
  // VisitIntersectingShapeIds for every cell, CSR: pairs of cell i are [offsets[i], offsets[i+1]) of (shape_ids, contains),
  // ascending by shape id; contains[k] = 1 when the shape contains the whole cell.
  bool GetIntersectingShapeIds(const uint64_t* cell_ids, size_t n, std::vector<uint32_t>* offsets, std::vector<int32_t>* shape_ids,
                               std::vector<uint8_t>* contains) const {
    offsets->resize(n + 1);
    size_t total = 0;
    int rc = s2b_region_intersecting_shapes(index_->handle(), cell_ids, n, offsets->data(), nullptr, nullptr, 0, &total);
    if (rc == S2B_ECAPACITY) {
      shape_ids->resize(total); contains->resize(total);
      rc = s2b_region_intersecting_shapes(index_->handle(), cell_ids, n, offsets->data(), shape_ids->data(), contains->data(), total, &total);
    } else {
      shape_ids->clear(); contains->clear();
    }
    return rc == S2B_OK;
  }

 private:
  const DeviceShapeIndex* index_;
};

// ---- S2CellUnion::Contains(S2Point), batch form ------------------------------------------------------------------
// cell_ids = S2CellUnion::cell_ids() (sorted, non-overlapping).
inline bool S2CellUnionContainsBatch(const uint64_t* cell_ids, size_t num_cells, const Point* points, size_t n, uint8_t* out) {
  return s2b_cellunion_contains(cell_ids, num_cells, &points->x, n, out) == S2B_OK;
}

// ---- S2CellIndex, batch form ---------------------------------------------------------------------------------------
// Add() collects (cell_id, label) pairs on the host, Build() uploads the index to the current CUDA device; the queries take a
// batch of targets. Targets are S2CellUnions in CSR form (cell ids sorted and non-overlapping per target).
class S2CellIndexBatch {
 public:
  S2CellIndexBatch() = default;
  S2CellIndexBatch(const S2CellIndexBatch&) = delete;
  S2CellIndexBatch& operator=(const S2CellIndexBatch&) = delete;
  ~S2CellIndexBatch() { s2b_cell_index_destroy(h_); }
  void Add(uint64_t cell_id, int32_t label) { ids_.push_back(cell_id); labels_.push_back(label); }
  void Add(const std::vector<uint64_t>& cell_union, int32_t label) { for (uint64_t id : cell_union) Add(id, label); }
  int num_cells() const { return (int)ids_.size(); }
  bool Build() {
    s2b_cell_index_destroy(h_);
    h_ = nullptr;
    return s2b_cell_index_create(ids_.data(), labels_.data(), ids_.size(), &h_) == S2B_OK;
  }
  void Clear() { s2b_cell_index_destroy(h_); h_ = nullptr; ids_.clear(); labels_.clear(); }
  // VisitIntersectingCells for every target: pairs of target t are (cell_ids, labels)[offsets[t] .. offsets[t+1]).
  bool GetIntersectingCells(const std::vector<uint64_t>& target_ids, const std::vector<uint32_t>& target_offsets,
                            std::vector<uint32_t>* offsets, std::vector<uint64_t>* cell_ids, std::vector<int32_t>* labels) const {
    const size_t T = target_offsets.empty() ? 0 : target_offsets.size() - 1;
    offsets->assign(T + 1, 0);
    size_t total = 0;
    int rc = s2b_cell_index_intersecting_cells(h_, target_ids.data(), target_offsets.data(), T, offsets->data(), nullptr, nullptr, 0, &total);
    if (rc != S2B_OK && rc != S2B_ECAPACITY) return false;
    cell_ids->resize(total);
    labels->resize(total);
    if (total == 0) return true;
    return s2b_cell_index_intersecting_cells(h_, target_ids.data(), target_offsets.data(), T, offsets->data(), cell_ids->data(), labels->data(), total,
                                             &total) == S2B_OK;
  }
  // GetIntersectingLabels for every target: distinct labels, ascending.
  bool GetIntersectingLabels(const std::vector<uint64_t>& target_ids, const std::vector<uint32_t>& target_offsets,
                             std::vector<uint32_t>* offsets, std::vector<int32_t>* labels) const {
    const size_t T = target_offsets.empty() ? 0 : target_offsets.size() - 1;
    offsets->assign(T + 1, 0);
    labels->resize(2 * T + 1024);
    size_t total = 0;
    int rc = s2b_cell_index_intersecting_labels(h_, target_ids.data(), target_offsets.data(), T, offsets->data(), labels->data(), labels->size(), &total);
    if (rc == S2B_ECAPACITY && total > labels->size()) {
      labels->resize(total);
      rc = s2b_cell_index_intersecting_labels(h_, target_ids.data(), target_offsets.data(), T, offsets->data(), labels->data(), labels->size(), &total);
    }
    if (rc != S2B_OK) return false;
    labels->resize(total);
    return true;
  }
  // counts[label] = number of (point, indexed cell) pairs whose cell contains the point (VisitIntersectingCells on the leaf cell).
  bool CountPointsPerLabel(const Point* points, size_t n, std::vector<uint64_t>* counts) const {
    int32_t num_labels = 0;
    if (s2b_cell_index_info(h_, nullptr, &num_labels, nullptr) != S2B_OK) return false;
    counts->assign((size_t)num_labels, 0);
    return s2b_cell_index_point_histogram(h_, &points->x, n, counts->data()) == S2B_OK;
  }
  const S2bCellIndex* handle() const { return h_; }

 private:
  std::vector<uint64_t> ids_;
  std::vector<int32_t> labels_;
  S2bCellIndex* h_ = nullptr;
};

// ---- S2PointIndex<int> + S2ClosestPointQuery<int>, batch form (s2point_index.h, s2closest_point_query.h) -----------------
// Add() collects points on the host (data = the order of insertion, as S2PointIndex<int>::Add(point, i)); Build() sorts them
// by leaf cell id on the current CUDA device.
class S2PointIndexBatch {
 public:
  S2PointIndexBatch() = default;
  S2PointIndexBatch(const S2PointIndexBatch&) = delete;
  S2PointIndexBatch& operator=(const S2PointIndexBatch&) = delete;
  ~S2PointIndexBatch() { s2b_point_index_destroy(h_); }
  void Add(const Point& p) { points_.push_back(p); }
  void Add(const Point* points, size_t n) { points_.insert(points_.end(), points, points + n); }
  int num_points() const { return (int)points_.size(); }
  bool Build() {
    s2b_point_index_destroy(h_);
    h_ = nullptr;
    return s2b_point_index_create(points_.empty() ? nullptr : &points_[0].x, points_.size(), &h_) == S2B_OK;
  }
  const S2bPointIndex* handle() const { return h_; }

 private:
  std::vector<Point> points_;
  S2bPointIndex* h_ = nullptr;
};

// Options mirror S2ClosestPointQuery::Options: max_results (0 = unlimited; per-target lists hold at most 32), max_distance
// as S1ChordAngle::length2() (exclusive; INFINITY = none), max_error in radians. Results per target: rows of max_results
// (data, length2) entries, nearest first, padded with -1 / INFINITY, and the number found.
class S2ClosestPointQueryBatch {
 public:
  explicit S2ClosestPointQueryBatch(const S2PointIndexBatch* index) : index_(index) {}
  void set_max_results(int k) { max_results_ = k; }
  void set_max_distance_length2(double length2) { max_length2_ = length2; }
  void set_max_error_radians(double e) { max_error_ = e; }
  // FindClosestPoints(PointTarget(targets[i]))
  bool FindClosestPoints(const Point* targets, size_t n, std::vector<int32_t>* data, std::vector<double>* length2, std::vector<uint32_t>* counts) const {
    return Find(S2B_TARGET_POINT, &targets->x, n, data, length2, counts);
  }
  // FindClosestPoints(EdgeTarget(a, b)): targets = n x (a, b)
  bool FindClosestPointsToEdges(const Point* endpoints_ab, size_t n, std::vector<int32_t>* data, std::vector<double>* length2, std::vector<uint32_t>* counts) const {
    return Find(S2B_TARGET_EDGE, &endpoints_ab->x, n, data, length2, counts);
  }
  // FindClosestPoints(CellTarget(S2Cell(cell_ids[i])))
  bool FindClosestPointsToCells(const uint64_t* cell_ids, size_t n, std::vector<int32_t>* data, std::vector<double>* length2, std::vector<uint32_t>* counts) const {
    return Find(S2B_TARGET_CELL, cell_ids, n, data, length2, counts);
  }
  // FindClosestPoints(ShapeIndexTarget(&index)) with include_interiors: ONE target, a list of (data, length2), nearest first.
  bool FindClosestPointsToShapeIndex(const DeviceShapeIndex& target, bool include_interiors, std::vector<int32_t>* data, std::vector<double>* length2) const {
    size_t total = 0, cap = max_results_ > 0 ? (size_t)max_results_ : 4096;
    for (;;) {
      data->resize(cap); length2->resize(cap);
      const int rc = s2b_closest_points_to_shape_index(index_->handle(), target.handle(), include_interiors ? 1 : 0, max_results_, max_length2_, max_error_,
                                                       data->data(), length2->data(), cap, &total);
      if (rc == S2B_ECAPACITY) { cap = total; continue; }
      if (rc != S2B_OK) return false;
      data->resize(total); length2->resize(total);
      return true;
    }
  }

 private:
  bool Find(int kind, const void* targets, size_t n, std::vector<int32_t>* data, std::vector<double>* length2, std::vector<uint32_t>* counts) const {
    const size_t k = (size_t)(max_results_ > 0 ? max_results_ : 0);
    data->assign(n * (k ? k : 1), -1);
    length2->assign(n * (k ? k : 1), HUGE_VAL);
    counts->assign(n, 0);
    return s2b_closest_points_to_targets(index_->handle(), kind, targets, n, (int)k, max_length2_, max_error_, data->data(), length2->data(), counts->data()) == S2B_OK;
  }
  const S2PointIndexBatch* index_;
  int max_results_ = 1;
  double max_length2_ = HUGE_VAL, max_error_ = 0.0;
};

}  // namespace s2b

#endif  // S2B_BATCH_H_
