// oracle/ref_capi.cc — TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
//
// A C ABI over the UNMODIFIED reference library (google/s2geometry sources compiled in place from
// /root/reference/src against oracle/absl_shim). It is the "reference" oracle the GPU path is checked
// against and the multithreaded CPU baseline bench.py times (cpu_baseline.kind = "reference").
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Every function calls the reference's own public API for the path:
//   S2CellId(S2Point)/S2CellId(S2LatLng)      s2cell_id.cc:309-317
//   S2ContainsPointQuery                      s2contains_point_query.h:255-389
//   MutableS2ShapeIndex                       mutable_s2shape_index.h
//   s2pred::Sign / S2::CrossingSign / ...     s2predicates.h, s2edge_crossings.h
//   S2CellUnion::Contains, S2RegionCoverer    s2cell_union.cc:564, s2region_coverer.h

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "s2/mutable_s2shape_index.h"
#include "s2/s1angle.h"
#include "s2/s2cap.h"
#include "s2/s2cell_id.h"
#include "s2/s2cell_index.h"
#include "s2/s2cell_union.h"
#include "s2/s2region_sharder.h"
#include "s2/s2closest_edge_query.h"
#include "s2/s2closest_point_query.h"
#include "s2/s2point_index.h"
#include "s2/s2contains_point_query.h"
#include "s2/s2edge_crosser.h"
#include "s2/s2edge_crossings.h"
#include "s2/s2latlng.h"
#include "s2/s2loop.h"
#include "s2/s2polygon.h"
#include "s2/s2polyline.h"
#include "s2/s2lax_polygon_shape.h"
#include "s2/s2lax_polyline_shape.h"
#include "s2/s2point_vector_shape.h"
#include "s2/encoded_string_vector.h"
#include "s2/s2point.h"
#include "s2/s2pointutil.h"
#include "s2/s2predicates.h"
#include "s2/s2predicates_internal.h"
#include "s2/s2region_coverer.h"
#include "s2/s2shape.h"
#include "s2/s2cell.h"
#include "s2/s2edge_clipping.h"
#include "s2/s2shape_index_region.h"
#include "s2/s2shapeutil_contains_brute_force.h"
#include "s2/s2shapeutil_get_reference_point.h"
#include "s2/util/coding/coder.h"

namespace {

// Geometry enters through the reference's own plugin interface (S2Shape, s2shape.h:64-260): a shape is
// a list of vertex chains; dimension 2 chains are closed loops (S2LaxPolygonShape semantics), dimension 1
// chains are polylines (S2LaxPolylineShape), dimension 0 vertices are degenerate edges (S2PointVectorShape).
class SoupShape : public S2Shape {
 public:
  SoupShape(int dim, std::vector<std::vector<S2Point>> chains) : dim_(dim), chains_(std::move(chains)) {
    int start = 0;
    for (const auto& c : chains_) {
      int n = (int)c.size();
      int ne = dim_ == 2 ? n : (dim_ == 1 ? std::max(0, n - 1) : n);
      starts_.push_back(start);
      start += ne;
    }
    starts_.push_back(start);
  }
  int num_edges() const override { return starts_.back(); }
  Edge edge(int e) const override {
    int c = (int)(std::upper_bound(starts_.begin(), starts_.end(), e) - starts_.begin()) - 1;
    return chain_edge(c, e - starts_[c]);
  }
  int dimension() const override { return dim_; }
  ReferencePoint GetReferencePoint() const override {
    if (dim_ < 2) return ReferencePoint::Contained(false);
    return s2shapeutil::GetReferencePoint(*this);
  }
  int num_chains() const override { return (int)chains_.size(); }
  Chain chain(int i) const override { return Chain(starts_[i], starts_[i + 1] - starts_[i]); }
  Edge chain_edge(int i, int j) const override {
    const auto& c = chains_[i];
    if (dim_ == 0) return Edge(c[j], c[j]);
    int k = j + 1;
    if (dim_ == 2 && k == (int)c.size()) k = 0;
    return Edge(c[j], c[k]);
  }
  ChainPosition chain_position(int e) const override {
    int c = (int)(std::upper_bound(starts_.begin(), starts_.end(), e) - starts_.begin()) - 1;
    return ChainPosition(c, e - starts_[c]);
  }

 private:
  int dim_;
  std::vector<std::vector<S2Point>> chains_;
  std::vector<int> starts_;
};

struct RefIndex {
  MutableS2ShapeIndex index;
  explicit RefIndex(const MutableS2ShapeIndex::Options& o) : index(o) {}
};

inline S2Point P(const double* p) { return S2Point(p[0], p[1], p[2]); }

template <class F>
void ParallelFor(size_t n, int threads, F f) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads == 1 || n < 1024) { f(0, n, 0); return; }
  std::vector<std::thread> pool;
  size_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    size_t b = std::min(n, (size_t)t * chunk), e = std::min(n, b + chunk);
    if (b < e) pool.emplace_back([=] { f(b, e, t); });
  }
  for (auto& th : pool) th.join();
}

S2VertexModel Model(int m) { return m == 0 ? S2VertexModel::OPEN : (m == 1 ? S2VertexModel::SEMI_OPEN : S2VertexModel::CLOSED); }

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---------------------------------------------------------------- cell ids
void ref_encode_points(const double* xyz, size_t n, int level, uint64_t* out, int threads) {
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) {
      S2CellId id(P(xyz + 3 * i));
      out[i] = (level >= 30 ? id : id.parent(level)).id();
    }
  });
}

void ref_encode_latlng(const double* latlng_rad, size_t n, int level, uint64_t* out, int threads) {
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) {
      S2CellId id(S2LatLng::FromRadians(latlng_rad[2 * i], latlng_rad[2 * i + 1]));
      out[i] = (level >= 30 ? id : id.parent(level)).id();
    }
  });
}

void ref_encode_latlng_deg(const double* latlng_deg, size_t n, int level, uint64_t* out, int threads) {
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) {
      S2CellId id(S2LatLng::FromDegrees(latlng_deg[2 * i], latlng_deg[2 * i + 1]));
      out[i] = (level >= 30 ? id : id.parent(level)).id();
    }
  });
}

void ref_encode_latlng_e7(const int32_t* latlng_e7, size_t n, int level, uint64_t* out, int threads) {
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) {
      S2CellId id(S2LatLng::FromE7(latlng_e7[2 * i], latlng_e7[2 * i + 1]));
      out[i] = (level >= 30 ? id : id.parent(level)).id();
    }
  });
}

// BASELINE config 2 as the reference would run it: one S2CellId(S2LatLng) per point, parent() at each requested
// level, ToToken() of the id at levels[token_level_index]. ids_out: num_levels arrays of n; tokens: 16-byte slots.
void ref_encode_latlng_levels_tokens(const double* latlng_rad, size_t n, const int* levels, int num_levels, uint64_t* ids_out,
                                     int token_level_index, char* tokens16, uint8_t* len_out, int threads) {
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) {
      S2CellId leaf(S2LatLng::FromRadians(latlng_rad[2 * i], latlng_rad[2 * i + 1]));
      for (int k = 0; k < num_levels; ++k) {
        S2CellId id = levels[k] >= 30 ? leaf : leaf.parent(levels[k]);
        ids_out[(size_t)k * n + i] = id.id();
        if (k == token_level_index) {
          std::string t = id.ToToken();
          std::memset(tokens16 + 16 * i, 0, 16);
          std::memcpy(tokens16 + 16 * i, t.data(), std::min<size_t>(16, t.size()));
          if (len_out) len_out[i] = (uint8_t)t.size();
        }
      }
    }
  });
}

// S2LatLng::ToPoint only (the libm-dependent front end), for diagnosing lat/lng parity.
void ref_latlng_to_point(const double* latlng_rad, size_t n, double* xyz) {
  for (size_t i = 0; i < n; ++i) {
    S2Point p = S2LatLng::FromRadians(latlng_rad[2 * i], latlng_rad[2 * i + 1]).ToPoint();
    xyz[3 * i] = p[0]; xyz[3 * i + 1] = p[1]; xyz[3 * i + 2] = p[2];
  }
}

void ref_cellid_parent(const uint64_t* ids, size_t n, int level, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = S2CellId(ids[i]).parent(level).id();
}

void ref_cellid_level(const uint64_t* ids, size_t n, int32_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = S2CellId(ids[i]).level();
}

void ref_cellid_range(const uint64_t* ids, size_t n, uint64_t* rmin, uint64_t* rmax) {
  for (size_t i = 0; i < n; ++i) { rmin[i] = S2CellId(ids[i]).range_min().id(); rmax[i] = S2CellId(ids[i]).range_max().id(); }
}

// tokens: 16 bytes per id, NUL padded; len_out = strlen.
void ref_cellid_to_token(const uint64_t* ids, size_t n, char* tokens16, uint8_t* len_out, int threads) {
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) {
      std::string t = S2CellId(ids[i]).ToToken();
      std::memset(tokens16 + 16 * i, 0, 16);
      std::memcpy(tokens16 + 16 * i, t.data(), std::min<size_t>(16, t.size()));
      if (len_out) len_out[i] = (uint8_t)t.size();
    }
  });
}

void ref_cellid_from_token(const char* tokens16, const uint8_t* len, size_t n, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = S2CellId::FromToken(absl::string_view(tokens16 + 16 * i, len[i])).id();
}

void ref_edge_neighbors(const uint64_t* ids, size_t n, uint64_t* out4) {
  for (size_t i = 0; i < n; ++i) {
    S2CellId nb[4];
    S2CellId(ids[i]).GetEdgeNeighbors(nb);
    for (int k = 0; k < 4; ++k) out4[4 * i + k] = nb[k].id();
  }
}

void ref_cellid_to_latlng(const uint64_t* ids, size_t n, double* latlng) {
  for (size_t i = 0; i < n; ++i) {
    S2LatLng ll = S2CellId(ids[i]).ToLatLng();
    latlng[2 * i] = ll.lat().radians(); latlng[2 * i + 1] = ll.lng().radians();
  }
}

// AppendVertexNeighbors(level): out4 (unused = 0) + count.
void ref_vertex_neighbors(const uint64_t* ids, size_t n, int level, uint64_t* out4, uint8_t* count) {
  for (size_t i = 0; i < n; ++i) {
    std::vector<S2CellId> v;
    S2CellId(ids[i]).AppendVertexNeighbors(level, &v);
    for (int k = 0; k < 4; ++k) out4[4 * i + k] = k < (int)v.size() ? v[k].id() : 0;
    count[i] = (uint8_t)v.size();
  }
}

// AppendAllNeighbors(nbr_level): first `stride` neighbours (unused = 0) + full count.
void ref_all_neighbors(const uint64_t* ids, size_t n, int nbr_level, uint64_t* out, size_t stride, uint32_t* count) {
  for (size_t i = 0; i < n; ++i) {
    std::vector<S2CellId> v;
    S2CellId(ids[i]).AppendAllNeighbors(nbr_level, &v);
    for (size_t k = 0; k < stride; ++k) out[stride * i + k] = k < v.size() ? v[k].id() : 0;
    count[i] = (uint32_t)v.size();
  }
}

void ref_cellid_to_point(const uint64_t* ids, size_t n, double* xyz) {
  for (size_t i = 0; i < n; ++i) {
    S2Point p = S2CellId(ids[i]).ToPoint();
    xyz[3 * i] = p[0]; xyz[3 * i + 1] = p[1]; xyz[3 * i + 2] = p[2];
  }
}

uint64_t ref_cellid_from_face_pos_level(int face, uint64_t pos, int level) { return S2CellId::FromFacePosLevel(face, pos, level).id(); }

// ---------------------------------------------------------------- predicates (unit-test hooks)
// kind: 0 s2pred::Sign, 1 TriageSign, 2 StableSign, 3 ExactSign(perturb=true), 4 ExpensiveSign, 5 ExactSign(perturb=false)
void ref_sign(const double* abc, size_t n, int kind, int8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    S2Point a = P(abc + 9 * i), b = P(abc + 9 * i + 3), c = P(abc + 9 * i + 6);
    int s;
    switch (kind) {
      case 1: s = s2pred::TriageSign(a, b, c, a.CrossProd(b)); break;
      case 2: s = s2pred::StableSign(a, b, c); break;
      case 3: s = s2pred::ExactSign(a, b, c, true); break;
      case 4: s = s2pred::ExpensiveSign(a, b, c); break;
      case 5: s = s2pred::ExactSign(a, b, c, false); break;
      default: s = s2pred::Sign(a, b, c);
    }
    out[i] = (int8_t)s;
  }
}

// kind: 0 S2::CrossingSign (fresh S2EdgeCrosser), 1 S2CopyingEdgeCrosser, 2 VertexCrossing, 3 EdgeOrVertexCrossing
void ref_crossing(const double* abcd, size_t n, int kind, int8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    S2Point a = P(abcd + 12 * i), b = P(abcd + 12 * i + 3), c = P(abcd + 12 * i + 6), d = P(abcd + 12 * i + 9);
    int s;
    switch (kind) {
      case 1: { S2CopyingEdgeCrosser cr(a, b); s = cr.CrossingSign(c, d); break; }
      case 2: s = S2::VertexCrossing(a, b, c, d); break;
      case 3: s = S2::EdgeOrVertexCrossing(a, b, c, d); break;
      default: s = S2::CrossingSign(a, b, c, d);
    }
    out[i] = (int8_t)s;
  }
}

void ref_normalize(const double* a, size_t n, double* out) {
  for (size_t i = 0; i < n; ++i) { S2Point o = P(a + 3 * i).Normalize(); out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2]; }
}

void ref_ortho(const double* a, size_t n, double* out) {
  for (size_t i = 0; i < n; ++i) { S2Point o = S2::Ortho(P(a + 3 * i)); out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2]; }
}

// ---------------------------------------------------------------- index
// Shape soup: shape s has dimension shape_dim[s] and chains [shape_chain_begin[s], shape_chain_begin[s+1]);
// chain c has vertices [chain_vertex_begin[c], chain_vertex_begin[c+1]) of `vertices` (N×3 doubles).
void* ref_index_create(int num_shapes, const int8_t* shape_dim, const uint32_t* shape_chain_begin,
                       const uint32_t* chain_vertex_begin, const double* vertices, int max_edges_per_cell) {
  MutableS2ShapeIndex::Options opt;
  if (max_edges_per_cell > 0) opt.set_max_edges_per_cell(max_edges_per_cell);
  auto* r = new RefIndex(opt);
  for (int s = 0; s < num_shapes; ++s) {
    std::vector<std::vector<S2Point>> chains;
    for (uint32_t c = shape_chain_begin[s]; c < shape_chain_begin[s + 1]; ++c) {
      std::vector<S2Point> v;
      for (uint32_t k = chain_vertex_begin[c]; k < chain_vertex_begin[c + 1]; ++k) v.push_back(P(vertices + 3 * (size_t)k));
      chains.push_back(std::move(v));
    }
    r->index.Add(std::make_unique<SoupShape>(shape_dim[s], std::move(chains)));
  }
  r->index.ForceBuild();
  return r;
}

void ref_index_destroy(void* h) { delete (RefIndex*)h; }

// Sizes of the flattened form: C cells, K clipped shapes, E clipped edges.
void ref_index_flat_sizes(void* h, uint64_t* C, uint64_t* K, uint64_t* E) {
  auto& index = ((RefIndex*)h)->index;
  uint64_t c = 0, k = 0, e = 0;
  for (MutableS2ShapeIndex::Iterator it(&index, S2ShapeIndex::BEGIN); !it.done(); it.Next()) {
    ++c;
    const S2ShapeIndexCell& cell = it.cell();
    k += cell.num_clipped();
    for (int i = 0; i < cell.num_clipped(); ++i) e += cell.clipped(i).num_edges();
  }
  *C = c; *K = k; *E = e;
}

// Walks the reference's public iteration API (s2shape_index.h:244-258) and writes the flat arrays.
void ref_index_flatten(void* h, uint64_t* cell_ids, double* cell_center, uint32_t* cell_clip_begin,
                       int32_t* clip_shape_id, uint8_t* clip_flags, uint32_t* clip_edge_begin,
                       int32_t* edge_id, double* edge_v0, double* edge_v1) {
  auto& index = ((RefIndex*)h)->index;
  uint64_t c = 0, k = 0, e = 0;
  for (MutableS2ShapeIndex::Iterator it(&index, S2ShapeIndex::BEGIN); !it.done(); it.Next()) {
    cell_ids[c] = it.id().id();
    S2Point ctr = it.id().ToPoint();
    cell_center[3 * c] = ctr[0]; cell_center[3 * c + 1] = ctr[1]; cell_center[3 * c + 2] = ctr[2];
    cell_clip_begin[c] = (uint32_t)k;
    const S2ShapeIndexCell& cell = it.cell();
    for (int i = 0; i < cell.num_clipped(); ++i) {
      const S2ClippedShape& cl = cell.clipped(i);
      const S2Shape* shape = index.shape(cl.shape_id());
      clip_shape_id[k] = cl.shape_id();
      clip_flags[k] = (uint8_t)((cl.contains_center() ? 1 : 0) | (shape->dimension() << 1));
      clip_edge_begin[k] = (uint32_t)e;
      for (int j = 0; j < cl.num_edges(); ++j) {
        S2Shape::Edge ed = shape->edge(cl.edge(j));
        if (edge_id) edge_id[e] = cl.edge(j);
        for (int q = 0; q < 3; ++q) { edge_v0[3 * e + q] = ed.v0[q]; edge_v1[3 * e + q] = ed.v1[q]; }
        ++e;
      }
      ++k;
    }
    ++c;
  }
  cell_clip_begin[c] = (uint32_t)k;
  clip_edge_begin[k] = (uint32_t)e;
}

// MutableS2ShapeIndex::Encode (mutable_s2shape_index.cc:1988-2009). Returns the encoded size; writes if it fits.
size_t ref_index_encode(void* h, uint8_t* out, size_t capacity) {
  Encoder enc;
  ((RefIndex*)h)->index.Encode(&enc);
  if (out && enc.length() <= capacity) memcpy(out, enc.base(), enc.length());
  return enc.length();
}

// Round trip through the reference's own decoder: MutableS2ShapeIndex::Init over `bytes` with the shapes of `h` is
// rebuilt and its cells compared with h's. Returns 1 if Init succeeds and every cell/clip/edge id matches, else 0.
int ref_index_init_matches(void* h, const uint8_t* bytes, size_t len) {
  auto& index = ((RefIndex*)h)->index;
  struct Borrowed : public S2Shape {
    const S2Shape* s;
    explicit Borrowed(const S2Shape* s) : s(s) {}
    int num_edges() const override { return s->num_edges(); }
    Edge edge(int e) const override { return s->edge(e); }
    int dimension() const override { return s->dimension(); }
    ReferencePoint GetReferencePoint() const override { return s->GetReferencePoint(); }
    int num_chains() const override { return s->num_chains(); }
    Chain chain(int i) const override { return s->chain(i); }
    Edge chain_edge(int i, int j) const override { return s->chain_edge(i, j); }
    ChainPosition chain_position(int e) const override { return s->chain_position(e); }
  };
  struct Factory : public S2ShapeIndex::ShapeFactory {
    const MutableS2ShapeIndex* src;
    explicit Factory(const MutableS2ShapeIndex* src) : src(src) {}
    int size() const override { return src->num_shape_ids(); }
    std::unique_ptr<S2Shape> operator[](int i) const override { return std::make_unique<Borrowed>(src->shape(i)); }
    std::unique_ptr<ShapeFactory> Clone() const override { return std::make_unique<Factory>(src); }
  } factory(&index);
  MutableS2ShapeIndex copy;
  Decoder dec(bytes, len);
  if (!copy.Init(&dec, factory)) return 0;
  MutableS2ShapeIndex::Iterator a(&index, S2ShapeIndex::BEGIN), b(&copy, S2ShapeIndex::BEGIN);
  for (; !a.done() && !b.done(); a.Next(), b.Next()) {
    if (a.id() != b.id() || a.cell().num_clipped() != b.cell().num_clipped()) return 0;
    for (int i = 0; i < a.cell().num_clipped(); ++i) {
      const S2ClippedShape &x = a.cell().clipped(i), &y = b.cell().clipped(i);
      if (x.shape_id() != y.shape_id() || x.contains_center() != y.contains_center() || x.num_edges() != y.num_edges()) return 0;
      for (int j = 0; j < x.num_edges(); ++j) if (x.edge(j) != y.edge(j)) return 0;
    }
  }
  return a.done() && b.done();
}

// Iterator::Locate(S2Point): id of the containing index cell, or 0.
void ref_index_locate_points(void* h, const double* xyz, size_t n, uint64_t* at, int threads) {
  auto* index = &((RefIndex*)h)->index;
  ParallelFor(n, threads, [&](size_t b, size_t e, int) {
    MutableS2ShapeIndex::Iterator it(index);
    for (size_t i = b; i < e; ++i) at[i] = it.Locate(P(xyz + 3 * i)) ? it.id().id() : 0;
  });
}

// Iterator::Locate(S2CellId): relation (0 INDEXED, 1 SUBDIVIDED, 2 DISJOINT) and the id of the cell the iterator is
// positioned at afterwards (0 for DISJOINT, where the position is unspecified).
void ref_index_locate_cells(void* h, const uint64_t* ids, size_t n, uint8_t* relation, uint64_t* at) {
  auto& index = ((RefIndex*)h)->index;
  MutableS2ShapeIndex::Iterator it(&index);
  for (size_t i = 0; i < n; ++i) {
    S2CellRelation r = it.Locate(S2CellId(ids[i]));
    relation[i] = (uint8_t)static_cast<int>(r);
    at[i] = r == S2CellRelation::DISJOINT ? 0 : it.id().id();
  }
}

// S2ShapeIndexRegion::MayIntersect(S2Cell) (mode 0) / Contains(S2Cell) (mode 1) per target cell.
void ref_region_cells(void* h, const uint64_t* ids, size_t n, int mode, uint8_t* out) {
  auto& index = ((RefIndex*)h)->index;
  S2ShapeIndexRegion<MutableS2ShapeIndex> region(&index);
  for (size_t i = 0; i < n; ++i) {
    S2Cell cell{S2CellId(ids[i])};
    out[i] = mode == 0 ? region.MayIntersect(cell) : region.Contains(cell);
  }
}

// S2ShapeIndexRegion::VisitIntersectingShapeIds per target cell, pairs sorted by shape id (the visiting order of a
// SUBDIVIDED target is the hash map's). Returns the number of pairs; writes them when capacity suffices.
size_t ref_region_intersecting_shapes(void* h, const uint64_t* ids, size_t n, uint32_t* offsets, int32_t* shape_ids,
                                      uint8_t* contains, size_t capacity) {
  auto& index = ((RefIndex*)h)->index;
  S2ShapeIndexRegion<MutableS2ShapeIndex> region(&index);
  size_t total = 0;
  std::vector<std::pair<int, bool>> found;
  for (size_t i = 0; i < n; ++i) {
    offsets[i] = (uint32_t)total;
    found.clear();
    region.VisitIntersectingShapeIds(S2Cell(S2CellId(ids[i])), [&](int id, bool c) { found.emplace_back(id, c); return true; });
    std::sort(found.begin(), found.end());
    for (auto& f : found) {
      if (total < capacity) { shape_ids[total] = f.first; contains[total] = f.second; }
      ++total;
    }
  }
  offsets[n] = (uint32_t)total;
  return total;
}

// S2ShapeIndexRegion::GetCellUnionBound: returns the number of cells (at most 6) written to out.
int ref_region_cell_union_bound(void* h, uint64_t* out) {
  auto& index = ((RefIndex*)h)->index;
  S2ShapeIndexRegion<MutableS2ShapeIndex> region(&index);
  std::vector<S2CellId> cells;
  region.GetCellUnionBound(&cells);
  for (size_t i = 0; i < cells.size(); ++i) out[i] = cells[i].id();
  return (int)cells.size();
}

// S2::ClipToPaddedFace + S2::IntersectsRect on raw inputs: out_uv = (a.u, a.v, b.u, b.v) per edge, ok = return value.
void ref_clip_to_padded_face(const double* a, const double* b, const int32_t* face, size_t n, double padding, double* out_uv, uint8_t* ok) {
  for (size_t i = 0; i < n; ++i) {
    R2Point p0(0, 0), p1(0, 0);
    ok[i] = S2::ClipToPaddedFace(P(a + 3 * i), P(b + 3 * i), face[i], padding, &p0, &p1);
    out_uv[4 * i] = p0[0]; out_uv[4 * i + 1] = p0[1]; out_uv[4 * i + 2] = p1[0]; out_uv[4 * i + 3] = p1[1];
  }
}
void ref_intersects_rect(const double* uv4, const double* rect4, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    const double* r = rect4 + 4 * i;
    out[i] = S2::IntersectsRect(R2Point(uv4[4 * i], uv4[4 * i + 1]), R2Point(uv4[4 * i + 2], uv4[4 * i + 3]),
                                R2Rect(R1Interval(r[0], r[1]), R1Interval(r[2], r[3])));
  }
}

int ref_index_num_shape_ids(void* h) { return ((RefIndex*)h)->index.num_shape_ids(); }

// S2ContainsPointQuery::Contains, one query object per thread (s2contains_point_query.h:86-91).
void ref_contains(void* h, const double* xyz, size_t n, int vertex_model, uint8_t* out, int threads) {
  auto* index = &((RefIndex*)h)->index;
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    S2ContainsPointQuery<MutableS2ShapeIndex> q(index, Model(vertex_model));
    for (size_t i = b; i < e; ++i) out[i] = q.Contains(P(xyz + 3 * i));
  });
}

void ref_shape_contains(void* h, int shape_id, const double* xyz, size_t n, int vertex_model, uint8_t* out, int threads) {
  auto* index = &((RefIndex*)h)->index;
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    S2ContainsPointQuery<MutableS2ShapeIndex> q(index, Model(vertex_model));
    for (size_t i = b; i < e; ++i) out[i] = q.ShapeContains(shape_id, P(xyz + 3 * i));
  });
}

// Spatial join: counts[s] += 1 for every point contained by shape s (VisitContainingShapeIds).
void ref_shape_histogram(void* h, const double* xyz, size_t n, int vertex_model, uint64_t* counts, int threads) {
  auto* index = &((RefIndex*)h)->index;
  int S = index->num_shape_ids();
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  std::vector<std::vector<uint64_t>> local(threads, std::vector<uint64_t>(S, 0));
  ParallelFor(n, threads, [&](size_t b, size_t e, int t) {
    S2ContainsPointQuery<MutableS2ShapeIndex> q(index, Model(vertex_model));
    uint64_t* cnt = local[t].data();
    for (size_t i = b; i < e; ++i) q.VisitContainingShapeIds(P(xyz + 3 * i), [cnt](int id) { ++cnt[id]; return true; });
  });
  for (int s = 0; s < S; ++s) { uint64_t v = 0; for (auto& l : local) v += l[s]; counts[s] = v; }
}

// VisitIncidentEdges: offsets[n+1] and (shape id, edge id) pairs per point in visiting order. Returns the total count.
uint64_t ref_incident_edges(void* h, const double* xyz, size_t n, uint32_t* offsets, int32_t* shape_ids, int32_t* edge_ids, uint64_t capacity) {
  auto* index = &((RefIndex*)h)->index;
  S2ContainsPointQuery<MutableS2ShapeIndex> q(index);
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    offsets[i] = (uint32_t)total;
    q.VisitIncidentEdges(P(xyz + 3 * i), [&](const s2shapeutil::ShapeEdge& e) {
      if (total < capacity) { shape_ids[total] = e.id().shape_id; edge_ids[total] = e.id().edge_id; }
      ++total;
      return true;
    });
  }
  offsets[n] = (uint32_t)total;
  return total;
}

// GetContainingShapeIds: offsets[n+1] and ascending shape ids per point (capacity-limited). Returns total count.
uint64_t ref_containing_shapes(void* h, const double* xyz, size_t n, int vertex_model, uint32_t* offsets, int32_t* ids, uint64_t capacity) {
  auto* index = &((RefIndex*)h)->index;
  S2ContainsPointQuery<MutableS2ShapeIndex> q(index, Model(vertex_model));
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    offsets[i] = (uint32_t)total;
    for (int id : q.GetContainingShapeIds(P(xyz + 3 * i))) { if (total < capacity) ids[total] = id; ++total; }
  }
  offsets[n] = (uint32_t)total;
  return total;
}

// Ground truth of the reference's own tests: s2shapeutil::ContainsBruteForce (SEMI_OPEN).
void ref_brute_force_contains(void* h, int shape_id, const double* xyz, size_t n, uint8_t* out, int threads) {
  auto* index = &((RefIndex*)h)->index;
  const S2Shape* shape = index->shape(shape_id);
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) out[i] = s2shapeutil::ContainsBruteForce(*shape, P(xyz + 3 * i));
  });
}

// ---------------------------------------------------------------- cell unions
void ref_cellunion_contains(const uint64_t* ids, size_t m, const double* xyz, size_t n, uint8_t* out, int threads) {
  std::vector<S2CellId> v(m);
  for (size_t i = 0; i < m; ++i) v[i] = S2CellId(ids[i]);
  S2CellUnion u = S2CellUnion::FromVerbatim(std::move(v));
  const S2CellUnion* up = &u;
  ParallelFor(n, threads, [=](size_t b, size_t e, int) {
    for (size_t i = b; i < e; ++i) out[i] = up->Contains(P(xyz + 3 * i));
  });
}

// S2CellUnion normalisation (sort + merge) of arbitrary ids; returns count written.
// ---------------------------------------------------------------- tagged shapes (geometry half of a serialized index)
// The soup as the reference's own shape classes: dim 0 -> S2PointVectorShape, dim 1 (one chain) -> S2LaxPolylineShape,
// dim 2 -> S2LaxPolygonShape; encoded the way s2shapeutil::EncodeTaggedShapes does (s2shapeutil_coding.cc:134-148:
// StringVectorEncoder of [varint32 type_tag, shape.Encode(hint)]) with hint 0 = FAST, 1 = COMPACT.
static std::unique_ptr<S2Shape> SoupToLaxShape(int s, const int8_t* dim, const uint32_t* scb, const uint32_t* cvb, const double* v) {
  std::vector<std::vector<S2Point>> chains;
  for (uint32_t c = scb[s]; c < scb[s + 1]; ++c) {
    std::vector<S2Point> pts;
    for (uint32_t k = cvb[c]; k < cvb[c + 1]; ++k) pts.push_back(P(v + 3 * (size_t)k));
    chains.push_back(std::move(pts));
  }
  if (dim[s] == 0) {
    std::vector<S2Point> all;
    for (auto& c : chains) all.insert(all.end(), c.begin(), c.end());
    return std::make_unique<S2PointVectorShape>(std::move(all));
  }
  if (dim[s] == 1) return std::make_unique<S2LaxPolylineShape>(chains.empty() ? std::vector<S2Point>() : chains[0]);
  return std::make_unique<S2LaxPolygonShape>(chains);
}
// hint bit 0: CodingHint::COMPACT instead of FAST; bit 1: polygons and polylines as S2Polygon::Shape (tag 1, loops
// through S2Polygon::InitOriented) and S2Polyline::Shape (tag 2) instead of the lax shapes.
static std::unique_ptr<S2Shape> SoupToClassicShape(int s, const int8_t* dim, const uint32_t* scb, const uint32_t* cvb, const double* v) {
  if (dim[s] == 1) {
    std::vector<S2Point> pts;
    for (uint32_t c = scb[s]; c < scb[s + 1]; ++c)
      for (uint32_t k = cvb[c]; k < cvb[c + 1]; ++k) pts.push_back(P(v + 3 * k));
    auto shape = std::make_unique<S2Polyline::OwningShape>();
    shape->Init(std::make_unique<S2Polyline>(pts, S2Debug::DISABLE));
    return shape;
  }
  if (dim[s] == 2) {
    std::vector<std::unique_ptr<S2Loop>> loops;
    for (uint32_t c = scb[s]; c < scb[s + 1]; ++c) {
      std::vector<S2Point> pts;
      for (uint32_t k = cvb[c]; k < cvb[c + 1]; ++k) pts.push_back(P(v + 3 * k));
      if (pts.empty()) pts.push_back(S2Loop::kFull()[0]);   // a zero-vertex chain is the full loop (S2LaxPolygonShape convention)
      loops.push_back(std::make_unique<S2Loop>(pts, S2Debug::DISABLE));
    }
    auto poly = std::make_unique<S2Polygon>();
    poly->set_s2debug_override(S2Debug::DISABLE);
    poly->InitOriented(std::move(loops));
    auto shape = std::make_unique<S2Polygon::OwningShape>();
    shape->Init(std::move(poly));
    return shape;
  }
  return SoupToLaxShape(s, dim, scb, cvb, v);
}
size_t ref_tagged_shapes_encode(int num_shapes, const int8_t* dim, const uint32_t* scb, const uint32_t* cvb, const double* v, int hint,
                                uint8_t* out, size_t capacity) {
  s2coding::StringVectorEncoder shape_vector;
  for (int s = 0; s < num_shapes; ++s) {
    auto shape = (hint & 2) ? SoupToClassicShape(s, dim, scb, cvb, v) : SoupToLaxShape(s, dim, scb, cvb, v);
    hint &= ~0;
    Encoder* sub = shape_vector.AddViaEncoder();
    sub->Ensure(Encoder::kVarintMax32);
    sub->put_varint32(shape->type_tag());
    shape->Encode(sub, (hint & 1) ? s2coding::CodingHint::COMPACT : s2coding::CodingHint::FAST);
  }
  Encoder enc;
  shape_vector.Encode(&enc);
  if (out && enc.length() <= capacity) memcpy(out, enc.base(), enc.length());
  return enc.length();
}
// Decodes tagged shapes with the reference's own Init methods and writes every edge of every shape (v0, v1) in edge-id
// order: edges_out holds 6 doubles per edge, num_edges_out[s] the edge count of shape s, dims_out[s] its dimension.
// Returns the total number of edges, or (size_t)-1 if the reference rejects the input.
size_t ref_tagged_shapes_decode_edges(const uint8_t* bytes, size_t len, int max_shapes, int32_t* dims_out, uint32_t* num_edges_out,
                                      double* edges_out, size_t edge_capacity, int* num_shapes_out) {
  Decoder dec(bytes, len);
  s2coding::EncodedStringVector sv;
  if (!sv.Init(&dec)) return (size_t)-1;
  *num_shapes_out = (int)sv.size();
  size_t total = 0;
  for (int s = 0; s < (int)sv.size() && s < max_shapes; ++s) {
    Decoder d = sv.GetDecoder(s);
    uint32_t tag;
    std::unique_ptr<S2Shape> shape;
    if (d.avail() == 0) { dims_out[s] = 0; num_edges_out[s] = 0; continue; }
    if (!d.get_varint32(&tag)) return (size_t)-1;
    if (tag == S2PointVectorShape::kTypeTag) { auto p = std::make_unique<S2PointVectorShape>(); if (!p->Init(&d)) return (size_t)-1; shape = std::move(p); }
    else if (tag == S2LaxPolylineShape::kTypeTag) { auto p = std::make_unique<S2LaxPolylineShape>(); if (!p->Init(&d)) return (size_t)-1; shape = std::move(p); }
    else if (tag == S2LaxPolygonShape::kTypeTag) { auto p = std::make_unique<S2LaxPolygonShape>(); if (!p->Init(&d)) return (size_t)-1; shape = std::move(p); }
    else if (tag == S2Polygon::Shape::kTypeTag) { auto p = std::make_unique<S2Polygon::OwningShape>(); if (!p->Init(&d)) return (size_t)-1; shape = std::move(p); }
    else if (tag == S2Polyline::Shape::kTypeTag) { auto p = std::make_unique<S2Polyline::OwningShape>(); if (!p->Init(&d)) return (size_t)-1; shape = std::move(p); }
    else return (size_t)-1;
    dims_out[s] = shape->dimension();
    num_edges_out[s] = (uint32_t)shape->num_edges();
    for (int e = 0; e < shape->num_edges(); ++e, ++total) {
      if (total < edge_capacity) {
        S2Shape::Edge ed = shape->edge(e);
        for (int q = 0; q < 3; ++q) { edges_out[6 * total + q] = ed.v0[q]; edges_out[6 * total + 3 + q] = ed.v1[q]; }
      }
    }
  }
  return total;
}

// ---------------------------------------------------------------- S2PointIndex / S2ClosestPointQuery
struct RefPointIndex { S2PointIndex<int> index; };
void* ref_point_index_create(const double* xyz, size_t m) {
  auto* r = new RefPointIndex();
  for (size_t i = 0; i < m; ++i) r->index.Add(P(xyz + 3 * i), (int)i);
  return r;
}
void ref_point_index_destroy(void* h) { delete (RefPointIndex*)h; }
// FindClosestPoints(PointTarget) with options.max_distance = S1Angle::Radians(r) (negative r = no limit): result size.
void ref_closest_points_count(void* h, const double* targets, size_t n, double max_distance_rad, uint32_t* counts, int threads) {
  auto* index = &((RefPointIndex*)h)->index;
  ParallelFor(n, threads, [&](size_t b, size_t e, int) {
    S2ClosestPointQuery<int> query(index);
    if (max_distance_rad >= 0) query.mutable_options()->set_max_distance(S1Angle::Radians(max_distance_rad));
    for (size_t i = b; i < e; ++i) {
      S2ClosestPointQuery<int>::PointTarget target(P(targets + 3 * i));
      counts[i] = (uint32_t)query.FindClosestPoints(&target).size();
    }
  });
}
// FindClosestPoint(PointTarget): data (-1 if none) and the S1ChordAngle length2 of the result.
void ref_closest_point(void* h, const double* targets, size_t n, double max_distance_rad, int32_t* data, double* length2, int threads) {
  auto* index = &((RefPointIndex*)h)->index;
  ParallelFor(n, threads, [&](size_t b, size_t e, int) {
    S2ClosestPointQuery<int> query(index);
    if (max_distance_rad >= 0) query.mutable_options()->set_max_distance(S1Angle::Radians(max_distance_rad));
    for (size_t i = b; i < e; ++i) {
      S2ClosestPointQuery<int>::PointTarget target(P(targets + 3 * i));
      auto r = query.FindClosestPoint(&target);
      data[i] = r.is_empty() ? -1 : r.data();
      length2[i] = r.is_empty() ? HUGE_VAL : r.distance().length2();
    }
  });
}
// FindClosestPoints with max_results = k: rows of k (data, length2), padded with -1 / inf, and the result counts.
void ref_closest_points(void* h, const double* targets, size_t n, int k, double max_distance_rad, int32_t* data, double* length2, uint32_t* counts, int threads) {
  auto* index = &((RefPointIndex*)h)->index;
  ParallelFor(n, threads, [&](size_t b, size_t e, int) {
    S2ClosestPointQuery<int> query(index);
    query.mutable_options()->set_max_results(k);
    if (max_distance_rad >= 0) query.mutable_options()->set_max_distance(S1Angle::Radians(max_distance_rad));
    for (size_t i = b; i < e; ++i) {
      S2ClosestPointQuery<int>::PointTarget target(P(targets + 3 * i));
      auto r = query.FindClosestPoints(&target);
      counts[i] = (uint32_t)r.size();
      for (int j = 0; j < k; ++j) {
        data[i * k + j] = j < (int)r.size() ? r[j].data() : -1;
        length2[i * k + j] = j < (int)r.size() ? r[j].distance().length2() : HUGE_VAL;
      }
    }
  });
}
// FindClosestPoints for any target kind (0 PointTarget: 3 doubles, 1 EdgeTarget: 6 doubles a, b, 2 CellTarget: uint64 id) with
// max_results = k (k == 0: unlimited, counts only), max_distance (negative = none) and max_error (radians).
void ref_closest_points_target(void* h, int kind, const void* targets, size_t n, int k, double max_distance_rad, double max_error_rad,
                               int32_t* data, double* length2, uint32_t* counts, int threads) {
  auto* index = &((RefPointIndex*)h)->index;
  ParallelFor(n, threads, [&](size_t b, size_t e, int) {
    S2ClosestPointQuery<int> query(index);
    if (k > 0) query.mutable_options()->set_max_results(k);
    if (max_distance_rad >= 0) query.mutable_options()->set_max_distance(S1Angle::Radians(max_distance_rad));
    if (max_error_rad > 0) query.mutable_options()->set_max_error(S1Angle::Radians(max_error_rad));
    for (size_t i = b; i < e; ++i) {
      std::vector<S2ClosestPointQuery<int>::Result> r;
      if (kind == 0) {
        S2ClosestPointQuery<int>::PointTarget target(P((const double*)targets + 3 * i));
        r = query.FindClosestPoints(&target);
      } else if (kind == 1) {
        S2ClosestPointQuery<int>::EdgeTarget target(P((const double*)targets + 6 * i), P((const double*)targets + 6 * i + 3));
        r = query.FindClosestPoints(&target);
      } else {
        S2ClosestPointQuery<int>::CellTarget target(S2Cell(S2CellId(((const uint64_t*)targets)[i])));
        r = query.FindClosestPoints(&target);
      }
      counts[i] = (uint32_t)r.size();
      for (int j = 0; j < k; ++j) {
        data[i * k + j] = j < (int)r.size() ? r[j].data() : -1;
        length2[i * k + j] = j < (int)r.size() ? r[j].distance().length2() : HUGE_VAL;
      }
    }
  });
}
// FindClosestPoints(ShapeIndexTarget(&shape index)) with include_interiors, max_results = k (0: unlimited), max_distance
// (negative = none), max_error. Returns the number of results; the first `capacity` are written, nearest first.
size_t ref_closest_points_shape_index(void* h, void* index_h, int include_interiors, int k, double max_distance_rad, double max_error_rad,
                                      int32_t* data, double* length2, size_t capacity) {
  auto* index = &((RefPointIndex*)h)->index;
  S2ClosestPointQuery<int> query(index);
  if (k > 0) query.mutable_options()->set_max_results(k);
  if (max_distance_rad >= 0) query.mutable_options()->set_max_distance(S1Angle::Radians(max_distance_rad));
  if (max_error_rad > 0) query.mutable_options()->set_max_error(S1Angle::Radians(max_error_rad));
  S2ClosestPointQuery<int>::ShapeIndexTarget target(&((RefIndex*)index_h)->index);
  target.set_include_interiors(include_interiors != 0);
  const auto r = query.FindClosestPoints(&target);
  for (size_t j = 0; j < r.size() && j < capacity; ++j) {
    data[j] = r[j].data();
    length2[j] = r[j].distance().length2();
  }
  return r.size();
}
// S2CellId::child(k), advance / advance_wrap, GetCommonAncestorLevel element-wise (valid ids only).
void ref_cellid_children(const uint64_t* ids, size_t n, uint64_t* out4) {
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) out4[4 * i + k] = S2CellId(ids[i]).child(k).id();
}
void ref_cellid_advance(const uint64_t* ids, const int64_t* steps, size_t n, int wrap, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = (wrap ? S2CellId(ids[i]).advance_wrap(steps[i]) : S2CellId(ids[i]).advance(steps[i])).id();
}
void ref_cellid_next_prev(const uint64_t* ids, size_t n, uint64_t* next, uint64_t* prev, uint64_t* next_wrap, uint64_t* prev_wrap) {
  for (size_t i = 0; i < n; ++i) {
    const S2CellId c(ids[i]);
    next[i] = c.next().id(); prev[i] = c.prev().id(); next_wrap[i] = c.next_wrap().id(); prev_wrap[i] = c.prev_wrap().id();
  }
}
void ref_cellid_common_ancestor_level(const uint64_t* a, const uint64_t* b, size_t n, int8_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = (int8_t)S2CellId(a[i]).GetCommonAncestorLevel(S2CellId(b[i]));
}
// S2ClosestEdgeQuery(&index).GetDistance(PointTarget(p)).length2() per point (HUGE_VAL for S1ChordAngle::Infinity()).
void ref_index_point_distances(void* index_h, const double* xyz, size_t n, int include_interiors, double max_distance_rad, double* out, int threads) {
  auto& index = ((RefIndex*)index_h)->index;
  ParallelFor(n, threads, [&](size_t b, size_t e, int) {
    S2ClosestEdgeQuery query(&index);
    query.mutable_options()->set_include_interiors(include_interiors != 0);
    if (max_distance_rad >= 0) query.mutable_options()->set_max_distance(S1Angle::Radians(max_distance_rad));
    for (size_t i = b; i < e; ++i) {
      S2ClosestEdgeQuery::PointTarget target(P(xyz + 3 * i));
      const S1ChordAngle d = query.GetDistance(&target);
      out[i] = d.is_infinity() ? HUGE_VAL : d.length2();
    }
  });
}
double ref_chord_length2(double angle_rad) { return S1ChordAngle(S1Angle::Radians(angle_rad)).length2(); }

// S2CellUnion::Contains(S2CellId) (bit0) and ::Intersects(S2CellId) (bit1).
void ref_cellunion_contains_cells(const uint64_t* u, size_t m, const uint64_t* ids, size_t n, uint8_t* out) {
  std::vector<S2CellId> v(m);
  for (size_t i = 0; i < m; ++i) v[i] = S2CellId(u[i]);
  S2CellUnion cu = S2CellUnion::FromVerbatim(std::move(v));
  for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)((cu.Contains(S2CellId(ids[i])) ? 1 : 0) | (cu.Intersects(S2CellId(ids[i])) ? 2 : 0));
}

size_t ref_cellunion_normalize(const uint64_t* ids, size_t m, uint64_t* out) {
  std::vector<S2CellId> v(m);
  for (size_t i = 0; i < m; ++i) v[i] = S2CellId(ids[i]);
  S2CellUnion u(std::move(v));
  for (size_t i = 0; i < u.cell_ids().size(); ++i) out[i] = u.cell_ids()[i].id();
  return u.cell_ids().size();
}

// S2RegionCoverer covering of a cap (centre, radius in radians); returns number of cells (≤ capacity written).
size_t ref_cover_cap(const double* center, double radius_rad, int max_cells, int min_level, int max_level,
                     uint64_t* out, size_t capacity) {
  S2RegionCoverer::Options opt;
  opt.set_max_cells(max_cells);
  if (min_level >= 0) opt.set_min_level(min_level);
  if (max_level >= 0) opt.set_max_level(max_level);
  S2RegionCoverer coverer(opt);
  S2Cap cap(P(center).Normalize(), S1Angle::Radians(radius_rad));
  S2CellUnion u = coverer.GetCovering(cap);
  size_t k = 0;
  for (S2CellId id : u.cell_ids()) { if (k < capacity) out[k] = id.id(); ++k; }
  return k;
}

// S2::RobustCrossProd (s2edge_crossings.cc:147-177) element-wise, all of its stages as this platform runs them.
void ref_robust_cross_prod(const double* a, const double* b, size_t n, double* out) {
  for (size_t i = 0; i < n; ++i) {
    const S2Point r = S2::RobustCrossProd(P(a + 3 * i), P(b + 3 * i));
    out[3 * i] = r[0]; out[3 * i + 1] = r[1]; out[3 * i + 2] = r[2];
  }
}

// S2RegionCoverer(max_cells).GetCovering(S2ShapeIndexRegion(index)) with default levels (s2region_coverer.h:106):
// BASELINE configs[4]'s "S2RegionCoverer covering (max 1000 cells)". Returns the number of cells.
size_t ref_index_covering(void* h, int max_cells, uint64_t* out, size_t capacity) {
  auto& index = ((RefIndex*)h)->index;
  S2ShapeIndexRegion<MutableS2ShapeIndex> region(&index);
  S2RegionCoverer::Options opt;
  opt.set_max_cells(max_cells);
  S2RegionCoverer coverer(opt);
  S2CellUnion u = coverer.GetCovering(region);
  size_t k = 0;
  for (S2CellId id : u.cell_ids()) { if (k < capacity) out[k] = id.id(); ++k; }
  return k;
}

// S2RegionCoverer::GetCovering(S2ShapeIndexRegion) with min_level = max_level = level; returns the number of cells.
size_t ref_region_covering(void* h, int level, uint64_t* out, size_t capacity) {
  auto& index = ((RefIndex*)h)->index;
  S2ShapeIndexRegion<MutableS2ShapeIndex> region(&index);
  S2RegionCoverer::Options opt;
  opt.set_min_level(level);
  opt.set_max_level(level);
  S2RegionCoverer coverer(opt);
  S2CellUnion u = coverer.GetCovering(region);
  size_t k = 0;
  for (S2CellId id : u.cell_ids()) { if (k < capacity) out[k] = id.id(); ++k; }
  return k;
}

// ---------------------------------------------------------------- S2CellIndex (s2cell_index.h) / S2RegionSharder
struct RefCellIndex { S2CellIndex index; };
void* ref_cell_index_create(const uint64_t* ids, const int32_t* labels, size_t m) {
  auto* r = new RefCellIndex();
  for (size_t i = 0; i < m; ++i) r->index.Add(S2CellId(ids[i]), labels[i]);
  r->index.Build();
  return r;
}
void ref_cell_index_destroy(void* h) { delete (RefCellIndex*)h; }
// VisitIntersectingCells for T target unions (CSR: ids[offsets[t]..offsets[t+1])); results CSR, each target's pairs sorted by
// (cell id, label) so the visiting order does not matter. Returns the total (only the first `capacity` pairs are written).
uint64_t ref_cell_index_visit(void* h, const uint64_t* target_ids, const uint64_t* offsets, size_t T, uint64_t* out_offsets,
                              uint64_t* out_cells, int32_t* out_labels, uint64_t capacity) {
  const S2CellIndex& index = ((RefCellIndex*)h)->index;
  uint64_t total = 0;
  out_offsets[0] = 0;
  std::vector<S2CellIndex::LabelledCell> got;
  for (size_t t = 0; t < T; ++t) {
    std::vector<S2CellId> v;
    for (uint64_t k = offsets[t]; k < offsets[t + 1]; ++k) v.push_back(S2CellId(target_ids[k]));
    S2CellUnion target = S2CellUnion::FromVerbatim(std::move(v));
    got.clear();
    index.VisitIntersectingCells(target, [&](S2CellId id, S2CellIndex::Label label) {
      got.push_back({id, label});
      return true;
    });
    std::sort(got.begin(), got.end());
    for (const auto& lc : got) {
      if (total < capacity) { out_cells[total] = lc.cell_id.id(); out_labels[total] = lc.label; }
      ++total;
    }
    out_offsets[t + 1] = total;
  }
  return total;
}
// GetIntersectingLabels per target union: distinct labels, ascending.
uint64_t ref_cell_index_labels(void* h, const uint64_t* target_ids, const uint64_t* offsets, size_t T, uint64_t* out_offsets,
                               int32_t* out_labels, uint64_t capacity) {
  const S2CellIndex& index = ((RefCellIndex*)h)->index;
  uint64_t total = 0;
  out_offsets[0] = 0;
  for (size_t t = 0; t < T; ++t) {
    std::vector<S2CellId> v;
    for (uint64_t k = offsets[t]; k < offsets[t + 1]; ++k) v.push_back(S2CellId(target_ids[k]));
    S2CellUnion target = S2CellUnion::FromVerbatim(std::move(v));
    auto labels = index.GetIntersectingLabels(target);
    std::vector<int32_t> sorted(labels.begin(), labels.end());
    std::sort(sorted.begin(), sorted.end());
    for (int32_t l : sorted) { if (total < capacity) out_labels[total] = l; ++total; }
    out_offsets[t + 1] = total;
  }
  return total;
}
// Per-label histogram of the (point, indexed cell) pairs VisitIntersectingCells reports for the leaf cell of every point.
void ref_cell_index_point_histogram(void* h, const double* xyz, size_t n, uint64_t* counts, int num_labels, int threads) {
  const S2CellIndex& index = ((RefCellIndex*)h)->index;
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  std::vector<std::vector<uint64_t>> local(T, std::vector<uint64_t>(num_labels, 0));
  ParallelFor(n, T, [&](size_t b, size_t e, int tid) {
    auto& c = local[tid];
    for (size_t i = b; i < e; ++i) {
      S2CellUnion target = S2CellUnion::FromVerbatim({S2CellId(P(xyz + 3 * i))});
      index.VisitIntersectingCells(target, [&](S2CellId, S2CellIndex::Label label) { ++c[label]; return true; });
    }
  });
  for (int l = 0; l < num_labels; ++l) { counts[l] = 0; for (int t = 0; t < T; ++t) counts[l] += local[t][l]; }
}
// S2RegionSharder over shard coverings (CSR) queried with S2CellUnion regions (CSR): GetMostIntersectingShard.
void ref_sharder_most_intersecting(const uint64_t* shard_ids, const uint64_t* shard_offsets, size_t num_shards,
                                   const uint64_t* region_ids, const uint64_t* region_offsets, size_t T, int default_shard,
                                   int32_t* out) {
  std::vector<S2CellUnion> shards;
  for (size_t s = 0; s < num_shards; ++s) {
    std::vector<S2CellId> v;
    for (uint64_t k = shard_offsets[s]; k < shard_offsets[s + 1]; ++k) v.push_back(S2CellId(shard_ids[k]));
    shards.push_back(S2CellUnion::FromVerbatim(std::move(v)));
  }
  S2RegionSharder sharder(shards);
  for (size_t t = 0; t < T; ++t) {
    std::vector<S2CellId> v;
    for (uint64_t k = region_offsets[t]; k < region_offsets[t + 1]; ++k) v.push_back(S2CellId(region_ids[k]));
    S2CellUnion region = S2CellUnion::FromVerbatim(std::move(v));
    out[t] = sharder.GetMostIntersectingShard(region, default_shard);
  }
}

// The covering S2RegionSharder derives from an S2CellUnion region: S2CellUnion(region.GetCellUnionBound()) (normalized).
size_t ref_cellunion_cell_union_bound(const uint64_t* ids, size_t m, uint64_t* out, size_t capacity) {
  std::vector<S2CellId> v(m);
  for (size_t i = 0; i < m; ++i) v[i] = S2CellId(ids[i]);
  S2CellUnion region = S2CellUnion::FromVerbatim(std::move(v));
  std::vector<S2CellId> bound;
  region.GetCellUnionBound(&bound);
  S2CellUnion covering(std::move(bound));
  size_t k = 0;
  for (S2CellId id : covering.cell_ids()) { if (k < capacity) out[k] = id.id(); ++k; }
  return k;
}

}  // extern "C"
#pragma GCC visibility pop
