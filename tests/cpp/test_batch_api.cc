// tests/cpp/test_batch_api.cc — exercises include/s2b/s2b_batch.h (the C++ sugar over the C ABI) the way a C++ caller of
// the reference would use the batch entry points. Needs a GPU to run; tests/test_cpp_api.py compiles it everywhere.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime_api.h>

#include "s2b/s2b_batch.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s (%s)\n", __FILE__, __LINE__, #c, s2b::LastError()); return 1; } } while (0)

static s2b::Point FromDegrees(double lat, double lng) {
  const double d = M_PI / 180;
  return s2b::Point{std::cos(lng * d) * std::cos(lat * d), std::sin(lng * d) * std::cos(lat * d), std::sin(lat * d)};
}

int main() {
  // S2CellId(S2Point(1,2,3)) and S2CellId(S2LatLng::FromDegrees(39,-120))  (s2region_test.cc:66-68)
  std::vector<s2b::Point> pts = {{1, 2, 3}};
  std::vector<uint64_t> ids;
  CHECK(s2b::S2CellIdBatch::FromPoints(pts, &ids));
  CHECK(ids[0] == 0x43CC5DF3E09213F5ull);
  s2b::LatLng ll{39 * M_PI / 180, -120 * M_PI / 180};
  uint64_t id2 = 0;
  CHECK(s2b::S2CellIdBatch::FromLatLngs(&ll, 1, &id2));
  CHECK(id2 == 0x809984952A960863ull);
  uint64_t parent = 0;
  CHECK(s2b::S2CellIdBatch::Parent(&id2, 1, 10, &parent));
  CHECK((parent & 0x1FFFFFFFFFFull) == (1ull << 40));   // lsb of a level-10 cell is bit 2*(30-10)
  auto tok = s2b::S2CellIdBatch::ToTokens({0x487604C489F841C3ull, 0x4870000000000000ull, 0});
  CHECK(tok.size() == 3 && tok[0] == "487604c489f841c3" && tok[1] == "487" && tok[2] == "X");

  // index "0:0 # -1:1, 1:1 # 0:5, 0:7, 2:6" (s2contains_point_query_test.cc:55-123) built by the library's own builder
  std::vector<s2b::Point> verts = {FromDegrees(0, 0), FromDegrees(-1, 1), FromDegrees(1, 1), FromDegrees(0, 5), FromDegrees(0, 7), FromDegrees(2, 6)};
  int8_t dims[3] = {0, 1, 2};
  uint32_t shape_chain_begin[4] = {0, 1, 2, 3};
  uint32_t chain_vertex_begin[4] = {0, 1, 3, 6};
  s2b::DeviceShapeIndex index;
  CHECK(index.InitFromShapes(3, dims, shape_chain_begin, chain_vertex_begin, &verts[0].x));
  CHECK(index.num_shape_ids() == 3);
  std::vector<s2b::Point> q = {FromDegrees(0, 0), FromDegrees(-1, 1), FromDegrees(1, 1), FromDegrees(0, 2), FromDegrees(0, 5),
                               FromDegrees(0, 7), FromDegrees(2, 6), FromDegrees(1, 6), FromDegrees(10, 10)};
  const uint8_t want_open[9] = {0, 0, 0, 0, 0, 0, 0, 1, 0}, want_semi[9] = {0, 0, 0, 0, 0, 1, 0, 1, 0}, want_closed[9] = {1, 1, 1, 0, 1, 1, 1, 1, 0};
  std::vector<uint8_t> out(q.size());
  CHECK(s2b::S2ContainsPointQueryBatch(&index, s2b::S2VertexModel::OPEN).Contains(q.data(), q.size(), out.data()));
  CHECK(std::memcmp(out.data(), want_open, 9) == 0);
  CHECK(s2b::S2ContainsPointQueryBatch(&index, s2b::S2VertexModel::SEMI_OPEN).Contains(q.data(), q.size(), out.data()));
  CHECK(std::memcmp(out.data(), want_semi, 9) == 0);
  s2b::S2ContainsPointQueryBatch closed(&index, s2b::S2VertexModel::CLOSED);
  CHECK(closed.Contains(q.data(), q.size(), out.data()));
  CHECK(std::memcmp(out.data(), want_closed, 9) == 0);
  CHECK(closed.ShapeContains(2, q.data(), q.size(), out.data()));
  CHECK(out[4] == 1 && out[7] == 1 && out[0] == 0);
  std::vector<uint64_t> counts;
  CHECK(closed.CountContainingShapes(q.data(), q.size(), &counts));
  CHECK(counts.size() == 3 && counts[0] == 1 && counts[1] == 2 && counts[2] == 4);
  std::vector<uint32_t> offsets;
  std::vector<int32_t> shape_ids;
  CHECK(closed.GetContainingShapeIds(q.data(), q.size(), &offsets, &shape_ids));
  CHECK(offsets.size() == 10 && offsets[9] == 7 && shape_ids.size() == 7 && shape_ids[0] == 0);
  {
    // S2ShapeIndexRegion over the same index: the six face cells cover the sphere, so some of them must intersect; a leaf cell
    // far below an index cell is contained by a polygon exactly when MayIntersect and the point query agree.
    s2b::S2ShapeIndexRegionBatch region(&index);
    std::vector<uint64_t> faces;
    for (uint64_t f = 0; f < 6; ++f) faces.push_back((f << 61) | (1ull << 60));
    std::vector<uint8_t> may(6), con(6);
    CHECK(region.MayIntersect(faces.data(), faces.size(), may.data()));
    CHECK(region.Contains(faces.data(), faces.size(), con.data()));
    int any = 0;
    for (int f = 0; f < 6; ++f) { any += may[f]; CHECK(!con[f] || may[f]); }
    CHECK(any > 0);
    std::vector<uint32_t> roff; std::vector<int32_t> rshape; std::vector<uint8_t> rcon;
    CHECK(region.GetIntersectingShapeIds(faces.data(), faces.size(), &roff, &rshape, &rcon));
    CHECK(roff.size() == 7 && roff[6] == rshape.size() && rshape.size() == rcon.size() && !rshape.empty());
    for (int f = 0; f < 6; ++f) CHECK((roff[f + 1] > roff[f]) == (may[f] != 0));
  }

  // serialized round trip: geometry + index to the reference's wire formats, read back, same answers
  {
    S2bBuiltIndex* built = nullptr;
    CHECK(s2b_index_build(3, dims, shape_chain_begin, chain_vertex_begin, &verts[0].x, S2B_MAX_EDGES_PER_CELL_DEVICE, &built) == S2B_OK);
    S2bFlatIndex flat;
    const int32_t* edge_ids = nullptr;
    CHECK(s2b_built_index_flat(built, &flat, &edge_ids) == S2B_OK);
    size_t need = 0, need_shapes = 0;
    CHECK(s2b_index_encode(&flat, edge_ids, S2B_MAX_EDGES_PER_CELL_DEVICE, nullptr, 0, &need) == S2B_ECAPACITY && need > 0);
    std::vector<uint8_t> blob(need);
    CHECK(s2b_index_encode(&flat, edge_ids, S2B_MAX_EDGES_PER_CELL_DEVICE, blob.data(), blob.size(), &need) == S2B_OK);
    CHECK(s2b_shapes_encode(3, dims, shape_chain_begin, chain_vertex_begin, &verts[0].x, nullptr, 0, &need_shapes) == S2B_ECAPACITY);
    std::vector<uint8_t> shape_blob(need_shapes);
    CHECK(s2b_shapes_encode(3, dims, shape_chain_begin, chain_vertex_begin, &verts[0].x, shape_blob.data(), shape_blob.size(), &need_shapes) == S2B_OK);
    s2b_built_index_destroy(built);
    S2bShapeSoup* soup = nullptr;
    CHECK(s2b_shapes_decode(shape_blob.data(), shape_blob.size(), &soup, nullptr) == S2B_OK);
    int ns = 0; const int8_t* d2; const uint32_t *scb2, *cvb2; const double* v2; size_t nv = 0;
    CHECK(s2b_shape_soup_arrays(soup, &ns, &d2, &scb2, &cvb2, &v2, &nv) == S2B_OK && ns == 3 && nv == 6);
    s2b::DeviceShapeIndex index2;
    CHECK(index2.InitFromEncoded(blob.data(), blob.size(), ns, d2, scb2, cvb2, v2));
    s2b_shape_soup_destroy(soup);
    CHECK(s2b::S2ContainsPointQueryBatch(&index2, s2b::S2VertexModel::CLOSED).Contains(q.data(), q.size(), out.data()));
    CHECK(std::memcmp(out.data(), want_closed, 9) == 0);
    uint8_t garbage[5] = {0xFF, 0xFF, 0xFF, 0xFF, 0x7F};
    CHECK(!index2.InitFromEncoded(garbage, 5, ns, dims, shape_chain_begin, chain_vertex_begin, &verts[0].x));
  }

  // S2PointIndex + S2ClosestPointQuery: the six vertices as an index
  {
    S2bPointIndex* pi = nullptr;
    CHECK(s2b_point_index_create(&verts[0].x, verts.size(), &pi) == S2B_OK && s2b_point_index_size(pi) == 6);
    s2b::Point t = FromDegrees(0.1, 5.1);
    int32_t data = -2; double l2 = -1; uint32_t cnt = 99;
    CHECK(s2b_closest_point(pi, &t.x, 1, INFINITY, &data, &l2) == S2B_OK && data == 3 && l2 > 0 && l2 < 1e-4);
    CHECK(s2b_closest_points_count(pi, &t.x, 1, 4.0 * std::sin(0.5 * 0.05) * std::sin(0.5 * 0.05), &cnt) == S2B_OK && cnt == 3);   // within 0.05 rad: 0:5, 0:7, 2:6
    s2b_point_index_destroy(pi);
  }

  // S2CellIndex: IntersectionOptimization scene (s2cell_index_test.cc:387-398): 1/001 -> 1, 1/333 -> 2, 2/00 -> 3, 2/0232 -> 4
  {
    auto child = [](uint64_t id, int k) { uint64_t nl = (id & (~id + 1)) >> 2; return id + (uint64_t)(2 * k + 1 - 4) * nl; };
    auto cell = [&](int face, const char* digits) { uint64_t id = ((uint64_t)face << 61) | (1ull << 60); for (const char* d = digits; *d; ++d) id = child(id, *d - '0'); return id; };
    s2b::S2CellIndexBatch ci;
    ci.Add(cell(1, "001"), 1); ci.Add(cell(1, "333"), 2); ci.Add(cell(2, "00"), 3); ci.Add(cell(2, "0232"), 4);
    CHECK(ci.Build());
    std::vector<uint64_t> targets = {cell(1, "010"), cell(1, "3"), cell(2, "010"), cell(2, "011"), cell(2, "02")};
    std::vector<uint32_t> toff = {0, 2, 5}, off;
    std::vector<int32_t> labels;
    CHECK(ci.GetIntersectingLabels(targets, toff, &off, &labels));
    CHECK(off.size() == 3 && off[1] == 1 && off[2] == 2 && labels.size() == 2 && labels[0] == 2 && labels[1] == 4);   // the reference's answer
    std::vector<uint64_t> cells_out;
    CHECK(ci.GetIntersectingCells(targets, toff, &off, &cells_out, &labels) && cells_out.size() == 2 && cells_out[0] == cell(1, "333") &&
          cells_out[1] == cell(2, "0232"));
    std::vector<uint64_t> counts;
    s2b::Point on_face2 = {0.0, 0.0, 1.0};   // the centre of face 2 lies in none of the four cells
    CHECK(ci.CountPointsPerLabel(&on_face2, 1, &counts) && counts.size() == 5 && counts[1] + counts[2] + counts[3] + counts[4] == 0);
  }

  // S2PointIndex + S2ClosestPointQuery through the C++ sugar: five points on the equator, point / edge / cell targets and a
  // shape-index target (a small square around (0, 0)); answers follow from the geometry.
  {
    s2b::S2PointIndexBatch pi;
    for (int k = 0; k < 5; ++k) pi.Add(FromDegrees(0.0, 10.0 * k));           // data 0..4 at longitudes 0, 10, 20, 30, 40
    CHECK(pi.Build() && pi.num_points() == 5);
    s2b::S2ClosestPointQueryBatch q(&pi);
    q.set_max_results(2);
    std::vector<int32_t> data; std::vector<double> l2; std::vector<uint32_t> cnt;
    s2b::Point t = FromDegrees(1.0, 19.0);
    CHECK(q.FindClosestPoints(&t, 1, &data, &l2, &cnt) && cnt[0] == 2 && data[0] == 2 && data[1] == 1 && l2[0] < l2[1]);
    s2b::Point edge[2] = {FromDegrees(5.0, 28.0), FromDegrees(5.0, 41.0)};     // runs above points 3 and 4
    CHECK(q.FindClosestPointsToEdges(edge, 1, &data, &l2, &cnt) && cnt[0] == 2 && ((data[0] == 3 && data[1] == 4) || (data[0] == 4 && data[1] == 3)));
    std::vector<uint64_t> leaf;
    std::vector<s2b::Point> at = {FromDegrees(0.0, 10.0)};
    CHECK(s2b::S2CellIdBatch::FromPoints(at, &leaf));
    const uint64_t lsb10 = 1ull << (2 * (30 - 10));
    const uint64_t cell10 = (leaf[0] & (~lsb10 + 1)) | lsb10;                  // the level-10 cell containing point 1
    CHECK(q.FindClosestPointsToCells(&cell10, 1, &data, &l2, &cnt) && cnt[0] == 2 && data[0] == 1 && l2[0] == 0.0);
    std::vector<s2b::Point> sq = {FromDegrees(-1, -1), FromDegrees(-1, 1), FromDegrees(1, 1), FromDegrees(1, -1)};
    const int8_t dim2[] = {2};
    const uint32_t scb1[] = {0, 1}, cvb1[] = {0, 4};
    s2b::DeviceShapeIndex square;
    CHECK(square.InitFromShapes(1, dim2, scb1, cvb1, &sq[0].x));
    q.set_max_results(0);
    q.set_max_distance_length2(0.04);                                           // chord^2 of ~11.5 degrees: points 0 and 1
    CHECK(q.FindClosestPointsToShapeIndex(square, true, &data, &l2) && data.size() == 2 && data[0] == 0 && l2[0] == 0.0 && data[1] == 1 && l2[1] > 0.0);
    CHECK(q.FindClosestPointsToShapeIndex(square, false, &data, &l2) && data.size() == 2 && data[0] == 0 && l2[0] > 0.0);   // the boundary is 1 degree away
  }

  // Multi-GPU driver behind the C ABI: shard over every visible GPU (2 on the test box; 1 still runs every code path but
  // the exchange), results must equal the single-GPU calls; both reductions (NCCL and the library's peer-memory kernel).
  {
    const int ngpu = s2b_device_count() < 8 ? s2b_device_count() : 8;
    CHECK(ngpu >= 1);
    std::vector<int> devices;
    for (int d = 0; d < ngpu; ++d) devices.push_back(d);
    // a scene with some weight: 300 small squares on a grid, 3,000,001 pseudo-random points (splitmix-style hash)
    std::vector<s2b::Point> v;
    std::vector<int8_t> dim; std::vector<uint32_t> scb = {0}, cvb = {0};
    for (int a = 0; a < 20; ++a)
      for (int b = 0; b < 15; ++b) {
        const double lat = -60 + 8.0 * b, lng = -170 + 17.0 * a, h = 3.0;
        v.push_back(FromDegrees(lat, lng)); v.push_back(FromDegrees(lat, lng + h)); v.push_back(FromDegrees(lat + h, lng + h)); v.push_back(FromDegrees(lat + h, lng));
        dim.push_back(2); scb.push_back((uint32_t)dim.size()); cvb.push_back((uint32_t)v.size());
      }
    s2b::DeviceShapeIndex big;
    CHECK(big.InitFromShapes((int)dim.size(), dim.data(), scb.data(), cvb.data(), &v[0].x));
    const size_t n = 3000001;
    std::vector<s2b::Point> p(n);
    std::vector<s2b::LatLng> pll(n);
    uint64_t st = 12345;
    auto next = [&] { uint64_t z = (st += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return (double)((z ^ (z >> 31)) >> 11) * (1.0 / 9007199254740992.0); };
    for (size_t i = 0; i < n; ++i) {
      const double z = 2 * next() - 1, t = 2 * M_PI * next(), r = std::sqrt(1 - z * z);
      p[i] = s2b::Point{r * std::cos(t), r * std::sin(t), z};
      pll[i] = s2b::LatLng{std::asin(z), t - M_PI};
    }
    s2b::S2ContainsPointQueryBatch query(&big);
    std::vector<uint64_t> single, multi;
    std::vector<uint8_t> in1(n), inN(n);
    CHECK(query.CountContainingShapes(p.data(), n, &single));         // no replicas yet: the single-GPU path
    CHECK(query.Contains(p.data(), n, in1.data()));
    CHECK(big.Replicate(devices) && big.num_replicas() == ngpu);
    uint64_t hits = 0;
    for (uint64_t c : single) hits += c;
    CHECK(hits > 1000);
    for (int mode : {S2B_REDUCE_NCCL, S2B_REDUCE_PEER}) {
      if (ngpu > 1) {
        int rc = s2b_index_set_reduce_mode(const_cast<S2bIndex*>(big.handle()), mode);
        if (rc == S2B_EUNSUPPORTED) { std::printf("reduce mode %d unavailable: %s\n", mode, s2b::LastError()); continue; }
        CHECK(rc >= 0);
      }
      CHECK(query.CountContainingShapes(p.data(), n, &multi));
      CHECK(multi == single);
      // device-resident shards: the same points, uploaded to each replica's GPU by hand
      int rdev[8];
      CHECK(s2b_index_replica_devices(big.handle(), rdev, 8) == S2B_OK);
      std::vector<double*> d_pts(ngpu);
      std::vector<uint64_t*> d_counts(ngpu);
      std::vector<size_t> cnt(ngpu);
      for (int k = 0; k < ngpu; ++k) {
        const size_t lo = n * k / ngpu, hi = n * (k + 1) / ngpu;   // any partition will do: the histogram is a sum
        cnt[k] = hi - lo;
        CHECK(cudaSetDevice(rdev[k]) == cudaSuccess);
        CHECK(cudaMalloc((void**)&d_pts[k], 24 * cnt[k] + 24) == cudaSuccess && cudaMalloc((void**)&d_counts[k], 8 * single.size()) == cudaSuccess);
        CHECK(cudaMemcpy(d_pts[k], &p[lo].x, 24 * cnt[k], cudaMemcpyHostToDevice) == cudaSuccess);
      }
      for (int rep = 0; rep < 3; ++rep)   // repeated calls reuse the partial buffers: ordering between reads and the next memset
        CHECK(s2b_shape_histogram_multi_dev(big.handle(), (const double* const*)d_pts.data(), cnt.data(), S2B_VERTEX_MODEL_SEMI_OPEN, d_counts.data(), rep == 2) == S2B_OK);
      for (int k = 0; k < ngpu; ++k) {
        std::vector<uint64_t> got(single.size());
        CHECK(cudaSetDevice(rdev[k]) == cudaSuccess);
        CHECK(cudaMemcpy(got.data(), d_counts[k], 8 * got.size(), cudaMemcpyDeviceToHost) == cudaSuccess);
        CHECK(got == single);                                    // every GPU holds the global histogram
        cudaFree(d_pts[k]); cudaFree(d_counts[k]);
      }
      CHECK(cudaSetDevice(0) == cudaSuccess);
    }
    CHECK(query.Contains(p.data(), n, inN.data()));
    CHECK(in1 == inN);
    std::vector<uint64_t> ids1(n), idsN(n);
    CHECK(s2b_encode_points(&p[0].x, n, 30, ids1.data()) == S2B_OK);
    CHECK(s2b_encode_points_multi(devices.data(), ngpu, &p[0].x, n, 30, idsN.data()) == S2B_OK && ids1 == idsN);
    const int levels[3] = {12, 20, 30};
    std::vector<uint64_t> l1(3 * n), lN(3 * n);
    std::vector<char> t1(16 * n), tN(16 * n);
    std::vector<uint8_t> len1(n), lenN(n);
    CHECK(s2b_encode_latlng_levels_tokens(&pll[0].lat_rad, n, levels, 3, l1.data(), 2, t1.data(), len1.data()) == S2B_OK);
    CHECK(s2b_encode_latlng_levels_tokens_multi(devices.data(), ngpu, &pll[0].lat_rad, n, levels, 3, lN.data(), 2, tN.data(), lenN.data()) == S2B_OK);
    CHECK(l1 == lN && t1 == tN && len1 == lenN);
    const int twice[2] = {0, 0};
    CHECK(s2b_encode_points_multi(twice, 2, &p[0].x, n, 30, idsN.data()) == S2B_EINVAL);
    std::printf("multi-GPU driver: %d GPU(s), %llu hits, histograms / flags / ids equal to the single-GPU calls\n", ngpu, (unsigned long long)hits);
  }

  // S2CellUnion::Contains: Trondheim (python/s2geometry_test.py:254-259)
  uint64_t cells[2] = {0x466D319000000000ull, 0x466D31B000000000ull};
  s2b::Point trondheim = FromDegrees(63.431052, 10.395083);
  uint8_t hit = 0;
  CHECK(s2b::S2CellUnionContainsBatch(cells, 2, &trondheim, 1, &hit));
  CHECK(hit == 1);
  std::printf("test_batch_api: all checks passed\n");
  return 0;
}
