// s2b_index_handle.h — the opaque S2bIndex of include/s2b_capi.h, shared by the translation units that query it.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "s2b_contains.cuh"

struct S2bReplicaSet;                            // s2b_multi.cu: the copies of this index on other GPUs + the reduction plumbing
void s2b_replica_set_destroy(S2bReplicaSet* set);

struct S2bIndex {
  int device = 0;
  s2b::FlatIndexView view{};
  const s2b::FlatIndexView* view_dev = nullptr;   // the same view, stored at the end of the device blob
  void* blob = nullptr;   // one allocation holding every array
  size_t blob_bytes = 0;
  uint64_t num_clips = 0, num_edges = 0;
  int32_t* edge_ids = nullptr;   // optional (s2b_index_set_edge_ids): shape-local id of every clipped edge, for VisitIncidentEdges
  S2bReplicaSet* replicas = nullptr;             // set by s2b_index_replicate (owned by the primary index)
  uint64_t extended_region_edges = 0;   // statistic: edges whose RobustCrossProd goes beyond the double-precision stage
};
