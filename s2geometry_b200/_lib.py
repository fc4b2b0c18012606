"""ctypes binding of libs2b.so (the C ABI declared in include/s2b_capi.h).

The library is the ONLY implementation of the batch entry points: if it cannot be loaded, or a call fails,
an exception is raised — nothing falls back to Python/NumPy/the oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S2B_LIB_PATH") or os.path.join(_HERE, "libs2b.so")   # the override is for kernel experiments

_c = ctypes
_vp, _sz, _i, _u64 = _c.c_void_p, _c.c_size_t, _c.c_int, _c.c_uint64

# name -> (restype, argtypes); mirrors include/s2b_capi.h one to one (tests/test_capi_symbols.py checks it).
SIGNATURES = {
    "s2b_last_error": (_c.c_char_p, []),
    "s2b_version": (_c.c_char_p, []),
    "s2b_device_count": (_i, []),
    "s2b_launch_count": (_u64, []),
    "s2b_diag_fp64_peak": (_i, [_vp]),
    "s2b_diag_force_correctly_rounded": (_i, [_i]),
    "s2b_diag_filter_deviation_dev": (_i, [_i, _vp, _sz, _vp, _vp, _vp]),
    "s2b_synth_sphere_points_dev": (_i, [_u64, _u64, _sz, _vp, _vp]),
    "s2b_synth_sphere_latlng_dev": (_i, [_u64, _u64, _sz, _vp, _vp]),
    "s2b_synth_cap_points_dev": (_i, [_u64, _u64, _sz, _vp, _c.c_double, _vp, _vp]),
    "s2b_set_latlng_mode": (_i, [_i]),
    "s2b_latlng_strict_stats": (_i, [_vp, _vp, _i]),
    "s2b_encode_latlng_strict_dev": (_i, [_i, _vp, _sz, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "s2b_encode_points": (_i, [_vp, _sz, _i, _vp]),
    "s2b_encode_points_dev": (_i, [_vp, _sz, _i, _vp, _vp]),
    "s2b_encode_latlng": (_i, [_vp, _sz, _i, _vp, _vp]),
    "s2b_encode_latlng_dev": (_i, [_vp, _sz, _i, _vp, _vp, _vp]),
    "s2b_encode_latlng_e7": (_i, [_vp, _sz, _i, _vp, _vp]),
    "s2b_encode_latlng_e7_dev": (_i, [_vp, _sz, _i, _vp, _vp, _vp]),
    "s2b_encode_latlng_levels": (_i, [_vp, _sz, _vp, _i, _vp]),
    "s2b_encode_latlng_levels_dev": (_i, [_vp, _sz, _vp, _i, _vp, _vp]),
    "s2b_encode_points_levels_dev": (_i, [_vp, _sz, _vp, _i, _vp, _vp]),
    "s2b_encode_latlng_levels_tokens": (_i, [_vp, _sz, _vp, _i, _vp, _i, _vp, _vp]),
    "s2b_encode_latlng_levels_tokens_dev": (_i, [_vp, _sz, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "s2b_cellid_parent": (_i, [_vp, _sz, _i, _vp]),
    "s2b_cellid_parent_dev": (_i, [_vp, _sz, _i, _vp, _vp]),
    "s2b_cellid_to_token": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_cellid_to_token_dev": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "s2b_cellid_edge_neighbors": (_i, [_vp, _sz, _vp]),
    "s2b_cellid_edge_neighbors_dev": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_cellid_to_point": (_i, [_vp, _sz, _vp]),
    "s2b_cellid_to_point_dev": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_cellid_to_latlng": (_i, [_vp, _sz, _vp]),
    "s2b_cellid_to_latlng_dev": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_cellid_from_token": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_cellid_from_token_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_cellid_children": (_i, [_vp, _sz, _vp]),
    "s2b_cellid_children_dev": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_cellid_advance": (_i, [_vp, _vp, _sz, _i, _vp]),
    "s2b_cellid_advance_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp]),
    "s2b_cellid_common_ancestor_level": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_cellid_common_ancestor_level_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_cellid_info": (_i, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "s2b_cellid_info_dev": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "s2b_cellid_vertex_neighbors": (_i, [_vp, _sz, _i, _vp, _vp]),
    "s2b_cellid_vertex_neighbors_dev": (_i, [_vp, _sz, _i, _vp, _vp, _vp]),
    "s2b_cellid_all_neighbors": (_i, [_vp, _sz, _i, _vp, _sz, _vp]),
    "s2b_cellid_all_neighbors_dev": (_i, [_vp, _sz, _i, _vp, _sz, _vp, _vp]),
    "s2b_cellunion_normalize": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_cellunion_normalize_dev": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "s2b_cellunion_contains_cells": (_i, [_vp, _sz, _vp, _sz, _vp]),
    "s2b_cellunion_contains_cells_dev": (_i, [_vp, _sz, _vp, _sz, _vp, _vp]),
    "s2b_point_index_create": (_i, [_vp, _sz, _vp]),
    "s2b_point_index_destroy": (None, [_vp]),
    "s2b_point_index_size": (_sz, [_vp]),
    "s2b_closest_points_count": (_i, [_vp, _vp, _sz, ctypes.c_double, _vp]),
    "s2b_closest_points_count_dev": (_i, [_vp, _vp, _sz, ctypes.c_double, _vp, _vp]),
    "s2b_closest_points": (_i, [_vp, _vp, _sz, _i, ctypes.c_double, _vp, _vp, _vp]),
    "s2b_closest_points_dev": (_i, [_vp, _vp, _sz, _i, ctypes.c_double, _vp, _vp, _vp, _vp]),
    "s2b_closest_point": (_i, [_vp, _vp, _sz, ctypes.c_double, _vp, _vp]),
    "s2b_closest_point_dev": (_i, [_vp, _vp, _sz, ctypes.c_double, _vp, _vp, _vp]),
    "s2b_closest_points_to_targets": (_i, [_vp, _i, _vp, _sz, _i, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp]),
    "s2b_closest_points_to_targets_dev": (_i, [_vp, _i, _vp, _sz, _i, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp]),
    "s2b_closest_points_to_shape_index": (_i, [_vp, _vp, _i, _i, ctypes.c_double, ctypes.c_double, _vp, _vp, _sz, _vp]),
    "s2b_index_point_distances": (_i, [_vp, _vp, _sz, _i, ctypes.c_double, _vp]),
    "s2b_index_point_distances_dev": (_i, [_vp, _vp, _sz, _i, ctypes.c_double, _vp]),
    "s2b_cell_index_create": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_cell_index_destroy": (None, [_vp]),
    "s2b_cell_index_info": (_i, [_vp, _vp, _vp, _vp]),
    "s2b_cell_index_intersecting_cells": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "s2b_cell_index_intersecting_cells_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp]),
    "s2b_cell_index_intersecting_labels": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp]),
    "s2b_cell_index_intersecting_labels_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "s2b_cell_index_most_intersecting_label": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _c.c_int32, _vp]),
    "s2b_cell_index_most_intersecting_label_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _c.c_int32, _vp, _vp]),
    "s2b_cell_index_point_histogram": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_cell_index_point_histogram_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_index_build": (_i, [_i, _vp, _vp, _vp, _vp, _i, _vp]),
    "s2b_built_index_flat": (_i, [_vp, _vp, _vp]),
    "s2b_built_index_destroy": (None, [_vp]),
    "s2b_shapes_decode": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_shape_soup_arrays": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2b_shape_soup_destroy": (None, [_vp]),
    "s2b_shapes_encode": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "s2b_shapes_encode_hint": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "s2b_index_decode": (_i, [_vp, _sz, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2b_index_encode": (_i, [_vp, _vp, _i, _vp, _sz, _vp]),
    "s2b_index_create": (_i, [_vp, _vp]),
    "s2b_index_create_from_cellunion": (_i, [_vp, _sz, _vp]),
    "s2b_index_destroy": (None, [_vp]),
    "s2b_index_set_edge_ids": (_i, [_vp, _vp, _sz]),
    "s2b_incident_edges": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "s2b_incident_edges_count_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_incident_edges_fill_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "s2b_index_locate_points": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_index_locate_points_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_index_locate_cells": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_index_locate_cells_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "s2b_region_may_intersect_cells": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_region_may_intersect_cells_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_region_contains_cells": (_i, [_vp, _vp, _sz, _vp]),
    "s2b_region_contains_cells_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "s2b_region_intersecting_shapes": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp]),
    "s2b_region_cell_union_bound": (_i, [_vp, _vp, _vp]),
    "s2b_region_covering_at_level": (_i, [_vp, _i, _vp, _sz, _vp]),
    "s2b_index_info": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "s2b_contains": (_i, [_vp, _vp, _sz, _i, _vp]),
    "s2b_contains_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp]),
    "s2b_count_flags_dev": (_i, [_vp, _sz, _vp, _vp]),
    "s2b_shape_contains": (_i, [_vp, _i, _vp, _sz, _i, _vp]),
    "s2b_shape_contains_dev": (_i, [_vp, _i, _vp, _sz, _i, _vp, _vp]),
    "s2b_shape_histogram": (_i, [_vp, _vp, _sz, _i, _vp]),
    "s2b_shape_histogram_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp]),
    "s2b_containing_shapes": (_i, [_vp, _vp, _sz, _i, _vp, _vp, _sz, _vp]),
    "s2b_containing_shapes_count_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp]),
    "s2b_containing_shapes_fill_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp, _vp]),
    "s2b_index_replicate": (_i, [_vp, _vp, _i]),
    "s2b_index_num_replicas": (_i, [_vp]),
    "s2b_index_replica_devices": (_i, [_vp, _vp, _i]),
    "s2b_index_set_reduce_mode": (_i, [_vp, _i]),
    "s2b_shape_histogram_multi": (_i, [_vp, _vp, _sz, _i, _vp]),
    "s2b_shape_histogram_multi_dev": (_i, [_vp, _vp, _vp, _i, _vp, _i]),
    "s2b_contains_multi": (_i, [_vp, _vp, _sz, _i, _vp]),
    "s2b_encode_points_multi": (_i, [_vp, _i, _vp, _sz, _i, _vp]),
    "s2b_encode_latlng_levels_tokens_multi": (_i, [_vp, _i, _vp, _sz, _vp, _i, _vp, _i, _vp, _vp]),
    "s2b_exact_stage_count": (_i, [_vp, _i]),
    "s2b_diag_force_exact_locate": (_i, [_i]),
    "s2b_cellunion_contains": (_i, [_vp, _sz, _vp, _sz, _vp]),
    "s2b_cellunion_contains_dev": (_i, [_vp, _sz, _vp, _sz, _vp, _vp]),
}


S2B_OK, S2B_EINVAL, S2B_ECUDA, S2B_ENOMEM, S2B_ECAPACITY, S2B_EUNSUPPORTED = 0, -1, -2, -3, -4, -5   # include/s2b_capi.h


class S2BError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libs2b error %d: %s" % (code, message))
        self.code = code


_lib = None


def load():
    """Loads libs2b.so (built by s2geometry_b200/build.py). Raises if it is missing — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -m s2geometry_b200.build` (needs nvcc). "
            "s2geometry_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise S2BError(rc, load().s2b_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load().s2b_launch_count())
