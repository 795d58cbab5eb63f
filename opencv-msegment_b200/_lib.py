"""ctypes binding of include/msegment.h (libmsegment_b200.so).  No fallback of any kind: if the CUDA
library is missing the import fails loudly; if there is no CUDA device, Context() raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsegment_b200.so")

MSG_OK, MSG_EINVAL, MSG_ECUDA, MSG_ENOMEM, MSG_ESTATE, MSG_ERANGE = 0, -1, -2, -3, -4, -5
LABELS_32S, LABELS_16U = 0, 1
TERM_COUNT, TERM_EPS = 1, 2
MAX_INFLIGHT = 4


class SegmentParams(C.Structure):
    _fields_ = [("sp", C.c_double), ("sr", C.c_double), ("max_level", C.c_int), ("term_type", C.c_int),
                ("max_count", C.c_int), ("eps", C.c_double), ("lo_diff", C.c_int), ("min_size", C.c_int),
                ("color_dist", C.c_int), ("render_depth", C.c_int), ("connectivity", C.c_int), ("labels_type", C.c_int)]


MAX_STRIPS, SHARD_TABLE_HEADER = 64, 192


class ShardPlan(C.Structure):
    _fields_ = [("n_strips", C.c_int), ("width", C.c_int), ("height", C.c_int), ("max_level", C.c_int), ("halo_rows", C.c_int),
                ("row0", C.c_int * 64), ("row1", C.c_int * 64), ("halo0", C.c_int * 64), ("halo1", C.c_int * 64)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "filter_ms", "label_ms", "merge_ms", "render_ms", "d2h_ms", "total_ms")]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("kernel_launches", "ms_overflow_items", "ms_active_items", "merge_rounds",
                                           "h2d_bytes", "d2h_bytes", "staged_bytes")]


class KernelProfile(C.Structure):
    _fields_ = [("launches", C.c_uint64 * 9), ("tile_ms", C.c_double * 9), ("overflow_ms", C.c_double * 9),
                ("tile_tests", C.c_uint64 * 9), ("tile_hits", C.c_uint64 * 9),
                ("overflow_tests", C.c_uint64 * 9), ("overflow_hits", C.c_uint64 * 9)]


# name -> (restype, argtypes); must list every symbol include/msegment.h declares (tests/test_abi.py checks)
_P, _SZ, _I, _D = C.c_void_p, C.c_size_t, C.c_int, C.c_double
SIGNATURES = {
    "msg_version": (_I, []),
    "msg_device_count": (_I, []),
    "msg_create": (_I, [_I, C.POINTER(_P)]),
    "msg_destroy": (None, [_P]),
    "msg_last_error": (C.c_char_p, [_P]),
    "msg_set_stream": (_I, [_P, _P]),
    "msg_synchronize": (_I, [_P]),
    "msg_set_option": (_I, [_P, C.c_char_p, _I]),
    "msg_get_option": (_I, [_P, C.c_char_p, C.POINTER(_I)]),
    "msg_register_host": (_I, [_P, _P, _SZ]),
    "msg_unregister_host": (_I, [_P, _P]),
    "msg_meanshift_filter": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _D, _D, _I, _I, _I, _D]),
    "msg_label_regions": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I, _I, C.POINTER(C.c_int32)]),
    "msg_merge_regions": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I, C.POINTER(C.c_int32)]),
    "msg_connected_components": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, C.POINTER(C.c_int32)]),
    "msg_render_labels": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _P]),
    "msg_watershed": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_watershed_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_watershed_batch_dev": (_I, [_P, _P, _SZ, _SZ, _P, _SZ, _SZ, _I, _I, _I, _P]),
    "msg_laplacian_sharpen": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _P, _I, _I]),
    "msg_bgr2gray": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_median_blur": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I]),
    "msg_canny": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _D, _D]),
    "msg_dilate": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I]),
    "msg_subtract": (_I, [_P, _P, _SZ, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_copy_masked": (_I, [_P, _P, _SZ, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_shape_seeds": (_I, [_P, _P, _SZ, _I, _I, _I, _D, _D, _P, _SZ, _P, _P, _SZ]),
    "msg_shape_seeds_dev": (_I, [_P, _P, _SZ, _I, _I, _I, _D, _D, _P, _SZ, _P, _P]),
    "msg_white_to_black": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_threshold": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _D, _D, _I, C.POINTER(_D)]),
    "msg_distance_transform": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I]),
    "msg_distance_transform_max_width": (_I, [_P]),
    "msg_normalize_minmax": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _D, _D]),
    "msg_threshold_f32": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _D, _D]),
    "msg_dilate_f32": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I]),
    "msg_convert_f32_to_u8": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I]),
    "msg_contour_markers": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, C.POINTER(C.c_int32)]),
    "msg_circle_filled": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _I, C.c_int32]),
    "msg_color_seeds": (_I, [_P, _P, _SZ, _I, _I, _P, _I, _I, _D, _P, _SZ, C.POINTER(C.c_int32), _P, _SZ, _P, _SZ, _P, _SZ,
                             _P, _SZ]),
    "msg_color_seeds_dev": (_I, [_P, _P, _SZ, _I, _I, _P, _I, _I, _D, _P, _SZ, C.POINTER(C.c_int32), _P, _P, _P, _P]),
    "msg_bilateral_filter": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I, _D, _D]),
    "msg_segment_params_default": (None, [C.POINTER(SegmentParams)]),
    "msg_segment": (_I, [_P, _P, _SZ, _I, _I, C.POINTER(SegmentParams), _P, _SZ, _P, _SZ, _P, _SZ, C.POINTER(C.c_int32)]),
    "msg_submit_segment": (_I, [_P, _P, _SZ, _I, _I, C.POINTER(SegmentParams), _P, _SZ, _P, _SZ, _P, _SZ, C.POINTER(_I)]),
    "msg_wait": (_I, [_P, _I, C.POINTER(C.c_int32)]),
    "msg_alloc_pinned": (_P, [_SZ]),
    "msg_free_pinned": (None, [_P]),
    "msg_segment_dev": (_I, [_P, _P, _SZ, _I, _I, C.POINTER(SegmentParams), _P, _SZ, _P, _SZ, _P, _SZ, _P]),
    "msg_meanshift_filter_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _D, _D, _I, _I, _I, _D]),
    "msg_label_regions_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I, _P]),
    "msg_connected_components_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _P]),
    "msg_merge_regions_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I, _P]),
    "msg_render_labels_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _P]),
    "msg_synth_bgr_dev": (_I, [_P, _P, _SZ, _I, _I, C.c_uint64]),
    "msg_synth_bgr_rows_dev": (_I, [_P, _P, _SZ, _I, _I, _I, _I, C.c_uint64]),
    "msg_meanshift_halo_rows": (_I, [_D, _I, _I, _I]),
    "msg_meanshift_filter_strip_dev": (_I, [_P, _P, _SZ, _I, _I, _P, _SZ, _I, _I, _I, _I, _D, _D, _I, _I, _I, _D]),
    "msg_label_strip_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _I, _I, _I]),
    "msg_seam_pairs_dev": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "msg_apply_label_map_dev": (_I, [_P, _P, _SZ, _I, _I, _P, _P, _I]),
    "msg_set_profiling": (_I, [_P, _I]),
    "msg_get_kernel_profile": (_I, [_P, C.POINTER(KernelProfile)]),
    "msg_strip_rank_dev": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _P]),
    "msg_strip_query_dense_dev": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "msg_strip_apply_dense_dev": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _I, _P, _P, _I]),
    "msg_seam_quads_dev": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "msg_strip_finalize_dense_dev": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _I, _P, _P, _I, _I]),
    "msg_shard_plan_make": (_I, [_I, _I, _I, _D, _I, _I, _I, C.POINTER(ShardPlan)]),
    "msg_strip_resolve_dense_dev": (_I, [_P, _P, _I, _I, C.POINTER(_I), _P, _SZ]),
    "msg_strip_finalize_tables_dev": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _I, _I, _P]),
    "msg_strip_merge_stats_dev": (_I, [_P, _P, _SZ, _P, _SZ, _I, _I, _P, _I, _P, _P, _P, C.c_longlong, _P]),
    "msg_strip_merge_finish_dev": (_I, [_P, _P, _SZ, _I, _I, C.c_longlong, _I, _P, _P, _P, C.c_longlong, _I, _I, _P]),
    "msg_get_timings": (_I, [_P, C.POINTER(Timings)]),
    "msg_get_stats": (_I, [_P, C.POINTER(Stats)]),
    "msg_debug_get_plane": (_I, [_P, _I, _I, _P, _SZ, C.POINTER(_I), C.POINTER(_I)]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
                              "There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
