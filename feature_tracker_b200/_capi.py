"""ctypes binding of libftk_b200.so (the C ABI in include/ftk_c.h).

The library is the product: hand-written sm_100a kernels behind a C ABI.  There is no CPU fallback -- importing this
module without the built library raises, and creating a context without a B200 raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FTK_LIB_PATH: A/B measurements of alternative builds of the same library (never a different implementation)
LIB_PATH = os.environ.get("FTK_LIB_PATH") or os.path.join(_HERE, "libftk_b200.so")

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_EMPTY_INPUT = -2
ERR_LEVEL_MISMATCH = -3
ERR_SIZE_MISMATCH = -4
ERR_CUDA = -5
ERR_UNSUPPORTED = -6

FLAG_DEVICE_POINTERS = 1
FLAG_NO_PREDICTION = 2
FLAG_NO_STATUS = 4
FLAG_SINGLE_LEVEL = 8
FLAG_NO_INDEX_INPUT = 16

# Every symbol include/ftk_c.h declares (tests check that the library exports all of them).
EXPORTED_SYMBOLS = [
    "ftk_abi_version", "ftk_klt_params_default", "ftk_create", "ftk_destroy", "ftk_last_error", "ftk_synchronize", "ftk_stream",
    "ftk_kernel_launches", "ftk_alloc_pinned", "ftk_free_pinned", "ftk_set_profiling", "ftk_last_kernel_ms", "ftk_pyramid_create", "ftk_pyramid_destroy", "ftk_pyramid_set_images", "ftk_pyramid_build",
    "ftk_pyramid_set_level", "ftk_pyramid_get_level", "ftk_pyramid_levels", "ftk_pyramid_images", "ftk_klt_track",
    "ftk_track_image_pairs", "ftk_track_image_pairs_multi", "ftk_track_image_sequence",
    "ftk_match_hamming_force", "ftk_match_hamming_nearby", "ftk_match_hamming_pairs", "ftk_match_cosine_pairs", "ftk_match_cosine_force", "ftk_match_cosine_nearby", "ftk_fill_matched", "ftk_last_cosine_exact_scan_items",
    "ftk_match_mutual_scores", "ftk_match_cross_check", "ftk_direct_params_default", "ftk_direct_method_track", "ftk_dense_flow_params_default", "ftk_dense_flow_track",
    "ftk_detector_params_default", "ftk_detect_features", "ftk_detect_features_batch", "ftk_detect_response", "ftk_brief_pattern_default", "ftk_describe_brief", "ftk_describe_brief_batch",
]


class KltParams(C.Structure):
    """ftk_klt_params (include/ftk_c.h) == OpticalFlowOptions + subclass extras of the reference."""

    _fields_ = [
        ("variant", C.c_int32),
        ("method", C.c_int32),
        ("max_track_points", C.c_uint32),
        ("max_iteration", C.c_uint32),
        ("max_tolerance_large_step", C.c_uint32),
        ("patch_row_half", C.c_int32),
        ("patch_col_half", C.c_int32),
        ("max_converge_step", C.c_float),
        ("predict", C.c_float * 4),
        ("consider_patch_luminance", C.c_int32),
        ("forward_backward_max_error", C.c_float),
    ]


class DirectParams(C.Structure):
    """ftk_direct_params (include/ftk_c.h) == DirectMethodOptions of the reference."""

    _fields_ = [
        ("max_track_points", C.c_uint32),
        ("max_iteration", C.c_uint32),
        ("patch_row_half", C.c_int32),
        ("patch_col_half", C.c_int32),
        ("max_converge_step", C.c_float),
        ("max_converge_residual", C.c_float),
        ("method", C.c_int32),
    ]


class DenseFlowParams(C.Structure):
    """ftk_dense_flow_params (include/ftk_c.h) == DenseOpticalFlow::Options of the reference."""

    _fields_ = [("max_iteration", C.c_int32), ("half_patch_size", C.c_int32), ("max_converge_step", C.c_float), ("max_delta_flow_step", C.c_float)]


class DetectorParams(C.Structure):
    """ftk_detector_params (include/ftk_c.h): FeaturePointDetector options by the names the reference's call sites use."""

    _fields_ = [("kind", C.c_int32), ("half_patch", C.c_int32), ("harris_k", C.c_float), ("min_response", C.c_float), ("min_distance", C.c_int32)]


def load_library():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); feature_tracker_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_float
    P = C.POINTER
    sig = {
        "ftk_abi_version": (C.c_int, []),
        "ftk_klt_params_default": (None, [P(KltParams)]),
        "ftk_create": (C.c_int, [C.c_int, P(vp)]),
        "ftk_destroy": (None, [vp]),
        "ftk_last_error": (C.c_char_p, [vp]),
        "ftk_synchronize": (C.c_int, [vp]),
        "ftk_stream": (vp, [vp]),
        "ftk_kernel_launches": (C.c_uint64, [vp]),
        "ftk_alloc_pinned": (C.c_int, [C.c_size_t, P(vp)]),
        "ftk_free_pinned": (None, [vp]),
        "ftk_set_profiling": (C.c_int, [vp, C.c_int]),
        "ftk_last_kernel_ms": (C.c_float, [vp]),
        "ftk_pyramid_create": (C.c_int, [vp, i32, i32, i32, i32, P(vp)]),
        "ftk_pyramid_destroy": (None, [vp, vp]),
        "ftk_pyramid_set_images": (C.c_int, [vp, vp, i32, i32, vp, u32]),
        "ftk_pyramid_build": (C.c_int, [vp, vp, i32, i32]),
        "ftk_pyramid_set_level": (C.c_int, [vp, vp, i32, i32, vp]),
        "ftk_pyramid_get_level": (C.c_int, [vp, vp, i32, i32, vp]),
        "ftk_pyramid_levels": (i32, [vp]),
        "ftk_pyramid_images": (i32, [vp]),
        "ftk_klt_track": (C.c_int, [vp, P(KltParams), vp, vp, i32, vp, vp, vp, vp, vp, vp, u32]),
        "ftk_track_image_pairs": (C.c_int, [vp, P(KltParams), i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, u32]),
        "ftk_track_image_pairs_multi": (C.c_int, [vp, P(KltParams), i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, u32]),
        "ftk_track_image_sequence": (C.c_int, [vp, P(KltParams), i32, i32, i32, i32, vp, vp, vp, vp, vp, u32]),
        "ftk_match_hamming_force": (C.c_int, [vp, vp, i32, vp, i32, i32, f32, vp, u32]),
        "ftk_match_hamming_nearby": (C.c_int, [vp, vp, i32, vp, i32, i32, vp, vp, i32, i32, f32, vp, u32]),
        "ftk_match_hamming_pairs": (C.c_int, [vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, f32, vp, u32]),
        "ftk_match_cosine_pairs": (C.c_int, [vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, f32, vp, u32]),
        "ftk_match_cosine_force": (C.c_int, [vp, vp, i32, vp, i32, i32, f32, vp, u32]),
        "ftk_match_cosine_nearby": (C.c_int, [vp, vp, i32, vp, i32, i32, vp, vp, i32, i32, f32, vp, u32]),
        "ftk_dense_flow_params_default": (None, [P(DenseFlowParams)]),
        "ftk_dense_flow_track": (C.c_int, [vp, P(DenseFlowParams), vp, vp, i32, i32, vp, vp, u32]),
        "ftk_detector_params_default": (None, [P(DetectorParams)]),
        "ftk_detect_features": (C.c_int, [vp, P(DetectorParams), vp, i32, vp, i32, i32, vp, vp, P(i32), u32]),
        "ftk_detect_features_batch": (C.c_int, [vp, P(DetectorParams), vp, i32, i32, i32, vp, vp, vp, u32]),
        "ftk_detect_response": (C.c_int, [vp, P(DetectorParams), vp, i32, vp, u32]),
        "ftk_brief_pattern_default": (None, [i32, i32, u32, vp]),
        "ftk_describe_brief": (C.c_int, [vp, vp, i32, vp, i32, vp, i32, i32, vp, vp, u32]),
        "ftk_describe_brief_batch": (C.c_int, [vp, vp, i32, i32, vp, vp, vp, i32, i32, vp, vp, u32]),
        "ftk_direct_params_default": (None, [P(DirectParams)]),
        "ftk_direct_method_track": (C.c_int, [vp, P(DirectParams), vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32]),
        "ftk_match_mutual_scores": (C.c_int, [vp, vp, i32, i32, f32, vp, u32]),
        "ftk_match_cross_check": (C.c_int, [vp, vp, i32, vp, i32, u32]),
        "ftk_fill_matched": (C.c_int, [vp, i32, vp, i32, vp, vp, i32]),
        "ftk_last_cosine_exact_scan_items": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
