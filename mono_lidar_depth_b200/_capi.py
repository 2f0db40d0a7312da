"""ctypes binding of libmld_cuda.so (include/mld_c_api.h).

The library is the product: if it is missing, or if no CUDA device is usable, every entry point
raises -- there is no CPU fallback and nothing here ever touches the CPU parity checker.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmld_cuda.so"


class MldError(RuntimeError):
    """A negative mld_error code from the C ABI; `.code` holds it."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[mld {code}] {message}")
        self.code = code
        self.message = message


# mld_error (include/mld_c_api.h)
MLD_OK = 0
MLD_ERR_INVALID_ARG = -1
MLD_ERR_NOT_CONFIGURED = -2
MLD_ERR_NOT_INITIALIZED = -3
MLD_ERR_NO_CLOUD = -4
MLD_ERR_BAD_SEARCH_MODE = -5
MLD_ERR_REGION_GROWING = -6
MLD_ERR_PCL_INVALID = -7
MLD_ERR_NO_ROAD_ESTIMATOR = -8
MLD_ERR_CAPACITY = -9
MLD_ERR_CUDA = -10
MLD_ERR_NO_MODEL = -11
MLD_ERR_IO = -12


class MldParams(C.Structure):
    """mld_params == Mono_Lidar::DepthEstimatorParameters (DepthEstimatorParameters.h:12-172)."""

    _fields_ = [
        ("neighbor_search_mode", C.c_int32),
        ("pixelarea_search_witdh", C.c_int32),
        ("pixelarea_search_height", C.c_int32),
        ("radiusSearch_count_min", C.c_int32),
        ("do_use_histogram_segmentation", C.c_int32),
        ("histogram_segmentation_min_pointcount", C.c_int32),
        ("histogram_segmentation_bin_witdh", C.c_double),
        ("do_use_depth_segmentation", C.c_int32),
        ("treshold_depth_enabled", C.c_int32),
        ("treshold_depth_mode", C.c_int32),
        ("treshold_depth_max", C.c_int32),
        ("treshold_depth_min", C.c_int32),
        ("treshold_depth_local_enabled", C.c_int32),
        ("treshold_depth_local_mode", C.c_int32),
        ("treshold_depth_local_valuetype", C.c_int32),
        ("treshold_depth_local_value", C.c_double),
        ("do_use_PCA", C.c_int32),
        ("pca_debug", C.c_int32),
        ("pca_treshold_3_abs_min", C.c_double),
        ("pca_treshold_3_2_rel_max", C.c_double),
        ("pca_treshold_2_1_rel_min", C.c_double),
        ("do_use_ransac_plane", C.c_int32),
        ("ransac_plane_max_iterations", C.c_int32),
        ("ransac_plane_distance_treshold", C.c_double),
        ("ransac_plane_min_z", C.c_double),
        ("ransac_plane_max_z", C.c_double),
        ("ransac_plane_use_refinement", C.c_int32),
        ("ransac_plane_use_camx_treshold", C.c_int32),
        ("ransac_plane_refinement_treshold", C.c_double),
        ("ransac_plane_treshold_camx", C.c_double),
        ("ransac_plane_point_distance_treshold", C.c_double),
        ("ransac_plane_probability", C.c_double),
        ("plane_estimator_use_triangle_maximation", C.c_int32),
        ("plane_estimator_use_leastsquares", C.c_int32),
        ("plane_estimator_use_mestimator", C.c_int32),
        ("do_use_cut_behind_camera", C.c_int32),
        ("plane_estimator_z_x_min_relation", C.c_double),
        ("do_use_triangle_size_maximation", C.c_int32),
        ("do_check_triangleplanar_condition", C.c_int32),
        ("triangleplanar_crossnorm_treshold", C.c_double),
        ("viewray_plane_orthoganality_treshold", C.c_double),
        ("set_all_depths_to_zero", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class MldPlane(C.Structure):
    _fields_ = [
        ("coeffs", C.c_float * 4),
        ("inlier_idx", C.POINTER(C.c_int32)),
        ("n_inliers", C.c_int64),
        ("inlier_capacity", C.c_int64),
        ("segmented", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class MldSynthConfig(C.Structure):
    """mld_synth_config (include/mld_synth.h)."""

    _fields_ = [
        ("rings", C.c_int32),
        ("azimuth_steps", C.c_int32),
        ("elev_top_deg", C.c_float),
        ("elev_bottom_deg", C.c_float),
        ("sensor_height", C.c_float),
        ("max_range", C.c_float),
        ("range_noise_sigma", C.c_float),
        ("dropout_prob", C.c_float),
        ("n_boxes", C.c_int32),
        ("image_width", C.c_int32),
        ("image_height", C.c_int32),
        ("band_top_frac", C.c_float),
        ("above_band_frac", C.c_float),
        ("object_frac", C.c_float),
        ("road_frac", C.c_float),
        ("cam_f", C.c_float),
        ("cam_cx", C.c_float),
        ("cam_cy", C.c_float),
        ("cam_T", C.c_float * 12),
        ("two_block_rings", C.c_int32),
    ]


# every symbol include/mld_c_api.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_PP = C.POINTER(MldParams)
_PL = C.POINTER(MldPlane)
_SC = C.POINTER(MldSynthConfig)
SYMBOLS = {
    "mld_sizeof_params": (C.c_int, []),
    "mld_default_params": (None, [_PP]),
    "mld_params_from_yaml": (C.c_int, [C.c_char_p, _PP]),
    "mld_status_name": (C.c_char_p, [C.c_int]),
    "mld_last_error": (C.c_char_p, [_H]),
    "mld_create": (C.c_int, [_PP, C.c_int, C.POINTER(_H)]),
    "mld_destroy": (C.c_int, [_H]),
    "mld_initialize": (C.c_int, [_H, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "mld_set_cloud": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int, _PL, C.c_uint64]),
    "mld_calculate_depth": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _PL]),
    "mld_estimate_ground_plane": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_int, C.c_uint64, _PL, C.POINTER(C.c_int32)]),
    "mld_semantic_ground_plane": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
         C.c_int, C.c_double, _PL],
    ),
    "mld_semantic_ground_labelled": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
         C.c_int, C.c_void_p],
    ),
    "mld_semantic_ground_plane_device": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p,
         C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mld_get_visible_points": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "mld_process_frames_device": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
         C.c_uint64, C.c_void_p, C.c_void_p],
    ),
    "mld_process_frames_device_semantic": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p,
         C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mld_process_frames_device_planes": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
         C.c_void_p],
    ),
    "mld_process_frames_host": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
         C.c_uint64, C.c_void_p],
    ),
    "mld_calculate_depth_pair": (
        C.c_int,
        [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _PL, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
         C.c_void_p, C.c_void_p, _PL, C.c_int, C.c_uint64],
    ),
    "mld_set_statistics": (C.c_int, [_H, C.c_int]),
    "mld_set_semantic_exact": (C.c_int, [_H, C.c_int]),
    "mld_last_status_histogram": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "mld_get_points_camera_indexed": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p]),
    "mld_get_triangle_corners": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mld_yaml_int": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mld_calculate_depth_pair_resident": (
        C.c_int,
        [_H, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _PL, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _PL,
         C.c_int, C.c_uint64],
    ),
    "mld_has_resident_cloud": (C.c_int, [_H]),
    "mld_status_histogram_host": (C.c_int, [_H, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "mld_status_histogram_device": (C.c_int, [_H, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "mld_pack_feature_points_device": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mld_kernel_launch_count": (C.c_int64, [_H]),
    "mld_neighbor_capacity": (C.c_int, []),
    "mld_host_pipeline_stats": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "mld_profile_enable": (C.c_int, [_H, C.c_int]),
    "mld_profile_read": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mld_fused_chunk_frames": (C.c_int, [_H]),
    "mld_chunk_frames": (C.c_int, [_H]),
    "mld_get_pixel_map": (C.c_int, [_H, C.c_void_p]),
    "mld_get_neighbors": (C.c_int, [_H, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "mld_get_visible": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_int64)]),
    "mld_get_points_camera": (C.c_int, [_H, C.c_void_p]),
    "mld_synth_points_device": (C.c_int, [_H, _SC, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "mld_synth_features_device": (C.c_int, [_H, _SC, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
}

# libmld_synth.so (include/mld_synth.h): host generators of the synthetic input, plain C++ without CUDA
SYNTH_LIB_PATH = _PKG / "libmld_synth.so"
SYNTH_SYMBOLS = {
    "mld_synth_default_config": (None, [_SC, C.c_int]),
    "mld_synth_config_for": (None, [_SC, C.c_int, C.c_int]),
    "mld_synth_points_per_frame": (C.c_int64, [_SC]),
    "mld_synth_points_host": (C.c_int, [_SC, C.c_uint64, C.c_int64, C.c_void_p]),
    "mld_synth_points_host_xyzi32": (C.c_int, [_SC, C.c_uint64, C.c_int64, C.c_void_p]),
    "mld_synth_features_host": (C.c_int, [_SC, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
}

_lib = None
_synth_lib = None


def load_synth() -> C.CDLL:
    """Load libmld_synth.so (no CUDA involved: usable by the CPU arm of bench.py and the CPU tests)."""
    global _synth_lib
    if _synth_lib is not None:
        return _synth_lib
    path = Path(os.environ.get("MLD_SYNTH_LIB", str(SYNTH_LIB_PATH)))
    if not path.exists():
        raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(str(path))
    for name, (res, args) in SYNTH_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _synth_lib = lib
    return lib


def load() -> C.CDLL:
    """Load libmld_cuda.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("MLD_CUDA_LIB", str(LIB_PATH)))
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the depth-estimation path)"
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mld_sizeof_params() != C.sizeof(MldParams):
        raise ImportError("mld_params layout mismatch between include/mld_c_api.h and _capi.py")
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    if rc == MLD_OK:
        return
    lib = load()
    msg = lib.mld_last_error(handle)
    raise MldError(rc, msg.decode("utf-8", "replace") if msg else "")
