"""Synthetic KITTI-shaped input (include/mld_synth.h): host generators of libmld_synth.so (plain C++, no
CUDA) and device generators of libmld_cuda.so, which agree bit for bit.

Not a reference component -- the reference ships no data. Shapes: K = HDL-64-like sweep of
64 x 1875 = 120 000 points, 1241 x 376 image, 2000 features; D = 128 x 2032 = 260 096 points,
2048 x 1024 image, 20 000 features.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .depth_estimator import CameraPinhole

# KITTI raw 2011_09_26 calib_velo_to_cam (public calibration values), row-major 3x4 [R|t]
KITTI_T_LIDAR_TO_CAM = np.array(
    [
        [7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
        [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
        [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01],
    ],
    dtype=np.float64,
)


def kitti_camera() -> CameraPinhole:
    return CameraPinhole(1241, 376, 718.856, 607.1928, 185.2157)


def dense_camera() -> CameraPinhole:
    return CameraPinhole(2048, 1024, 1400.0, 1024.0, 420.0)


def default_config(dense: bool = False, road: bool = False) -> _capi.MldSynthConfig:
    """dense: the 128-beam shape; road: the road / non-road feature mix of BASELINE.json configs[2] (half of the features
    in the lower third of the image, on ground returns)."""
    c = _capi.MldSynthConfig()
    _capi.load_synth().mld_synth_config_for(C.byref(c), 1 if dense else 0, 1 if road else 0)
    return c


def points_per_frame(cfg) -> int:
    return int(cfg.rings) * int(cfg.azimuth_steps)


def points_host(cfg, seed: int, frame: int) -> np.ndarray:
    """(n, 4) float32 x,y,z,intensity in the lidar frame; dropouts / no-returns are NaN."""
    out = np.empty((points_per_frame(cfg), 4), np.float32)
    if _capi.load_synth().mld_synth_points_host(C.byref(cfg), seed, frame, out.ctypes.data) != 0:
        raise ValueError("mld_synth_points_host: bad configuration")
    return out


def points_host_xyzi32(cfg, seed: int, frame: int) -> np.ndarray:
    """The same cloud as (n, 8) float32 records = 32-byte pcl::PointXYZI (x, y, z, 1, intensity, padding)."""
    out = np.empty((points_per_frame(cfg), 8), np.float32)
    if _capi.load_synth().mld_synth_points_host_xyzi32(C.byref(cfg), seed, frame, out.ctypes.data) != 0:
        raise ValueError("mld_synth_points_host_xyzi32: bad configuration")
    return out


def features_host(cfg, seed: int, frame: int, F: int) -> np.ndarray:
    """(F, 2) float64 integer pixel coordinates (memory layout of Eigen::Matrix2Xd)."""
    out = np.empty((F, 2), np.float64)
    if _capi.load_synth().mld_synth_features_host(C.byref(cfg), seed, frame, F, out.ctypes.data) != 0:
        raise ValueError("mld_synth_features_host: bad configuration")
    return out


def points_device(est, cfg, seed: int, frame0: int, nframes: int, d_out: int, frame_pitch_points: int = 0, stream: int = 0):
    pitch = frame_pitch_points or points_per_frame(cfg)
    _capi.check(_capi.load().mld_synth_points_device(est.handle, C.byref(cfg), seed, frame0, nframes, pitch, d_out, stream or None),
                est.handle)


def features_device(est, cfg, seed: int, frame0: int, nframes: int, F: int, d_out: int, stream: int = 0):
    _capi.check(_capi.load().mld_synth_features_device(est.handle, C.byref(cfg), seed, frame0, nframes, F, d_out, stream or None),
                est.handle)
