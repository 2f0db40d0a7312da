"""ctypes binding of the CPU parity oracle (oracle/libmld_oracle.so). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from mono_lidar_depth_b200._capi import MldParams

ROOT = Path(__file__).resolve().parent.parent
_lib = None


class OrcPlane(C.Structure):
    _fields_ = [("coeffs", C.c_float * 4), ("inlier_idx", C.POINTER(C.c_int32)), ("n_inliers", C.c_int64)]


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(ROOT / "oracle" / "libmld_oracle.so"))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(MldParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_default_params.argtypes = [C.POINTER(MldParams)]
        L.orc_yaml_params.argtypes = [C.POINTER(MldParams)]
        L.orc_initialize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_set_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_calculate_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(OrcPlane)]
        L.orc_visible_count.restype = C.c_int64
        L.orc_visible_count.argtypes = [C.c_void_p]
        for n in ("orc_get_point_index", "orc_get_image_points_visible", "orc_get_points_camera", "orc_get_pixel_map_visible",
                  "orc_get_pixel_map_raw"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_neighbors.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int]
        L.orc_histogram_filter.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_neighbor_finder.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                          C.c_double, C.c_void_p, C.c_int]
        L.orc_viewing_ray.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_image_point.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_ransac_plane.argtypes = [C.POINTER(MldParams), C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        _lib = L
    return _lib


def default_params() -> MldParams:
    p = MldParams()
    lib().orc_default_params(C.byref(p))
    return p


def yaml_params() -> MldParams:
    p = MldParams()
    lib().orc_yaml_params(C.byref(p))
    return p


class Oracle:
    """The reference's DepthEstimator call sequence on the CPU restatement."""

    def __init__(self, params: MldParams):
        self.L = lib()
        self.h = self.L.orc_create(C.byref(params))
        self.W = self.H = 0
        self.n = 0

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def initialize(self, W, H, f, cx, cy, T):
        T = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4])
        rc = self.L.orc_initialize(self.h, W, H, f, cx, cy, T.ctypes.data)
        assert rc == 0, rc
        self.W, self.H = W, H

    def set_cloud(self, cloud):
        a = np.ascontiguousarray(cloud, np.float32)
        self._cloud = a
        self.n = a.shape[0]
        rc = self.L.orc_set_cloud(self.h, a.ctypes.data, a.shape[0], a.shape[1])
        assert rc == 0, rc

    def calculate_depth(self, uv, plane=None):
        f = np.ascontiguousarray(uv, np.float64)
        F = f.shape[0]
        d = np.empty(F, np.float64)
        s = np.empty(F, np.int32)
        pl = None
        if plane is not None:
            coeffs, inl = plane
            inl = np.ascontiguousarray(inl, np.int32)
            pl = OrcPlane()
            for i in range(4):
                pl.coeffs[i] = float(coeffs[i])
            pl.inlier_idx = inl.ctypes.data_as(C.POINTER(C.c_int32))
            pl.n_inliers = len(inl)
        rc = self.L.orc_calculate_depth(self.h, f.ctypes.data, F, d.ctypes.data, s.ctypes.data, C.byref(pl) if pl is not None else None)
        if rc != 0:
            raise RuntimeError(f"oracle rc {rc}")
        return d, s

    def point_index(self):
        out = np.empty(self.L.orc_visible_count(self.h), np.int32)
        self.L.orc_get_point_index(self.h, out.ctypes.data)
        return out

    def image_points_visible(self):
        out = np.empty((self.L.orc_visible_count(self.h), 2), np.float64)
        self.L.orc_get_image_points_visible(self.h, out.ctypes.data)
        return out

    def points_camera(self):
        out = np.empty((self.n, 3), np.float64)
        self.L.orc_get_points_camera(self.h, out.ctypes.data)
        return out

    def pixel_map_visible(self):
        out = np.empty((self.H, self.W), np.int32)
        self.L.orc_get_pixel_map_visible(self.h, out.ctypes.data)
        return out

    def pixel_map_raw(self):
        out = np.empty((self.H, self.W), np.int32)
        self.L.orc_get_pixel_map_raw(self.h, out.ctypes.data)
        return out

    def neighbors(self, u, v, sw=1.0, sh=1.0):
        out = np.empty(4096, np.int32)
        k = self.L.orc_get_neighbors(self.h, u, v, sw, sh, out.ctypes.data, 4096)
        return out[:k].copy()


def histogram_filter(depths, bin_width, min_count):
    d = np.ascontiguousarray(depths, np.float64)
    pos = np.empty(len(d), np.int32)
    n = C.c_int(0)
    lo, hi = C.c_double(0), C.c_double(0)
    ok = lib().orc_histogram_filter(d.ctypes.data, len(d), bin_width, min_count, pos.ctypes.data, C.byref(n), C.byref(lo), C.byref(hi))
    return bool(ok), pos[: n.value].copy(), lo.value, hi.value


def ransac_plane(params: MldParams, cloud, seed):
    a = np.ascontiguousarray(cloud, np.float32)
    coeffs = np.zeros(4, np.float32)
    idx = np.empty(max(a.shape[0], 1), np.int32)
    n = C.c_int64(0)
    it = C.c_int32(0)
    rc = lib().orc_ransac_plane(C.byref(params), a.ctypes.data, a.shape[0], a.shape[1], seed, coeffs.ctypes.data, idx.ctypes.data,
                                C.byref(n), C.byref(it))
    return rc, coeffs, idx[: n.value].copy(), it.value
