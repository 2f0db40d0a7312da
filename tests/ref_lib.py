"""ctypes binding of oracle/_ref/libmld_ref.so: the REFERENCE's own monolidar_fusion sources compiled against the
stand-in Eigen/PCL/OpenCV headers (oracle/ref_standin, oracle/ref_bridge.cpp). Test infrastructure only; exists only
where /root/reference was present at build time (this container) or where the prebuilt .so travelled to."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from mono_lidar_depth_b200._capi import MldParams
from oracle_lib import OrcPlane

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "oracle" / "_ref" / "libmld_ref.so"
REF_YAML = Path("/root/reference/monolidar_fusion/parameters.yaml")
_lib = None


def available() -> bool:
    return SO.exists()


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(SO))
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.POINTER(MldParams)]
        L.ref_create_from_yaml.restype = C.c_void_p
        L.ref_create_from_yaml.argtypes = [C.c_char_p, C.POINTER(MldParams)]
        L.ref_default_params.argtypes = [C.POINTER(MldParams)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_initialize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.ref_set_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(OrcPlane)]
        L.ref_calculate_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_calculate_depth_single.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int]
        L.ref_get_plane.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.ref_visible_count.restype = C.c_int64
        L.ref_visible_count.argtypes = [C.c_void_p]
        for n in ("ref_get_point_index", "ref_get_image_points_visible", "ref_get_points_camera", "ref_get_pixel_map_visible",
                  "ref_get_pixel_map_raw"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_point_depth_cam_visible.restype = C.c_double
        L.ref_get_point_depth_cam_visible.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_neighbors.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int]
        L.ref_histogram_filter.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_neighbor_finder.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                          C.c_double, C.c_void_p, C.c_int]
        L.ref_viewing_ray.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.ref_image_point.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.ref_set_seed.argtypes = [C.c_uint]
        L.ref_ransac_plane.argtypes = [C.POINTER(MldParams), C.c_void_p, C.c_int64, C.c_int, C.c_uint, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int64)]
        L.ref_semantic_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_double, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _plane_struct(plane):
    coeffs, inl = plane
    inl = np.ascontiguousarray(inl, np.int32)
    pl = OrcPlane()
    for i in range(4):
        pl.coeffs[i] = float(coeffs[i])
    pl.inlier_idx = inl.ctypes.data_as(C.POINTER(C.c_int32))
    pl.n_inliers = len(inl)
    return pl, inl


def default_params() -> MldParams:
    p = MldParams()
    lib().ref_default_params(C.byref(p))
    return p


def yaml_params(path=REF_YAML) -> MldParams:
    """parameters.yaml as the reference's own loader reads it (absent keys -> 0)."""
    p = MldParams()
    h = lib().ref_create_from_yaml(str(path).encode(), C.byref(p))
    assert h, lib().ref_last_error()
    lib().ref_destroy(h)
    return p


class Reference:
    """Same call sequence as oracle_lib.Oracle, executed by the reference's own DepthEstimator."""

    def __init__(self, params: MldParams):
        self.L = lib()
        self.h = self.L.ref_create(C.byref(params))
        self.W = self.H = 0
        self.n = 0
        self.has_plane = False

    def __del__(self):
        try:
            self.L.ref_destroy(self.h)
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"reference threw (rc {rc}): {self.L.ref_last_error().decode(errors='replace')}")

    def initialize(self, W, H, f, cx, cy, T):
        T = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4])
        self._check(self.L.ref_initialize(self.h, W, H, f, cx, cy, T.ctypes.data))
        self.W, self.H = W, H

    def set_cloud(self, cloud, plane=None):
        a = np.ascontiguousarray(cloud, np.float32)
        self.n = a.shape[0]
        pl = None
        if plane is not None:
            pl, self._inl = _plane_struct(plane)
        self.has_plane = plane is not None
        self._check(self.L.ref_set_cloud(self.h, a.ctypes.data, a.shape[0], a.shape[1], C.byref(pl) if pl is not None else None))

    def calculate_depth(self, uv, with_plane=None):
        f = np.ascontiguousarray(uv, np.float64)
        F = f.shape[0]
        d = np.empty(F, np.float64)
        s = np.empty(F, np.int32)
        wp = self.has_plane if with_plane is None else with_plane
        self._check(self.L.ref_calculate_depth(self.h, f.ctypes.data, F, d.ctypes.data, s.ctypes.data, int(wp)))
        return d, s

    def calculate_depth_single(self, u, v, with_plane=None):
        """The single-feature overload (DepthEstimator.cpp:491-600). Throw sites are only observable here: inside the
        batch overload's OpenMP region an exception terminates the process (in the real reference too)."""
        d, s = C.c_double(0), C.c_int32(0)
        wp = self.has_plane if with_plane is None else with_plane
        self._check(self.L.ref_calculate_depth_single(self.h, float(u), float(v), C.byref(d), C.byref(s), int(wp)))
        return d.value, s.value

    def plane(self):
        coeffs = np.zeros(4, np.float32)
        idx = np.empty(max(self.n, 1), np.int32)
        n = C.c_int64(0)
        rc = self.L.ref_get_plane(self.h, coeffs.ctypes.data, idx.ctypes.data, C.byref(n))
        assert rc == 0
        return coeffs, idx[: n.value].copy()

    def point_index(self):
        out = np.empty(self.L.ref_visible_count(self.h), np.int32)
        self.L.ref_get_point_index(self.h, out.ctypes.data)
        return out

    def image_points_visible(self):
        out = np.empty((self.L.ref_visible_count(self.h), 2), np.float64)
        self.L.ref_get_image_points_visible(self.h, out.ctypes.data)
        return out

    def points_camera(self):
        out = np.empty((self.n, 3), np.float64)
        self.L.ref_get_points_camera(self.h, out.ctypes.data)
        return out

    def point_depth_cam_visible(self, i):
        return self.L.ref_get_point_depth_cam_visible(self.h, int(i))

    def pixel_map_visible(self):
        out = np.empty((self.H, self.W), np.int32)
        self.L.ref_get_pixel_map_visible(self.h, out.ctypes.data)
        return out

    def pixel_map_raw(self):
        out = np.empty((self.H, self.W), np.int32)
        self.L.ref_get_pixel_map_raw(self.h, out.ctypes.data)
        return out

    def neighbors(self, u, v, sw=1.0, sh=1.0):
        out = np.empty(4096, np.int32)
        k = self.L.ref_get_neighbors(self.h, u, v, sw, sh, out.ctypes.data, 4096)
        return out[:k].copy()


def histogram_filter(depths, bin_width, min_count):
    d = np.ascontiguousarray(depths, np.float64)
    pos = np.empty(len(d), np.int32)
    n = C.c_int(0)
    lo, hi = C.c_double(0), C.c_double(0)
    ok = lib().ref_histogram_filter(d.ctypes.data, len(d), bin_width, min_count, pos.ctypes.data, C.byref(n), C.byref(lo), C.byref(hi))
    return bool(ok), pos[: n.value].copy(), lo.value, hi.value


def neighbor_finder(W, H, sw, sh, img, cam, u, v):
    img = np.ascontiguousarray(img, np.float64)
    cam = np.ascontiguousarray(cam, np.float64)
    out = np.empty(4096, np.int32)
    k = lib().ref_neighbor_finder(W, H, sw, sh, img.ctypes.data, cam.ctypes.data, len(img), u, v, out.ctypes.data, 4096)
    return out[:k].copy()


def viewing_ray(W, H, f, cx, cy, u, v):
    out = np.empty(3, np.float64)
    lib().ref_viewing_ray(W, H, f, cx, cy, u, v, out.ctypes.data)
    return out


def image_point(W, H, f, cx, cy, p3):
    p = np.ascontiguousarray(p3, np.float64)
    out = np.empty(2, np.float64)
    ok = lib().ref_image_point(W, H, f, cx, cy, p.ctypes.data, out.ctypes.data)
    return bool(ok), out


def ransac_plane(params: MldParams, cloud, seed):
    a = np.ascontiguousarray(cloud, np.float32)
    coeffs = np.zeros(4, np.float32)
    idx = np.empty(max(a.shape[0], 1), np.int32)
    n = C.c_int64(0)
    rc = lib().ref_ransac_plane(C.byref(params), a.ctypes.data, a.shape[0], a.shape[1], seed, coeffs.ctypes.data, idx.ctypes.data, C.byref(n))
    return rc, coeffs, idx[: n.value].copy()


def semantic_plane(labels, f, cu, cv, T, ground_labels, inlier_threshold, cloud):
    lab = np.ascontiguousarray(labels, np.uint8)
    H, W = lab.shape
    T = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4])
    gl = np.ascontiguousarray(ground_labels, np.int32)
    a = np.ascontiguousarray(cloud, np.float32)
    coeffs = np.zeros(4, np.float32)
    idx = np.empty(max(a.shape[0], 1), np.int32)
    n = C.c_int64(0)
    rc = lib().ref_semantic_plane(lab.ctypes.data, W, H, f, cu, cv, T.ctypes.data, gl.ctypes.data, len(gl), inlier_threshold,
                                  a.ctypes.data, a.shape[0], a.shape[1], coeffs.ctypes.data, idx.ctypes.data, C.byref(n))
    return rc, coeffs, idx[: n.value].copy()
