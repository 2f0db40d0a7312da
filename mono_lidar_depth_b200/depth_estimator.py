"""Host-side mirror of the reference's operator interface, Mono_Lidar::DepthEstimator
(monolidar_fusion/include/monolidar_fusion/DepthEstimator.h:39-359), on top of the C ABI.

Same method names, argument meaning and error behaviour as the C++ class:
    InitConfig -> Initialize -> setInputCloud -> CalculateDepth
All arithmetic runs in libmld_cuda.so on the GPU; this file only moves buffers and mirrors the
reference's exceptions. Python has no reference arguments, so the in/out ``GroundPlane::Ptr&`` of
setInputCloud / CalculateDepth is returned instead.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _capi
from .params import DepthEstimatorParameters


class ExceptionPclInvalid(Exception):
    """GroundPlane::ExceptionPclInvalid (RansacPlane.h:46-50)."""

    def __str__(self):
        return "In GroundPlane: Input pointcloud is invalid"


class CameraPinhole:
    """CameraPinhole(width, height, focal_length, principal_point_x, principal_point_y) (camera_pinhole.h:21-26)."""

    def __init__(self, width: int, height: int, focal_length: float, principal_point_x: float, principal_point_y: float):
        self.width_, self.height_ = int(width), int(height)
        self.focal_length_ = float(focal_length)
        self.principal_point_x_, self.principal_point_y_ = float(principal_point_x), float(principal_point_y)

    def getImageSize(self) -> Tuple[int, int]:
        return self.width_, self.height_


class GroundPlane:
    """Mono_Lidar::GroundPlane (RansacPlane.h:38-126): model coefficients in the lidar frame, inlier
    indices into the raw cloud, isSegmented(). Construct it empty (to be fitted by RANSAC on the GPU)
    or with externally computed coefficients/inliers (e.g. a SemanticPlane computed by the caller)."""

    def __init__(self, coeffs=None, inliers=None):
        self._coeffs = np.zeros(4, np.float32) if coeffs is None else np.asarray(coeffs, np.float32).reshape(4).copy()
        self._inliers = np.zeros(0, np.int32) if inliers is None else np.ascontiguousarray(inliers, np.int32)
        self.is_segmented_ = coeffs is not None

    def isSegmented(self) -> bool:
        return bool(self.is_segmented_)

    def getModelCoeffs(self) -> np.ndarray:
        return self._coeffs

    def getInlinersIndex(self) -> np.ndarray:
        return self._inliers

    def CheckPointInPlane(self, index: int) -> bool:
        i = int(np.searchsorted(self._inliers, index))
        return i < len(self._inliers) and int(self._inliers[i]) == int(index)

    def _as_c(self, capacity: int = 0) -> _capi.MldPlane:
        pl = _capi.MldPlane()
        if capacity > len(self._inliers):
            buf = np.zeros(capacity, np.int32)
            buf[: len(self._inliers)] = self._inliers
            self._inliers_buf = buf
        else:
            self._inliers_buf = self._inliers
        for i in range(4):
            pl.coeffs[i] = float(self._coeffs[i])
        pl.inlier_idx = self._inliers_buf.ctypes.data_as(C.POINTER(C.c_int32)) if len(self._inliers_buf) else None
        pl.n_inliers = len(self._inliers)
        pl.inlier_capacity = len(self._inliers_buf)
        pl.segmented = 1 if self.is_segmented_ else 0
        return pl

    def _from_c(self, pl: _capi.MldPlane) -> None:
        self._coeffs = np.array([pl.coeffs[i] for i in range(4)], np.float32)
        self._inliers = self._inliers_buf[: min(pl.n_inliers, len(self._inliers_buf))].copy()
        self.is_segmented_ = bool(pl.segmented)


class RansacPlane(GroundPlane):
    """Mono_Lidar::RansacPlane (RansacPlane.h:132-164); the fit itself runs on the GPU."""

    def __init__(self, parameters: Optional[DepthEstimatorParameters] = None, seed: int = 0):
        super().__init__()
        self.parameters = parameters
        self.seed = int(seed)
        self.iterations = 0


class SemanticPlane(GroundPlane):
    """Mono_Lidar::SemanticPlane (RansacPlane.h:166-216, RansacPlane.cpp:159-274): the ground plane fitted to the lidar
    points whose projection carries a ground label in a semantic image; the fit runs on the GPU
    (mld_semantic_ground_plane). Same constructor arguments as the reference plus the estimator whose device handle is
    used: SemanticPlane(img, cam, groundplane_label, inlier_threshold, estimator)."""

    class Camera:
        """SemanticPlane::Camera: f, cu, cv and transform_cam_lidar (3x4 or 4x4, camera <- lidar)."""

        def __init__(self, f: float, cu: float, cv: float, transform_cam_lidar):
            self.f, self.cu, self.cv = float(f), float(cu), float(cv)
            self.transform_cam_lidar = np.ascontiguousarray(np.asarray(transform_cam_lidar, np.float64)[:3, :4])

    def __init__(self, img, cam: "SemanticPlane.Camera", groundplane_label=(6, 7, 8, 9), inlier_threshold: float = 0.1, estimator=None):
        super().__init__()
        self.semantic_image_ = np.ascontiguousarray(img, np.uint8)
        if self.semantic_image_.ndim != 2:
            raise ValueError("semantic image must be a single-channel 8-bit image (H, W)")
        self.cam_ = cam
        self.groundplane_label_ = sorted(int(x) for x in set(groundplane_label))
        self.inlier_threshold_ = float(inlier_threshold)
        self._estimator = estimator

    def CalculateInliersPlane(self, cloud, min_z: float = -1000.0, max_z: float = 1000.0) -> None:
        """RansacPlane.cpp:195-274 (min_z / max_z are ignored by the reference's override too, RansacPlane.h:205-207).
        Raises ExceptionPclInvalid when fewer than 3 points carry a ground label."""
        if self._estimator is None:
            raise RuntimeError("SemanticPlane needs the DepthEstimator whose GPU handle it runs on")
        self._estimator._semantic_plane(self, cloud)


def _cloud_buffer(cloud) -> Tuple[np.ndarray, int, int]:
    """Accepts (n,4) float32 [x,y,z,i] (float4) or (n,8) float32 (pcl::PointXYZI's 32-byte layout)."""
    a = np.ascontiguousarray(cloud, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("cloud must be (n,4) or (n,8) float32")
    return a, a.shape[0], a.shape[1] * 4


class DepthEstimator:
    def __init__(self, device: int = -1):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        self._device = device
        self._parameters: Optional[DepthEstimatorParameters] = None
        self._camera: Optional[CameraPinhole] = None
        self._transform = None
        self._isInitializedConfig = False
        self._isInitialized = False
        self._isInitializedPointCloud = False
        self.ransac_seed = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._lib.mld_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc == _capi.MLD_OK:
            return
        msg = self._lib.mld_last_error(self._h)
        msg = msg.decode("utf-8", "replace") if msg else ""
        if rc == _capi.MLD_ERR_PCL_INVALID:
            raise ExceptionPclInvalid()
        raise _capi.MldError(rc, msg)

    # -- DepthEstimator::InitConfig (DepthEstimator.cpp:129-154) ---------------------------------
    def InitConfig(self, parameters=None, printparams: bool = False) -> bool:
        if isinstance(parameters, (str, bytes)) or hasattr(parameters, "__fspath__"):
            p = DepthEstimatorParameters()
            p.fromFile(parameters)
            parameters = p
        elif parameters is None:
            parameters = DepthEstimatorParameters()
        self._parameters = parameters
        if printparams:
            parameters.print()
        self.close()
        rc = self._lib.mld_create(C.byref(parameters.c_struct), self._device, C.byref(self._h))
        if rc != _capi.MLD_OK:
            _capi.check(rc, None)
        self._params_on_device = bytes(parameters.c_struct)  # re-checked in Initialize
        self._isInitializedConfig = True
        self._isInitialized = False
        self._isInitializedPointCloud = False
        return True

    # -- DepthEstimator::Initialize (DepthEstimator.cpp:35-127) ----------------------------------
    def Initialize(self, camera: CameraPinhole, transform_lidar_to_cam) -> bool:
        if not self._isInitializedConfig:
            raise RuntimeError("Call 'InitConfig' before calling 'Initialize'.")
        # the reference builds its modules from the live parameter block here (DepthEstimator.cpp:46-127): a block that was
        # changed since InitConfig is handed to the device again
        if bytes(self._parameters.c_struct) != self._params_on_device:
            self.InitConfig(self._parameters, False)
        T = np.ascontiguousarray(np.asarray(transform_lidar_to_cam, np.float64)[:3, :4])
        if T.shape != (3, 4):
            raise ValueError("transform_lidar_to_cam must be 4x4 or 3x4")
        self._camera, self._transform = camera, T.copy()
        W, H = camera.getImageSize()
        self._check(self._lib.mld_initialize(self._h, W, H, camera.focal_length_, camera.principal_point_x_,
                                             camera.principal_point_y_, T.ctypes.data_as(C.POINTER(C.c_double))))
        self._isInitialized = True
        return True

    def getParameters(self):
        return self._parameters

    def getCamera(self):
        return self._camera

    def getTransformLidarToCam(self):
        return self._transform

    # -- DepthEstimator::setInputCloud (DepthEstimator.cpp:220-312) -----------------------------
    def setInputCloud(self, cloud, groundPlane: Optional[GroundPlane] = None) -> Optional[GroundPlane]:
        if not self._isInitialized:
            raise RuntimeError("call of 'setInputCloud' without 'initialize'")
        a, n, stride = _cloud_buffer(cloud)
        self._n = n
        pl_c = None
        if self._parameters.do_use_ransac_plane:
            if groundPlane is None:  # DepthEstimator.cpp:275-278
                groundPlane = RansacPlane(self._parameters, self.ransac_seed)
            if not groundPlane.isSegmented():
                if isinstance(groundPlane, RansacPlane):
                    pl_c = groundPlane._as_c(capacity=max(n, 1))  # fitted on the GPU, sharing the cloud's H2D copy
                else:
                    # any other GroundPlane segments itself (virtual CalculateInliersPlane, DepthEstimator.cpp:281-283),
                    # e.g. SemanticPlane -- what tracklets_depth hands in
                    groundPlane.CalculateInliersPlane(cloud, self._parameters.ransac_plane_min_z, self._parameters.ransac_plane_max_z)
        seed = getattr(groundPlane, "seed", self.ransac_seed) if groundPlane is not None else 0
        self._check(self._lib.mld_set_cloud(self._h, a.ctypes.data if n else None, n, stride,
                                            C.byref(pl_c) if pl_c is not None else None, seed))
        if pl_c is not None:
            groundPlane._from_c(pl_c)
        self._isInitializedPointCloud = True
        return groundPlane

    # -- DepthEstimator::CalculateDepth (DepthEstimator.cpp:404-488) -----------------------------
    def CalculateDepth(self, *args, layout: Optional[str] = None):
        """CalculateDepth(points_image_cs, ransacPlane=None) -> (depths, resultType)
        CalculateDepth(pointCloud, points_image_cs, ransacPlane=None) -> (depths, resultType, ransacPlane)

        points_image_cs: 2xF (reference layout, rows u and v) or Fx2 array of pixel coordinates. A 2x2 array is ambiguous:
        pass layout="2xF" or layout="Fx2" (without it a 2x2 array raises)."""
        if len(args) >= 2 and np.ndim(args[0]) == 2 and np.ndim(args[1]) == 2:
            cloud, feats = args[0], args[1]
            plane = args[2] if len(args) > 2 else None
            plane = self.setInputCloud(cloud, plane)
            d, s = self._calculate(feats, plane, layout)
            return d, s, plane
        feats = args[0]
        plane = args[1] if len(args) > 1 else None
        return self._calculate(feats, plane, layout)

    @staticmethod
    def _features(feats, layout: Optional[str] = None) -> np.ndarray:
        """(F, 2) C-contiguous doubles == the memory of a column-major Eigen::Matrix2Xd (u0, v0, u1, v1, ...)."""
        f = np.asarray(feats, np.float64)
        if f.ndim != 2 or 2 not in f.shape:
            raise ValueError("features must be 2xF or Fx2")
        if layout not in (None, "2xF", "Fx2"):
            raise ValueError("layout must be '2xF' or 'Fx2'")
        if layout is None and f.shape == (2, 2):
            raise ValueError("a 2x2 feature array is ambiguous: pass layout='2xF' (reference layout) or layout='Fx2'")
        if layout == "2xF" or (layout is None and f.shape[0] == 2):
            if f.shape[0] != 2:
                raise ValueError("layout '2xF' needs two rows")
            f = f.T
        elif f.shape[1] != 2:
            raise ValueError("layout 'Fx2' needs two columns")
        return np.ascontiguousarray(f)

    def _calculate(self, feats, plane, layout: Optional[str] = None):
        if not self._isInitializedPointCloud:
            raise RuntimeError("call of 'CalculateDepth' without 'SetInputCloud'")
        f = self._features(feats, layout)
        F = f.shape[0]
        depths = np.empty(F, np.float64)
        status = np.empty(F, np.int32)
        pl_c = plane._as_c() if plane is not None else None
        self._check(self._lib.mld_calculate_depth(self._h, f.ctypes.data, F, depths.ctypes.data, status.ctypes.data,
                                                  C.byref(pl_c) if pl_c is not None else None))
        return depths, status

    # -- tracklets_depth batch adaptor (TrackletDepthModule::process, tracklet_depth_module.cpp:318,330) ---------
    def CalculateDepthPair(self, cloud_last, feats_last, plane_last, cloud_cur, feats_cur, plane_cur, layout: Optional[str] = None,
                           resident: bool = False):
        """Previous and current cloud with their own feature sets in one call (both clouds are on the device
        concurrently). cloud_last may be None (first frame): its depths are -1. Planes follow setInputCloud:
        with do_use_ransac_plane a None / un-segmented plane is fitted on the GPU and returned.

        resident=True (cloud_last is then ignored): the previous cloud is the one that was current in the last call and is still
        on the device with its pixel map -- one upload and one projection per frame (mld_calculate_depth_pair_resident).
        Returns (depths_last, depths_cur, plane_last, plane_cur)."""
        if not self._isInitialized:
            raise RuntimeError("call of 'setInputCloud' without 'initialize'")

        def prep(feats):
            f = np.asarray(feats, np.float64)
            return np.empty((0, 2), np.float64) if f.size == 0 else self._features(f, layout)

        fl, fc = prep(feats_last), prep(feats_cur)
        dl, dc = np.empty(len(fl), np.float64), np.empty(len(fc), np.float64)
        use_plane = bool(self._parameters.do_use_ransac_plane)
        a_cur, n_cur, stride = _cloud_buffer(cloud_cur)
        have_last = bool(self._lib.mld_has_resident_cloud(self._h)) if resident else cloud_last is not None
        if resident:
            a_last, n_last = None, (self._n if have_last else 0)
        elif cloud_last is not None:
            a_last, n_last, stride_l = _cloud_buffer(cloud_last)
            if stride_l != stride:
                raise ValueError("both clouds must use the same point layout")
        else:
            a_last, n_last = None, 0
        planes, cplanes = [plane_last, plane_cur], [None, None]
        sizes = [n_last, n_cur]
        if use_plane:
            for i in range(2):
                if i == 0 and not have_last:
                    continue
                if planes[i] is None:
                    planes[i] = RansacPlane(self._parameters, self.ransac_seed)
                if not planes[i].isSegmented() and not isinstance(planes[i], RansacPlane):  # e.g. SemanticPlane: segments itself
                    if i == 0 and resident:
                        raise ValueError("an un-segmented non-RANSAC plane for the resident previous cloud needs the cloud: segment it first")
                    planes[i].CalculateInliersPlane(cloud_last if i == 0 else cloud_cur, self._parameters.ransac_plane_min_z,
                                                    self._parameters.ransac_plane_max_z)
                cplanes[i] = planes[i]._as_c(capacity=max(sizes[i], 1) if not planes[i].isSegmented() else 0)
        common = (a_cur.ctypes.data, n_cur, fc.ctypes.data if len(fc) else None, len(fc), dc.ctypes.data if len(fc) else None, None,
                  C.byref(cplanes[1]) if cplanes[1] is not None else None, stride, self.ransac_seed)
        prev = (fl.ctypes.data if len(fl) else None, len(fl), dl.ctypes.data if len(fl) else None, None,
                C.byref(cplanes[0]) if cplanes[0] is not None else None)
        if resident:
            self._check(self._lib.mld_calculate_depth_pair_resident(self._h, *prev, *common))
        else:
            self._check(self._lib.mld_calculate_depth_pair(self._h, a_last.ctypes.data if a_last is not None else None, n_last, *prev, *common))
        for i in range(2):
            if cplanes[i] is not None:
                planes[i]._from_c(cplanes[i])
        self._n = n_cur
        self._isInitializedPointCloud = True
        return dl, dc, planes[0], planes[1]

    # -- DepthCalculationStatistics (DepthEstimator.cpp:1039-1090) ---------------------------------------------
    def getDepthCalcStats(self, status) -> dict:
        """Counters per DepthResultType of a status array, computed on the GPU."""
        s = np.ascontiguousarray(status, np.int32).ravel()
        hist = (C.c_int64 * 21)()
        self._check(self._lib.mld_status_histogram_host(self._h, s.ctypes.data if len(s) else None, len(s), hist))
        names = {v: k for k, v in {
            "Unspecified": 0, "Success": 1, "RadiusSearchInsufficientPoints": 2, "HistogramNoLocalMax": 3,
            "TresholdDepthGlobalGreaterMax": 4, "TresholdDepthGlobalSmallerMin": 5, "TresholdDepthLocalGreaterMax": 6,
            "TresholdDepthLocalSmallerMin": 7, "TriangleNotPlanar": 8, "TriangleNotPlanarInsufficientPoints": 9,
            "CornerBehindCamera": 10, "PlaneViewrayNotOrthogonal": 11, "PcaIsPoint": 12, "PcaIsLine": 13, "PcaIsCubic": 14,
            "InsufficientRoadPoints": 15, "SuccessRoad": 16, "RegionGrowingNearestSeedNotAvailable": 17,
            "RegionGrowingSeedsOutOfRange": 18, "RegionGrowingInsufficientPoints": 19, "SuccessRegionGrowing": 20}.items()}
        return {names[i]: int(hist[i]) for i in range(21)}

    def statusHistogramDevice(self, d_status: int, n: int, stream: int = 0) -> np.ndarray:
        hist = (C.c_int64 * 21)()
        self._check(self._lib.mld_status_histogram_device(self._h, d_status, n, hist, stream or None))
        return np.array(list(hist), np.int64)

    def packFeaturePointsDevice(self, d_uv: int, d_depth: int, n: int, d_out: int, stream: int = 0) -> None:
        """matches_msg_depth_ros/FeaturePoint {float32 u, v, d} for n features, device buffers."""
        self._check(self._lib.mld_pack_feature_points_device(self._h, d_uv, d_depth, n, d_out, stream or None))

    # -- stand-alone RansacPlane::CalculateInliersPlane -------------------------------------------
    def estimateGroundPlane(self, cloud, seed: int = 0) -> RansacPlane:
        a, n, stride = _cloud_buffer(cloud)
        plane = RansacPlane(self._parameters, seed)
        pl_c = plane._as_c(capacity=max(n, 1))
        it = C.c_int32(0)
        self._check(self._lib.mld_estimate_ground_plane(self._h, a.ctypes.data if n else None, n, stride, seed, C.byref(pl_c), C.byref(it)))
        plane._from_c(pl_c)
        plane.iterations = it.value
        return plane

    def _semantic_plane(self, plane: SemanticPlane, cloud) -> None:
        a, n, stride = _cloud_buffer(cloud)
        pl_c = plane._as_c(capacity=max(n, 1))
        lab = plane.semantic_image_
        gl = np.ascontiguousarray(plane.groundplane_label_, np.int32)
        T = plane.cam_.transform_cam_lidar
        self._check(self._lib.mld_semantic_ground_plane(self._handle_for_plane(), a.ctypes.data if n else None, n, stride, lab.ctypes.data,
                                                        lab.shape[1], lab.shape[0], plane.cam_.f, plane.cam_.cu, plane.cam_.cv,
                                                        T.ctypes.data, gl.ctypes.data if len(gl) else None, len(gl),
                                                        plane.inlier_threshold_, C.byref(pl_c)))
        plane._from_c(pl_c)

    def semanticGroundLabelled(self, plane: SemanticPlane, cloud) -> np.ndarray:
        """bool per point: projects onto a ground-labelled pixel (the set RansacPlane.cpp:201-222 keeps); parity view."""
        a, n, stride = _cloud_buffer(cloud)
        lab = plane.semantic_image_
        gl = np.ascontiguousarray(plane.groundplane_label_, np.int32)
        T = plane.cam_.transform_cam_lidar
        out = np.zeros(max(n, 1), np.uint8)
        self._check(self._lib.mld_semantic_ground_labelled(self._handle_for_plane(), a.ctypes.data if n else None, n, stride, lab.ctypes.data,
                                                           lab.shape[1], lab.shape[0], plane.cam_.f, plane.cam_.cu, plane.cam_.cv, T.ctypes.data,
                                                           gl.ctypes.data if len(gl) else None, len(gl), out.ctypes.data))
        return out[:n].astype(bool)

    def _handle_for_plane(self):
        if not self._h:
            raise RuntimeError("Call 'InitConfig' before fitting a ground plane")
        return self._h

    # -- debug views (DepthEstimator.h:116-164) ------------------------------------------------
    def getPixelMap(self) -> np.ndarray:
        W, H = self._camera.getImageSize()
        out = np.empty((H, W), np.int32)
        self._check(self._lib.mld_get_pixel_map(self._h, out.ctypes.data))
        return out

    def getNeighbors(self, u: float, v: float, scale_w: float = 1.0, scale_h: float = 1.0) -> np.ndarray:
        cap = self._lib.mld_neighbor_capacity()
        out = np.empty(cap, np.int32)
        k = C.c_int(0)
        self._check(self._lib.mld_get_neighbors(self._h, u, v, scale_w, scale_h, out.ctypes.data, cap, C.byref(k)))
        return out[: min(k.value, cap)].copy()

    def getVisible(self) -> np.ndarray:
        out = np.zeros(max(self._n, 1), np.uint8)
        nv = C.c_int64(0)
        self._check(self._lib.mld_get_visible(self._h, out.ctypes.data, C.byref(nv)))
        return out[: self._n].astype(bool)

    def getPointsCloudCameraCs(self) -> np.ndarray:
        out = np.zeros((max(self._n, 1), 3), np.float64)
        self._check(self._lib.mld_get_points_camera(self._h, out.ctypes.data))
        return out[: self._n]

    def getVisiblePoints(self):
        """(_pointIndex, _points_cs_image_visible as (nvis, 2), camera-frame depth per visible point), compacted on the
        device in cloud order (DepthEstimator.cpp:189-207)."""
        cap = max(self._n, 1)
        idx = np.empty(cap, np.int32)
        img = np.empty((cap, 2), np.float64)
        dep = np.empty(cap, np.float64)
        nv = C.c_int64(0)
        self._check(self._lib.mld_get_visible_points(self._h, idx.ctypes.data, img.ctypes.data, dep.ctypes.data, cap, C.byref(nv)))
        k = nv.value
        return idx[:k].copy(), img[:k].copy(), dep[:k].copy()

    def getPointIndex(self) -> np.ndarray:
        """_pointIndex: raw index of every visible point, cloud order."""
        return self.getVisiblePoints()[0]

    def getPointsCloudImageCs(self) -> np.ndarray:
        """_points_cs_image_visible as (nvis, 2): projection of the visible points, cloud order (DepthEstimator.h:139)."""
        return self.getVisiblePoints()[1]

    def getPointDepthCamVisible(self, index: int) -> float:
        """_points_cs_camera(2, _pointIndex[index]) (DepthEstimator.h:116-118)."""
        return float(self.getVisiblePoints()[2][index])

    # -- batched sequences (frames are independent; see bench.py) ---------------------------------
    def processFramesDevice(self, d_points: int, n_points: int, frame_pitch_points: int, stride_bytes: int, d_uv: int, F: int,
                            d_depth: int, d_status: int, nframes: int, road: bool = False, seed: int = 0,
                            d_plane_coeffs_out: int = 0, stream: int = 0) -> None:
        """Raw device pointers (e.g. torch ``tensor.data_ptr()``); enqueues on ``stream`` without synchronising."""
        self._check(self._lib.mld_process_frames_device(self._h, d_points, n_points, frame_pitch_points, stride_bytes, d_uv, F,
                                                        d_depth, d_status, nframes, int(road), seed,
                                                        d_plane_coeffs_out or None, stream or None))

    def processFramesDeviceSemantic(self, d_points: int, n_points: int, frame_pitch_points: int, stride_bytes: int, d_labels: int,
                                    label_w: int, label_h: int, cam: "SemanticPlane.Camera", ground_labels, inlier_threshold: float,
                                    d_uv: int, F: int, d_depth: int, d_status: int, nframes: int, d_plane_coeffs_out: int = 0,
                                    d_plane_rc_out: int = 0, stream: int = 0) -> None:
        """SemanticPlane fit per frame on the device + depth estimation with the road path: the per-frame work of
        TrackletDepthModule::process (tracklet_depth_module.cpp:269-330) for a device-resident sequence."""
        gl = np.ascontiguousarray(sorted(int(x) for x in set(ground_labels)), np.int32)
        T = np.ascontiguousarray(cam.transform_cam_lidar, np.float64)
        self._check(self._lib.mld_process_frames_device_semantic(self._h, d_points, n_points, frame_pitch_points, stride_bytes, d_labels,
                                                                 label_w, label_h, cam.f, cam.cu, cam.cv, T.ctypes.data,
                                                                 gl.ctypes.data if len(gl) else None, len(gl), float(inlier_threshold), d_uv, F,
                                                                 d_depth, d_status, nframes, d_plane_coeffs_out or None,
                                                                 d_plane_rc_out or None, stream or None))

    def processFramesDevicePlanes(self, d_points: int, n_points: int, frame_pitch_points: int, stride_bytes: int, d_plane_coeffs: int,
                                  d_inlier_bits: int, d_uv: int, F: int, d_depth: int, d_status: int, nframes: int, stream: int = 0) -> None:
        """Device-resident sequence with caller-provided planes (nframes x 4 floats, nframes inlier bitmasks over raw indices)."""
        self._check(self._lib.mld_process_frames_device_planes(self._h, d_points, n_points, frame_pitch_points, stride_bytes, d_plane_coeffs,
                                                               d_inlier_bits, d_uv, F, d_depth, d_status, nframes, stream or None))

    def processFramesHost(self, points: np.ndarray, uv: np.ndarray, depth: np.ndarray, status: np.ndarray, road: bool = False,
                          seed: int = 0, plane_coeffs_out: Optional[np.ndarray] = None) -> None:
        """points (nframes, n, 4|8) float32, uv (nframes, F, 2) float64 -> depth (nframes, F) f64, status (nframes, F) i32."""
        nframes, n, k = points.shape
        F = uv.shape[1]
        self.processFramesHostPtr(points.ctypes.data, n, n, k * 4, uv.ctypes.data, F, depth.ctypes.data, status.ctypes.data,
                                  nframes, road, seed, plane_coeffs_out.ctypes.data if plane_coeffs_out is not None else 0)

    def processFramesHostPtr(self, points: int, n_points: int, frame_pitch_points: int, stride_bytes: int, uv: int, F: int,
                             depth: int, status: int, nframes: int, road: bool = False, seed: int = 0, plane_coeffs_out: int = 0):
        self._check(self._lib.mld_process_frames_host(self._h, points, n_points, frame_pitch_points, stride_bytes, uv, F, depth,
                                                      status, nframes, int(road), seed, plane_coeffs_out or None))

    def profileEnable(self, on: bool = True) -> None:
        self._check(self._lib.mld_profile_enable(self._h, int(on)))

    def profileRead(self):
        """{class: (total ms, launches)} for clear / project_scatter / ransac / feature_depth (= gather + solve + rest) and
        its three parts, frames sampled."""
        ms = (C.c_double * 7)()
        ln = (C.c_int64 * 7)()
        fr = C.c_int64(0)
        self._check(self._lib.mld_profile_read(self._h, ms, ln, C.byref(fr)))
        names = ("map_clear", "project_scatter", "ransac", "feature_depth", "feature_gather", "feature_solve", "feature_rest")
        return {n: (ms[i], ln[i]) for i, n in enumerate(names)}, fr.value

    def fusedChunkFrames(self) -> int:
        """Frames per fused K1 + gather launch of device-resident non-road sequences, 0 when that pipeline is off."""
        return int(self._lib.mld_fused_chunk_frames(self._h))

    def setSemanticExact(self, on: bool = True) -> None:
        """SemanticPlane fits of this estimator in PCL's sequential float accumulation order (bit-identical to the reference's
        coefficients and inlier set) instead of the default double-precision moments; see mld_set_semantic_exact."""
        self._check(self._lib.mld_set_semantic_exact(self._h, int(on)))

    def hostPipelineStats(self) -> dict:
        """Counters of processFramesHost since creation (mld_host_pipeline_stats)."""
        out = (C.c_int64 * 4)()
        self._check(self._lib.mld_host_pipeline_stats(self._h, out))
        return dict(zip(("h2d_bytes", "d2h_bytes", "frames_packed", "frames_direct"), [int(x) for x in out]))

    def chunkFrames(self) -> int:
        return int(self._lib.mld_chunk_frames(self._h))

    def kernelLaunchCount(self) -> int:
        return int(self._lib.mld_kernel_launch_count(self._h))

    @property
    def handle(self):
        return self._h
