"""DepthEstimatorParameters -- the reference's parameter object with the reference's field names.

Mirrors Mono_Lidar::DepthEstimatorParameters (monolidar_fusion/include/monolidar_fusion/
DepthEstimatorParameters.h:12-172) and its loader fromFile (src/DepthEstimatorParameters.cpp:16-114).
Parsing is done by the C ABI (mld_params_from_yaml) so that the C++ shim and this mirror agree.
"""
from __future__ import annotations

import ctypes as C

from . import _capi

_FIELDS = [n for n, _ in _capi.MldParams._fields_ if n != "reserved0"]


class DepthEstimatorParameters:
    """Attribute access by the reference's names, e.g. ``p.pixelarea_search_witdh = 6``."""

    def __init__(self):
        object.__setattr__(self, "_c", _capi.MldParams())
        _capi.load().mld_default_params(C.byref(self._c))

    def __getattr__(self, name):
        if name in _FIELDS:
            return getattr(self._c, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name not in _FIELDS:
            raise AttributeError(f"DepthEstimatorParameters has no field {name!r}")
        if isinstance(value, bool):
            value = int(value)
        setattr(self._c, name, value)

    def fromFile(self, filePath: str) -> None:
        """DepthEstimatorParameters::fromFile: flat OpenCV-YAML, absent keys read as 0."""
        rc = _capi.load().mld_params_from_yaml(str(filePath).encode(), C.byref(self._c))
        if rc != _capi.MLD_OK:
            _capi.check(rc, None)

    def as_dict(self) -> dict:
        return {n: getattr(self._c, n) for n in _FIELDS}

    def print(self) -> None:  # DepthEstimatorParameters::print
        print("DepthEstimator parameters: \n")
        for k, v in self.as_dict().items():
            print(f"{k}: {v}")

    def copy(self) -> "DepthEstimatorParameters":
        q = DepthEstimatorParameters()
        C.memmove(C.byref(q._c), C.byref(self._c), C.sizeof(_capi.MldParams))
        return q

    @property
    def c_struct(self) -> _capi.MldParams:
        return self._c

    @staticmethod
    def reference_yaml(do_use_ransac_plane: int = 1) -> "DepthEstimatorParameters":
        """The values of monolidar_fusion/parameters.yaml with do_use_depth_segmentation forced to 0
        (the shipped value 1 makes the reference throw "Region growing not supported!",
        DepthEstimator.cpp:608). Keys the yaml does not hold keep the struct defaults."""
        p = DepthEstimatorParameters()
        p.pixelarea_search_witdh = 6
        p.pixelarea_search_height = 9
        p.radiusSearch_count_min = 1
        p.histogram_segmentation_bin_witdh = 0.3
        p.histogram_segmentation_min_pointcount = 3
        p.do_use_depth_segmentation = 0
        p.pca_treshold_2_1_rel_min = 1.5
        p.do_use_ransac_plane = do_use_ransac_plane
        p.ransac_plane_distance_treshold = 0.3
        p.viewray_plane_orthoganality_treshold = 0.03
        return p
