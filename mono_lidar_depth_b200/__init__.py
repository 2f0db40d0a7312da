"""mono_lidar_depth_b200 -- B200-native (sm_100a) implementation of monolidar_fusion's per-frame
depth-estimation hot path behind the reference's DepthEstimator interface.

Layout: csrc/ holds the CUDA kernels and the C ABI (include/mld_c_api.h -> libmld_cuda.so); the Python
modules mirror the reference's operator interface on top of it. See DESIGN.md / INTEGRATION.md.
"""
from ._capi import MldError, LIB_PATH  # noqa: F401
from .params import DepthEstimatorParameters  # noqa: F401
from .depth_estimator import (  # noqa: F401
    CameraPinhole,
    DepthEstimator,
    ExceptionPclInvalid,
    GroundPlane,
    RansacPlane,
    SemanticPlane,
)
from . import sharding, synth  # noqa: F401

# Mono_Lidar::DepthResultType (eDepthResultType.h:9-31)
DepthResultType = {
    "Unspecified": 0, "Success": 1, "RadiusSearchInsufficientPoints": 2, "HistogramNoLocalMax": 3,
    "TresholdDepthGlobalGreaterMax": 4, "TresholdDepthGlobalSmallerMin": 5, "TresholdDepthLocalGreaterMax": 6,
    "TresholdDepthLocalSmallerMin": 7, "TriangleNotPlanar": 8, "TriangleNotPlanarInsufficientPoints": 9,
    "CornerBehindCamera": 10, "PlaneViewrayNotOrthogonal": 11, "PcaIsPoint": 12, "PcaIsLine": 13, "PcaIsCubic": 14,
    "InsufficientRoadPoints": 15, "SuccessRoad": 16, "RegionGrowingNearestSeedNotAvailable": 17,
    "RegionGrowingSeedsOutOfRange": 18, "RegionGrowingInsufficientPoints": 19, "SuccessRegionGrowing": 20,
}
