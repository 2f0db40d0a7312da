"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the
library runs at least once, in every K2 mode, on a reduced KITTI-shaped sweep."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, SemanticPlane, synth  # noqa: E402


def run(mode):
    if mode:
        os.environ["MLD_FEATURE_MODE"] = mode
    else:
        os.environ.pop("MLD_FEATURE_MODE", None)
    cfg = synth.default_config()
    cfg.azimuth_steps = 600  # 38400 points: keeps the sanitizer run short
    n = synth.points_per_frame(cfg)
    est = DepthEstimator()
    est.InitConfig(DepthEstimatorParameters.reference_yaml(1))
    est.Initialize(synth.kitti_camera(), synth.KITTI_T_LIDAR_TO_CAM)
    cloud = synth.points_host(cfg, 3, 0)
    uv = synth.features_host(cfg, 3, 0, 600)
    d, s, plane = est.CalculateDepth(cloud, uv, None)
    est.getPixelMap(); est.getNeighbors(600.0, 250.0); est.getVisible(); est.getPointsCloudCameraCs()
    est.getDepthCalcStats(s)
    idx, img, dep_vis = est.getVisiblePoints()  # visible-order stream compaction
    assert np.array_equal(idx, np.nonzero(est.getVisible())[0])
    lab = np.zeros((376, 1241), np.uint8)
    lab[200:, :] = 7
    sp = SemanticPlane(lab, SemanticPlane.Camera(718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM), (6, 7, 8, 9), 0.1, est)
    sp.CalculateInliersPlane(cloud)  # semantic_label / semantic_select kernels
    est.CalculateDepth(cloud, uv, sp)
    est.CalculateDepthPair(cloud, uv, None, cloud, uv, None)
    F, nframes = 500, 9
    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
    fu = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
    dep = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
    sta = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
    synth.points_device(est, cfg, 5, 0, nframes, pts.data_ptr())
    synth.features_device(est, cfg, 5, 0, nframes, F, fu.data_ptr())
    for road in (False, True):
        est.processFramesDevice(pts.data_ptr(), n, n, 16, fu.data_ptr(), F, dep.data_ptr(), sta.data_ptr(), nframes, road=road, seed=11,
                                stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    hp, hu = pts.cpu().numpy(), fu.cpu().numpy()
    hd, hs = np.empty((nframes, F)), np.empty((nframes, F), np.int32)
    est.processFramesHost(hp, hu, hd, hs, road=True, seed=11)
    assert np.array_equal(hs, sta.cpu().numpy())
    print("mode", mode or "split", "ok", np.bincount(hs.ravel(), minlength=17)[:17])


for m in (None, "fused", "warp"):
    run(m)
