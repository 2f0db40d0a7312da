"""End-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck) on the build that ships: every kernel of the library
runs at least once on full KITTI-shaped sweeps (so that the geometry tail is reached: the status counts are asserted), the
fused and separate-launch pipelines with more chunks than slots, both SemanticPlane fit modes, the host-buffer pipeline on
32-byte records with packing threads, the pair adaptors and the debug views."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("MLD_FUSE_CHUNK", "2")   # 14 frames -> 7 chunks over 5 slots: slots reused
os.environ.setdefault("MLD_CHUNK_FRAMES", "2")
os.environ.setdefault("MLD_PACK_THREADS", "3")
import torch  # noqa: E402

from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, GroundPlane, SemanticPlane, synth  # noqa: E402

seen = np.zeros(21, np.int64)


def count(s):
    global seen
    seen += np.bincount(np.asarray(s).ravel(), minlength=21)[:21]


def run(mode, fuse):
    for k, v in (("MLD_FEATURE_MODE", mode), ("MLD_FUSE", fuse)):
        if v:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    cfg = synth.default_config(road=True)
    n = synth.points_per_frame(cfg)
    est = DepthEstimator()
    est.InitConfig(DepthEstimatorParameters.reference_yaml(1))
    est.Initialize(synth.kitti_camera(), synth.KITTI_T_LIDAR_TO_CAM)
    cloud = synth.points_host(cfg, 3, 0)
    uv = synth.features_host(cfg, 3, 0, 2000)
    d, s, plane = est.CalculateDepth(cloud, uv, None)  # RANSAC plane fitted on the GPU
    count(s)
    est.getPixelMap(); est.getNeighbors(600.0, 250.0); est.getVisible(); est.getPointsCloudCameraCs()
    est.getDepthCalcStats(s)
    idx, img, dep_vis = est.getVisiblePoints()  # visible-order stream compaction
    assert np.array_equal(idx, np.nonzero(est.getVisible())[0])
    # a plane whose inlier set covers the ground returns: the road path succeeds (SuccessRoad)
    dist = np.abs(cloud[:, 2] + 1.73)
    gp = GroundPlane(np.array([0, 0, 1, 1.73], np.float32), np.nonzero(np.isfinite(dist) & (dist < 0.1))[0].astype(np.int32))
    count(est.CalculateDepth(uv, gp)[1])
    lab = np.zeros((376, 1241), np.uint8)
    lab[230:, :] = 7
    for exact in (False, True):
        est.setSemanticExact(exact)
        sp = SemanticPlane(lab, SemanticPlane.Camera(718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM), (6, 7, 8, 9), 0.2, est)
        sp.CalculateInliersPlane(cloud)  # semantic_label / semantic_select (/ semantic_exact_fit) kernels
        count(est.CalculateDepth(cloud, uv, sp)[1])
    est.setSemanticExact(False)
    est.CalculateDepthPair(cloud, uv, None, cloud, uv, None)
    est.CalculateDepthPair(None, uv[:100], None, cloud, uv, None, resident=True)  # previous cloud stays on the device
    F, nframes = 1500, 14
    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
    fu = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
    dep = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
    sta = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    synth.points_device(est, cfg, 5, 0, nframes, pts.data_ptr())
    synth.features_device(est, cfg, 5, 0, nframes, F, fu.data_ptr())
    for road in (False, True):
        for _ in range(2):  # second pass: slots and map epochs reused
            est.processFramesDevice(pts.data_ptr(), n, n, 16, fu.data_ptr(), F, dep.data_ptr(), sta.data_ptr(), nframes, road=road, seed=11, stream=st)
        torch.cuda.synchronize()
        count(sta.cpu().numpy())
    labs = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(lab, (nframes, 376, 1241)))).cuda()
    cam = SemanticPlane.Camera(718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM)
    est.processFramesDeviceSemantic(pts.data_ptr(), n, n, 16, labs.data_ptr(), 1241, 376, cam, [6, 7, 8, 9], 0.2, fu.data_ptr(), F, dep.data_ptr(),
                                    sta.data_ptr(), nframes, 0, 0, st)
    torch.cuda.synchronize()
    count(sta.cpu().numpy())
    # host-buffer pipeline on pcl::PointXYZI records (packing threads + whole-record chunks)
    hp = np.zeros((nframes, n, 8), np.float32)
    hp[:, :, :3] = pts.cpu().numpy()[:, :, :3]
    hu = fu.cpu().numpy()
    hd, hs = np.empty((nframes, F)), np.empty((nframes, F), np.int32)
    est.processFramesDevice(pts.data_ptr(), n, n, 16, fu.data_ptr(), F, dep.data_ptr(), sta.data_ptr(), nframes, road=True, seed=11, stream=st)
    torch.cuda.synchronize()
    est.processFramesHost(hp, hu, hd, hs, road=True, seed=11)
    assert np.array_equal(hs, sta.cpu().numpy())
    est.statusHistogramDevice(sta.data_ptr(), nframes * F)
    print("mode", mode or "split", "fuse", fuse or "1", "ok", flush=True)


for m, fu_ in ((None, None), (None, "0"), ("warp", None)):
    run(m, fu_)
print("status counts", seen[:17])
for st_ in (1, 2, 3, 8, 9, 11, 16):
    assert seen[st_] > 0, f"status {st_} never reached: the workload does not cover the geometry tail"
print("every status of the tail reached")
