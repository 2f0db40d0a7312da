#!/usr/bin/env python
"""Per-frame latency of the drop-in call sequence (BASELINE.json configs[0] shape: one KITTI-shaped frame, 2000 features):
setInputCloud + CalculateDepth through the C ABI with HOST buffers, one frame at a time like the reference's ROS callback
(10 Hz lidar), next to the CPU oracle on the same frames. Prints one JSON line. Not part of bench.py's contract."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, SemanticPlane, synth  # noqa: E402

cfg = synth.default_config()
frames = [synth.points_host(cfg, 5, f) for f in range(16)]
feats = [synth.features_host(cfg, 5, f, 2000) for f in range(16)]


def timed(fn, reps):
    for i in range(8):
        fn(i)
    ts = []
    for i in range(reps):
        t = time.perf_counter()
        fn(i)
        ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts)), 1e3 * float(np.percentile(ts, 95))


out = {"metric": "ms_per_frame", "workload": "one KITTI-shaped frame per call (120000 points, 2000 features), host buffers"}
for road in (0, 1):
    est = DepthEstimator()
    est.InitConfig(DepthEstimatorParameters.reference_yaml(road))
    est.Initialize(synth.kitti_camera(), synth.KITTI_T_LIDAR_TO_CAM)
    med, p95 = timed(lambda i: est.CalculateDepth(frames[i % 16], feats[i % 16], None), 300)
    out["gpu_road_ransac" if road else "gpu_non_road"] = {"median_ms": med, "p95_ms": p95}
    if not road:  # the same frames as 32-byte pcl::PointXYZI records in pageable memory: what the drop-in caller hands over
        frames32 = [synth.points_host_xyzi32(cfg, 5, f) for f in range(16)]
        med, p95 = timed(lambda i: est.CalculateDepth(frames32[i % 16], feats[i % 16], None), 300)
        out["gpu_non_road_pointxyzi"] = {"median_ms": med, "p95_ms": p95}
lab = np.zeros((376, 1241), np.uint8)
lab[200:] = 7
cam = SemanticPlane.Camera(718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM)


def semantic_frame(i):
    sp = SemanticPlane(lab, cam, (6, 7, 8, 9), 0.1, est)
    sp.CalculateInliersPlane(frames[i % 16])
    est.CalculateDepth(frames[i % 16], feats[i % 16], sp)


med, p95 = timed(semantic_frame, 300)
out["gpu_semantic_plane_plus_road"] = {"median_ms": med, "p95_ms": p95}
p = O.yaml_params()
p.do_use_ransac_plane = 0
orc = O.Oracle(p)
k = synth.kitti_camera()
orc.initialize(1241, 376, k.focal_length_, k.principal_point_x_, k.principal_point_y_, synth.KITTI_T_LIDAR_TO_CAM)


def cpu_frame(i):
    orc.set_cloud(frames[i % 16])
    orc.calculate_depth(feats[i % 16])


for threads in (1, 0):
    O.lib().orc_set_num_threads(threads)
    med, p95 = timed(cpu_frame, 60)
    out[f"cpu_oracle_non_road_{'1_thread' if threads == 1 else 'all_threads'}"] = {"median_ms": med, "p95_ms": p95, "threads": O.lib().orc_get_max_threads()}
print(json.dumps(out))
