#!/usr/bin/env python
"""Throughput of the production configuration: SemanticPlane fit + depth estimation with the road path per
frame, device-resident KITTI-shaped sequence, one mld_process_frames_device_semantic call per step. Prints one JSON line.
Not part of bench.py's contract (the headline metric is the non-road config); CUDA events, 3 warm-up steps."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, SemanticPlane, synth  # noqa: E402

nframes, F, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 2000, 4
cfg = synth.default_config()
n = synth.points_per_frame(cfg)
est = DepthEstimator()
est.InitConfig(DepthEstimatorParameters.reference_yaml(1))
est.Initialize(synth.kitti_camera(), synth.KITTI_T_LIDAR_TO_CAM)
pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
dep = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
sta = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
labs = torch.zeros((nframes, 376, 1241), dtype=torch.uint8, device="cuda")
labs[:, 200:, :] = 7
coeffs = torch.zeros((nframes, 4), dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
synth.points_device(est, cfg, 20261017, 0, nframes, pts.data_ptr(), stream=st)
synth.features_device(est, cfg, 20261017, 0, nframes, F, uv.data_ptr(), stream=st)
cam = SemanticPlane.Camera(718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM)


def step():
    est.processFramesDeviceSemantic(pts.data_ptr(), n, n, 16, labs.data_ptr(), 1241, 376, cam, (6, 7, 8, 9), 0.1, uv.data_ptr(), F,
                                    dep.data_ptr(), sta.data_ptr(), nframes, coeffs.data_ptr(), 0, st)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
s = sta.cpu().numpy()
print(json.dumps({"metric": "frames_per_sec", "workload": "semantic plane + road path per frame (production configuration)", "value": nframes / ms * 1e3,
                  "frames": nframes, "ms_per_step": ms, "success_fraction": float((s == 1).mean()), "success_road_fraction": float((s == 16).mean()),
                  "plane_of_frame_0": coeffs[0].cpu().tolist()}))
