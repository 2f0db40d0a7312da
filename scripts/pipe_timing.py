"""In-kernel timing of the persistent pipeline (MLD_PIPE_TIMING=1): where the block-cycles of one sequence go."""
import os, sys, time
os.environ["MLD_PIPE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, synth

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
est = DepthEstimator(); est.InitConfig(DepthEstimatorParameters.reference_yaml(do_use_ransac_plane=0)); est.Initialize(synth.kitti_camera(), synth.KITTI_T_LIDAR_TO_CAM)
cfg = synth.default_config(); n = synth.points_per_frame(cfg); F = 2000
pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda"); uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
depth = torch.empty((nframes, F), dtype=torch.float64, device="cuda"); status = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
synth.points_device(est, cfg, 1, 0, nframes, pts.data_ptr(), stream=st); synth.features_device(est, cfg, 1, 0, nframes, F, uv.data_ptr(), stream=st)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    est.processFramesDevice(pts.data_ptr(), n, n, 16, uv.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes, stream=st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
c = est.pipelineCounters()
tot = sum(v for k, v in c.items() if k != "warp_path_features")
print(f"{nframes} frames in {dt*1e3:.2f} ms = {nframes/dt:.0f} frames/s; aborted={est.pipelineAborted()}")
for k, v in c.items():
    if k == "warp_path_features": print(f"  {k}: {v} ({v/nframes:.1f} per frame)")
    else: print(f"  {k}: {v/1e6:.1f} Mcycles ({100*v/max(tot,1):.1f} %)  per frame {v/nframes/1965:.1f} us-blocks")
