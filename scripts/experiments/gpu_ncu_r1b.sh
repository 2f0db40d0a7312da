TAG=${1:-r1b}
python -m pytest tests/test_semantic_plane.py -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test_sem_$TAG.log 2>&1; tail -3 gpurun_out/test_sem_$TAG.log
export MLD_BENCH_FRAMES=512 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1
# launch list of the default configuration (3 overlapping streams, chunk 128): per-launch time and DRAM bytes
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
for k in project_scatter feature_gather feature_solve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
# SemanticPlane kernels (one KITTI-shaped sweep)
cat > /tmp/sem_once.py <<'PY'
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import importlib.util, numpy as np
spec = importlib.util.spec_from_file_location("mk", "tests/golden/make_ref_golden.py"); MK = importlib.util.module_from_spec(spec); spec.loader.exec_module(MK)
from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, SemanticPlane, synth
est = DepthEstimator(); est.InitConfig(DepthEstimatorParameters.reference_yaml(0)); est.Initialize(synth.kitti_camera(), synth.KITTI_T_LIDAR_TO_CAM)
cloud, lab, gl, thr = MK.semantic_case(1)
for _ in range(3):
    p = SemanticPlane(lab, SemanticPlane.Camera(718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM), gl, thr, est); p.CalculateInliersPlane(cloud)
print(p.getModelCoeffs(), len(p.getInlinersIndex()))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:semantic --csv --log-file gpurun_out/launches_semantic_$TAG.csv python /tmp/sem_once.py > gpurun_out/ncu_sem_$TAG.log 2>&1
tail -2 gpurun_out/ncu_sem_$TAG.log
ls gpurun_out | grep $TAG
