# usage: bash scripts/gpu_ncu.sh <tag>   -- launch list + full captures of the two hot kernels (small bench run)
TAG=${1:-r1}
export MLD_BENCH_FRAMES=256 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:feature_depth_thread -s 8 -c 1 -f -o gpurun_out/prof_feature_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_feature_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_scatter -s 8 -c 1 -f -o gpurun_out/prof_project_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_project_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
