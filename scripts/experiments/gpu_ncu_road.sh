TAG=${1:-r1road}
export MLD_BENCH_FRAMES=192 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1 MLD_OVERLAP=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --workload road --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ransac_cluster -s 4 -c 1 -f -o gpurun_out/prof_ransac_$TAG python bench.py --workload road --steps 2 --warmup 3 > gpurun_out/ncu_ransac_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:feature_depth_thread -s 4 -c 1 -f -o gpurun_out/prof_featroad_$TAG python bench.py --workload road --steps 2 --warmup 3 > gpurun_out/ncu_featroad_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
