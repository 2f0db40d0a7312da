set -x
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test3.log 2>&1; tail -15 gpurun_out/test3.log
export MLD_BENCH_FRAMES=128 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:feature_depth -s 20 -c 2 -f -o gpurun_out/prof_feature_r1 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_feature.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_scatter -s 20 -c 2 -f -o gpurun_out/prof_project_r1 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_project.log 2>&1
ls -la gpurun_out/
