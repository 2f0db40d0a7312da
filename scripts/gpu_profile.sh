# tests + ncu launch list + --set full captures of the two top kernels + default bench line: bash scripts/gpu_profile.sh TAG
TAG=${1:-r2m}
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test_$TAG.log 2>&1; tail -3 gpurun_out/test_$TAG.log
export MLD_BENCH_FRAMES=2560 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_NO_OTHERS=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 45 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
# (the solve launches alternate between the main instantiation and the small one for windows of 10-16 points: an even skip count lands on the main one)
for k in ${MLD_PROFILE_KERNELS:-fused_project_gather feature_solve}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
unset MLD_BENCH_FRAMES MLD_BENCH_E2E_FRAMES MLD_BENCH_CPU_SECONDS MLD_BENCH_NO_OTHERS
( time python bench.py ) > gpurun_out/bench_$TAG.log 2>&1; grep '^{' gpurun_out/bench_$TAG.log | tail -1 > gpurun_out/bench_${TAG}_kitti.json; tail -4 gpurun_out/bench_$TAG.log | cut -c1-300
ls gpurun_out | grep $TAG | wc -l
