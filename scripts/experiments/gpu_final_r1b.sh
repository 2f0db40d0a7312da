TAG=r1c
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test_$TAG.log 2>&1; tail -3 gpurun_out/test_$TAG.log
( time python bench.py ) > gpurun_out/bench_$TAG.log 2>&1; grep '^{' gpurun_out/bench_$TAG.log | tail -1 > gpurun_out/bench_${TAG}_kitti.json; tail -4 gpurun_out/bench_$TAG.log | cut -c1-300
python bench.py --impl reference > gpurun_out/bench_${TAG}_ref.log 2>&1; grep '^{' gpurun_out/bench_${TAG}_ref.log | tail -1 | cut -c1-400
MLD_BENCH_CPU_SECONDS=4 python bench.py --workload road --steps 4 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${TAG}_road.json
MLD_BENCH_CPU_SECONDS=4 python bench.py --workload dense --steps 4 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${TAG}_dense.json
python scripts/bench_semantic.py 4096 2>&1 | tail -1 > gpurun_out/bench_${TAG}_semantic.json
export MLD_BENCH_FRAMES=512 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
for k in project_scatter feature_gather feature_solve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
unset MLD_BENCH_FRAMES MLD_BENCH_E2E_FRAMES MLD_BENCH_CPU_SECONDS
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_semantic_batch_$TAG.csv python scripts/bench_semantic.py 512 > /dev/null 2>&1
ls gpurun_out | grep $TAG | wc -l
