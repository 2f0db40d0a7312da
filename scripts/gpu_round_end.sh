# round-end run on a B200 box: GPU test suite, ncu launch list of the default bench command, bench lines of every workload
TAG=${1:-r1f}
python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test_$TAG.log 2>&1; tail -3 gpurun_out/test_$TAG.log
MLD_BENCH_FRAMES=2048 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 45 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
bash scripts/gpu_bench_lines.sh $TAG
python scripts/bench_semantic.py 4096 2>&1 | tail -1 > gpurun_out/bench_${TAG}_semantic.json; cut -c1-200 gpurun_out/bench_${TAG}_semantic.json
