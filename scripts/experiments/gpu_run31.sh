python -m pytest tests/test_semantic_plane.py tests/test_shim_cpp.py -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test31.log 2>&1; tail -4 gpurun_out/test31.log
python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-300
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_semantic_batch2.csv python scripts/bench_semantic.py 512 > gpurun_out/ncu_sem_batch2.log 2>&1
python scripts/sanitize_workload.py > /dev/null 2>&1 && timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_workload.py 2>&1 | tail -1
