python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test16.log 2>&1; tail -4 gpurun_out/test16.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python scripts/sanitize_workload.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_workload.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; tail -5 gpurun_out/sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_workload.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo synccheck rc=$?; tail -3 gpurun_out/sanitizer_synccheck.log
MLD_BENCH_CPU_SECONDS=4 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.log 2>&1; tail -1 gpurun_out/bench3.log | cut -c1-600
