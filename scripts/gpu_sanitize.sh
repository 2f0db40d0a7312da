for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_workload.py > gpurun_out/r1b_$tool.txt 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r1b_$tool.txt
done
