# compute-sanitizer on the build that ships (scripts/sanitize_workload.py asserts that every status of the tail is reached)
TAG=${1:-r2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_workload.py > gpurun_out/${TAG}_$tool.txt 2>&1; echo "$tool rc=$?"; tail -6 gpurun_out/${TAG}_$tool.txt
done
