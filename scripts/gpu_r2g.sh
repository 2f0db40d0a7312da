cd "${GRAFT_REPO_ROOT:-.}"
TAG=r2g
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/test_$TAG.log 2>&1; tail -8 gpurun_out/test_$TAG.log
python scripts/pipe_timing.py 4000 | grep -v "unused\|road:"
for wl in kitti road dense; do
 for pipe in 1 0; do
  MLD_PIPE=$pipe MLD_BENCH_CPU_SECONDS=2 MLD_BENCH_E2E_FRAMES=64 timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl > gpurun_out/bench_${TAG}_${wl}_p$pipe.json 2> gpurun_out/bench_${TAG}_${wl}_p$pipe.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_${wl}_p$pipe.json").read().strip().splitlines()[-1])
    print("$wl pipe=$pipe", round(d["value"]), "f/s ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["parity"])
except Exception as e:
    print("$wl pipe=$pipe FAILED", e); print(open("gpurun_out/bench_${TAG}_${wl}_p$pipe.err").read()[-1500:])
PY
 done
done
