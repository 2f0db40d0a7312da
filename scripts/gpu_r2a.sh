# round 2, first run of the persistent pipeline: its parity tests, the whole GPU suite, bench lines pipeline vs chunked
TAG=${1:-r2a}
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf -x -k "persistent_pipeline" --timeout 300 > gpurun_out/test_pipe_$TAG.log 2>&1; tail -15 gpurun_out/test_pipe_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --timeout 600 > gpurun_out/test_$TAG.log 2>&1; tail -12 gpurun_out/test_$TAG.log
for wl in kitti road dense; do
  MLD_BENCH_CPU_SECONDS=4 timeout 600 python bench.py --steps 5 --warmup 3 --workload $wl > gpurun_out/bench_${TAG}_${wl}.json 2> gpurun_out/bench_${TAG}_${wl}.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_${wl}.json").read().strip().splitlines()[-1])
    print("$wl pipe", round(d["value"]), "f/s ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["parity"], d["roofline"] and (d["roofline"]["kernel"], round(d["roofline"]["frac"],3)))
except Exception as e:
    print("$wl pipe FAILED", e); print(open("gpurun_out/bench_${TAG}_${wl}.err").read()[-1500:])
PY
done
MLD_PIPE=0 MLD_BENCH_CPU_SECONDS=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_kitti_chunked.json 2> gpurun_out/bench_${TAG}_kitti_chunked.err; tail -c 600 gpurun_out/bench_${TAG}_kitti_chunked.json | cut -c1-300
for d in 6 20; do MLD_PIPE_DELAY=$d MLD_BENCH_NO_PARITY=1 MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=8 timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('delay $d', round(d['value']))"; done
MLD_PIPE_HINT=0 MLD_BENCH_NO_PARITY=1 MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=8 timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nohint', round(d['value']))"
for r in 16 64; do MLD_PIPE_RING=$r MLD_BENCH_NO_PARITY=1 MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=8 timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ring $r', round(d['value']))"; done
