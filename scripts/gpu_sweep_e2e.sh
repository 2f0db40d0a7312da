# A/B of the host-buffer pipeline's knobs on the end-to-end number: bash scripts/gpu_sweep_e2e.sh "ENV1=a" "ENV2=b" ...
cd "${GRAFT_REPO_ROOT:-.}"
export MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_NO_OTHERS=1 MLD_BENCH_FRAMES=2048 MLD_BENCH_NO_PROF=1
for v in "$@"; do
  env $v python bench.py --steps 2 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$v', 'e2e', round(e['value']), 'packed', round(e['frames_packed_fraction'],2), 'float4', round(e['float4_input_frames_per_s']))"
done
