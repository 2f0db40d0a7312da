# quick A/B of run-time knobs on the headline workload: bash scripts/gpu_sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...
cd "${GRAFT_REPO_ROOT:-.}"
export MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=16 MLD_BENCH_NO_PARITY=1 MLD_BENCH_NO_OTHERS=1
for v in "$@"; do
  env $v python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), round(d['ms_per_step'],3), {k:round(x['avg_launch_ms'],4) for k,x in d['roofline']['per_kernel'].items() if isinstance(x,dict) and x.get('avg_launch_ms')})"
done
