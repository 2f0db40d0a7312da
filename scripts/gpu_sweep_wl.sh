# A/B of run-time knobs / library variants on one workload: bash scripts/gpu_sweep_wl.sh WORKLOAD "ENV1=a" "ENV2=b" ...
cd "${GRAFT_REPO_ROOT:-.}"
WL=$1; shift
export MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=16 MLD_BENCH_NO_OTHERS=1
for v in "$@"; do
  env $v python bench.py --workload $WL --steps 4 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$WL', round(d['value']), round(d['ms_per_step'],3), d['parity']['status_exact'], {k:round(x['avg_launch_ms'],4) for k,x in d['roofline']['per_kernel'].items() if isinstance(x,dict) and x.get('avg_launch_ms')})"
done
