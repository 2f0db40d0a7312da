# quick look at a kernel change on the GPU box: parity tests, headline bench line, serialised ncu launch list (per-kernel time and DRAM bytes)
# usage: bash scripts/gpu_quick.sh TAG [pytest -k expression]
TAG=${1:-q}
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_ref_golden.py -m gpu -q -x --no-header ${2:+-k "$2"} 2>&1 | tail -3
export MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=16 MLD_BENCH_NO_OTHERS=1
python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_$TAG.json
python -c "import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('$TAG fps', round(d['value']), 'ms', round(d['ms_per_step'],3), 'parity', d['parity']['status_exact'], {k:round(x['avg_launch_ms'],4) for k,x in d['roofline']['per_kernel'].items() if isinstance(x,dict) and x.get('avg_launch_ms')})"
MLD_BENCH_FRAMES=2560 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 45 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
python scripts/launch_list_summary.py gpurun_out/launches_$TAG.csv gpurun_out/traffic_$TAG.json 512 "$TAG" > /dev/null
python -c "
import json; t=json.load(open('gpurun_out/traffic_$TAG.json'))
tot=0
for k,v in t.items():
    if isinstance(v,dict) and 'avg_launch_us_ncu_serialised' in v:
        print('  %-24s %8.2f us  %8.1f MB  (%d launches)'%(k, v['avg_launch_us_ncu_serialised'], v['dram_bytes_per_launch']/1e6, v['launches'])); tot+=v['avg_launch_us_ncu_serialised']
print('  serialised sum per 512 frames: %.1f us'%tot)"
