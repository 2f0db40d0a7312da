python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test41.log 2>&1; tail -4 gpurun_out/test41.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1 ${2:-kitti} fps',round(d['value']), 'chunk', d['config']['chunk_frames_per_launch'], 'dom', r['kernel'], round(r['frac'],3), {k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in r['per_kernel'].items()}, d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run fused; run fused; MLD_FUSE=0 run separate; run fused dense; run fused road
python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-160
