python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test20.log 2>&1; tail -4 gpurun_out/test20.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()}, d['parity'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
MLD_OVERLAP=1 run serial road; run ov3 road; MLD_CHUNK_FRAMES=64 run ov3 road; run ov3 kitti
