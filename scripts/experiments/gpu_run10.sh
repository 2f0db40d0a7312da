run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
for c in 16 32 64; do
MLD_CHUNK_FRAMES=$c MLD_OVERLAP=1 run rep0
MLD_CHUNK_FRAMES=$c MLD_OVERLAP=1 MLD_EXP_REPEAT_K2=1 run rep1
MLD_CHUNK_FRAMES=$c MLD_OVERLAP=1 MLD_EXP_REPEAT_K2=3 run rep3
done
