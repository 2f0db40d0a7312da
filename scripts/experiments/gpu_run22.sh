run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
for ov in 3 4 5 6; do MLD_OVERLAP=$ov run slots$ov; done
MLD_OVERLAP=4 MLD_CHUNK_FRAMES=96 run slots4; MLD_OVERLAP=6 MLD_CHUNK_FRAMES=96 run slots6; MLD_OVERLAP=3 MLD_CHUNK_FRAMES=192 run slots3; MLD_OVERLAP=3 MLD_CHUNK_FRAMES=256 run slots3
MLD_OVERLAP=4 run slots4 road; MLD_OVERLAP=6 run slots6 road
