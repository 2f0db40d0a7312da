run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
for ov in 3 4 6; do for c in 32 64; do MLD_OVERLAP=$ov MLD_CHUNK_FRAMES=$c run ov$ov; done; done
for v in scap8 scap12 scap8b128; do export MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so; MLD_OVERLAP=1 run ${v}_serial; run $v; unset MLD_CUDA_LIB; done
