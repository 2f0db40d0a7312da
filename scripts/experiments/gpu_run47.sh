python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test47.log 2>&1; tail -4 gpurun_out/test47.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1 ${2:-kitti} fps',round(d['value']), {k:(round(v['avg_launch_ms']*1000,1)) if isinstance(v,dict) and v['avg_launch_ms'] else None for k,v in r['per_kernel'].items()}, d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run tiled; MLD_CUDA_LIB=$PWD/build/variants/libmld_prev.so run prev; run tiled
MLD_FUSE_SERIAL=1 run tiled_serial; MLD_FUSE_SERIAL=1 MLD_CUDA_LIB=$PWD/build/variants/libmld_prev.so run prev_serial
run tiled dense; MLD_CUDA_LIB=$PWD/build/variants/libmld_prev.so run prev dense
run tiled road
