run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1 ${2:-kitti} fps',round(d['value']), {k:(round(v['avg_launch_ms']*1000,1)) if isinstance(v,dict) and v['avg_launch_ms'] else None for k,v in r['per_kernel'].items()}, d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run base
for v in scap8 scap10 sbtb64; do MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so run $v; done
MLD_FUSE_SERIAL=1 MLD_CUDA_LIB=$PWD/build/variants/libmld_scap8.so run scap8_serial
MLD_CUDA_LIB=$PWD/build/variants/libmld_scap8.so run scap8 dense
