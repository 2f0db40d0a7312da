python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test49.log 2>&1; tail -3 gpurun_out/test49.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1 ${2:-kitti} fps',round(d['value']), d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run scap8
for v in scap6 mb7 scap6mb8; do MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so run $v; done
