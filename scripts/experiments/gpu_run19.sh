run() { MLD_BENCH_NO_PARITY=1 MLD_OVERLAP=1 MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
run base
for v in d_nofp64 d_noatom d_ldg; do export MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so; run $v; unset MLD_CUDA_LIB; done
