python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test15.log 2>&1; tail -4 gpurun_out/test15.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
MLD_OVERLAP=1 run base_serial; run base; run base; MLD_CHUNK_FRAMES=128 run base
for v in scap12b128 scap10b128 scap8b128 scap16b128; do export MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so; MLD_OVERLAP=1 run ${v}_serial; run $v; run $v;  MLD_CHUNK_FRAMES=128 run $v; unset MLD_CUDA_LIB; done
