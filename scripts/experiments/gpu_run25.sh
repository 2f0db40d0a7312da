# new reference-golden GPU tests + persistent-K1 sweep
python -m pytest tests/test_ref_golden.py tests/test_golden.py -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test25.log 2>&1; tail -3 gpurun_out/test25.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
run base
MLD_OVERLAP=1 run serial
for b in 2 3 4 6; do MLD_K1_PERSIST=$b run persist$b; done
MLD_OVERLAP=1 MLD_K1_PERSIST=4 run serial_persist4
MLD_OVERLAP=1 MLD_K1_PERSIST=8 run serial_persist8
MLD_K1_PERSIST=3 MLD_OVERLAP_MODE=prio run prio_persist3
MLD_K1_PERSIST=4 MLD_K1_PERSIST=4 run persist4_road road
run base_road road
