python -m pytest tests -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test45.log 2>&1; tail -4 gpurun_out/test45.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1 ${2:-kitti} fps',round(d['value']), 'chunk', d['config']['chunk_frames_per_launch'], 'launches', d['gpu_launches'], d['parity'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run fused road; MLD_FUSE=2 run sep road; run fused road
python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-160; MLD_FUSE=2 python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-160
run fused kitti
