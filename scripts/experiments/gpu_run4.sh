set -x
python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test4.log 2>&1; tail -25 gpurun_out/test4.log
MLD_BENCH_CPU_SECONDS=6 python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.log 2>&1; tail -c 3500 gpurun_out/bench2.log
for c in 16 32 128; do MLD_CHUNK_FRAMES=$c MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=256 python bench.py --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),'e2e',round(d['e2e']['value']),{k:(round(v['avg_launch_ms']*1000,1) if v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
"; done
