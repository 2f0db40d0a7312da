run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} fps',round(d['value']), 'chunk', d['config']['chunk_frames_per_launch'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run fused_prof; MLD_BENCH_NO_PROF=1 run fused_noprof; run fused_prof; MLD_BENCH_NO_PROF=1 run fused_noprof
MLD_FUSE=0 run sep_prof; MLD_FUSE=0 MLD_BENCH_NO_PROF=1 run sep_noprof
MLD_FUSE_CHUNK=1024 run fused_c1024; MLD_FUSE_CHUNK=768 run fused_c768
