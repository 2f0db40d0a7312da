run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} fps',round(d['value']), 'launches', d['gpu_launches'], d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
MLD_CHUNK_FRAMES=256 run base_c256; MLD_CHUNK_FRAMES=192 run base_c192
MLD_FUSE=1 MLD_CHUNK_FRAMES=256 run fused_c256; MLD_FUSE=1 MLD_CHUNK_FRAMES=384 run fused_c384; MLD_FUSE=1 MLD_CHUNK_FRAMES=512 run fused_c512
MLD_FUSE=1 MLD_CHUNK_FRAMES=256 MLD_OVERLAP=2 run fused_c256_2slots; MLD_FUSE=1 MLD_CHUNK_FRAMES=256 MLD_OVERLAP=4 run fused_c256_4slots
