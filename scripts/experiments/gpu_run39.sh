# fused K1 + gather launches (MLD_FUSE=1): parity of the batched paths, then throughput A/B
MLD_FUSE=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf --timeout 600 -k "batched or long_sequence or epoch or overflow" > gpurun_out/test39.log 2>&1; tail -4 gpurun_out/test39.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 4 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} fps',round(d['value']), 'launches', d['gpu_launches'], d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run base; MLD_FUSE=1 run fused; run base; MLD_FUSE=1 run fused
MLD_FUSE=1 MLD_CHUNK_FRAMES=64 run fused_c64; MLD_FUSE=1 MLD_CHUNK_FRAMES=256 run fused_c256
MLD_FUSE=1 run fused dense; run base dense
