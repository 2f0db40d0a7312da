run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 timeout 300 python bench.py --workload ${2:-kitti} --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$1 ${2:-kitti} fps',round(d['value']), 'chunk', d['config']['chunk_frames_per_launch'], 'dom', r['kernel'], round(r['frac'],3), 'share', r['share_of_step'], {k:(round(v['avg_launch_ms']*1000,1), v['launches']) if isinstance(v,dict) and v['avg_launch_ms'] else None for k,v in r['per_kernel'].items()}, d['parity']['status_exact'])
    elif 'Error' in l or 'error' in l: print(l)
"; }
run fused; MLD_FUSE_SERIAL=1 run fused_serial; run fused; MLD_FUSE_SERIAL=1 run fused_serial
MLD_FUSE_SERIAL=1 MLD_FUSE_CHUNK=256 run fused_serial_c256; MLD_FUSE_SERIAL=1 MLD_FUSE_CHUNK=1024 run fused_serial_c1024
MLD_FUSE_SERIAL=1 run fused_serial dense; run fused dense
