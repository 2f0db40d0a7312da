run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} chunk',d['config']['chunk_frames_per_launch'],'fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
export MLD_CUDA_LIB=$PWD/build/variants/libmld_noevict.so
MLD_OVERLAP=1 run noevict_serial; MLD_CHUNK_FRAMES=32 run noevict_ov3; run noevict_ov3
unset MLD_CUDA_LIB
export MLD_BENCH_FRAMES=256 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1 MLD_OVERLAP=1
ncu --set full --clock-control none --import-source on -k regex:project_scatter -s 8 -c 1 -f -o gpurun_out/prof_project_r1v6 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_project_r1v6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:feature_depth_thread -s 8 -c 1 -f -o gpurun_out/prof_feature_r1v6 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_feature_r1v6.log 2>&1
