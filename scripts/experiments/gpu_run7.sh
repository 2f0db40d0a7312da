python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test7.log 2>&1; tail -5 gpurun_out/test7.log
run() { MLD_BENCH_CPU_SECONDS=1 MLD_BENCH_E2E_FRAMES=128 python bench.py --workload ${2:-kitti} --steps 3 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ${2:-kitti} fps',round(d['value']),{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()})
    elif 'Error' in l or 'error' in l: print(l)
"; }
MLD_OVERLAP=1 run base_serial; run base; MLD_OVERLAP=1 run base_serial road; run base road
for v in tcap8 ppt4; do export MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so; MLD_OVERLAP=1 run ${v}_serial; run $v; unset MLD_CUDA_LIB; done
export MLD_BENCH_FRAMES=256 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1 MLD_OVERLAP=1
MLD_CUDA_LIB=$PWD/build/variants/libmld_tcap8.so ncu --set full --clock-control none --import-source on -k regex:feature_depth_thread -s 8 -c 1 -f -o gpurun_out/prof_feature_tcap8 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_feature_tcap8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_scatter -s 8 -c 1 -f -o gpurun_out/prof_project_r1v5 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_project_r1v5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ransac_cluster -s 4 -c 1 -f -o gpurun_out/prof_ransac_r1v5 python bench.py --workload road --steps 2 --warmup 3 > gpurun_out/ncu_ransac_r1v5.log 2>&1
