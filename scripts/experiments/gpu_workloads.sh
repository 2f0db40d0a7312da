for w in kitti road dense; do
  MLD_BENCH_CPU_SECONDS=8 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.log 2>&1
  tail -1 gpurun_out/bench_$w.log | python -c "
import sys,json
l=sys.stdin.read()
try:
    d=json.loads(l)
    print('$w fps',round(d['value']),'depths/s %.3g'%d['feature_depths_per_sec'],'e2e',round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value'],1),'path frac %.3f'%d['roofline']['path']['frac'],{k:(round(v['avg_launch_ms']*1000,1) if isinstance(v,dict) and v['avg_launch_ms'] else None) for k,v in d['roofline']['per_kernel'].items()}, d['parity'])
except Exception as e:
    print('$w FAILED', e); print(l[-3000:])
"
done
