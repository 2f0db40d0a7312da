TAG=${1:-r1d}
( time python bench.py ) > gpurun_out/bench_$TAG.log 2>&1; grep '^{' gpurun_out/bench_$TAG.log | tail -1 > gpurun_out/bench_${TAG}_kitti.json; tail -4 gpurun_out/bench_$TAG.log | cut -c1-200
python bench.py --impl reference 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_${TAG}_reference_arm.json
MLD_BENCH_CPU_SECONDS=4 python bench.py --workload dense --steps 4 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${TAG}_dense.json
MLD_BENCH_CPU_SECONDS=4 python bench.py --workload road --steps 4 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_${TAG}_road.json
for w in kitti dense road; do python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_$w.json')); r=d['roofline']
print('$w','fps',round(d['value']),'e2e',round(d['e2e']['value']),'dom',r['kernel'],'frac',round(r['frac'],3),'path',round(r['path']['frac'],3),'cpu',round(d['cpu_baseline']['value']),'launches',d['gpu_launches'],'traffic',r['traffic'])
PY
done
