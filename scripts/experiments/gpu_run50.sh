python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_r1g_kitti.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1g_kitti.json')); r=d['roofline']
print('fps',round(d['value']),'e2e',round(d['e2e']['value']),'dom',r['kernel'],'achieved',round(r['achieved']),'frac',round(r['frac'],3),'traffic',r['traffic'],'path',round(r['path']['frac'],3),round(r['path']['frac_without_map_term'],3),'cpu',round(d['cpu_baseline']['value']),d['cpu_baseline']['kind'],'launches',d['gpu_launches'],d['clocks'])
PY
