cd "${GRAFT_REPO_ROOT:-.}"
TAG=r2k
MLD_BENCH_CPU_SECONDS=6 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_kitti.json 2> gpurun_out/bench_${TAG}_kitti.err; tail -c 1500 gpurun_out/bench_${TAG}_kitti.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_kitti.json").read().strip().splitlines()[-1])
print("kitti", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "float4 e2e", round(d["e2e"]["float4_input_frames_per_s"]), d["parity"])
print("roof", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"], d["roofline"].get("traffic_note"))
for k,v in (d["other_workloads"] or {}).items(): print(k, round(v["value"]), v["roofline_path"]["frac"], v["parity"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
MLD_BENCH_SEQ_FRAMES=30000 MLD_BENCH_E2E_FRAMES=64 timeout 900 python bench.py --steps 1 --warmup 3 --workload seq100k > gpurun_out/bench_${TAG}_seq.json 2> gpurun_out/bench_${TAG}_seq.err; tail -c 800 gpurun_out/bench_${TAG}_seq.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${TAG}_seq.json').read().strip().splitlines()[-1]); print('seq', round(d['value']), d['ms_per_step'], d['scaling'], d['config']['workload'][:120])"
for t in 4 8 14; do MLD_PACK_THREADS=$t MLD_BENCH_NO_OTHERS=1 MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_FRAMES=2048 timeout 300 python bench.py --steps 2 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pack threads $t e2e', round(d['e2e']['value']))"; done
MLD_HOST_PACK=0 MLD_BENCH_NO_OTHERS=1 MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_FRAMES=2048 timeout 300 python bench.py --steps 2 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no pack e2e', round(d['e2e']['value']))"
