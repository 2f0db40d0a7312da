cd "${GRAFT_REPO_ROOT:-.}"
TAG=r2h
export MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_E2E_FRAMES=16 MLD_BENCH_NO_PARITY=1
q() { python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],3))"; }
q base
MLD_SOLVE_PRIO=1 q solve_prio
MLD_FUSE_CHUNK=256 q chunk256
MLD_FUSE_CHUNK=384 q chunk384
MLD_FUSE_CHUNK=768 q chunk768
MLD_FUSE_SERIAL=1 q serial
MLD_OVERLAP=2 q overlap2
MLD_OVERLAP=4 q overlap4
export MLD_BENCH_FRAMES=2048
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 45 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
for k in fused_project_gather feature_solve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
ls -la gpurun_out | grep $TAG
