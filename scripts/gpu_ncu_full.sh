# ncu --set full capture of the named kernels (one launch each) in the default bench command: bash scripts/gpu_ncu_full.sh TAG kernel...
TAG=$1; shift
cd "${GRAFT_REPO_ROOT:-.}"
export MLD_BENCH_FRAMES=2560 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=0 MLD_BENCH_NO_OTHERS=1 MLD_BENCH_NO_PARITY=1
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  ls -la gpurun_out/prof_${k}_$TAG.ncu-rep
done
