# multi-GPU bench lines on one box: weak scaling (10k frames per GPU) and the 100k-frame sequence (strong scaling)
N=${1:-2}; TAG=${2:-r2}; STEPS=${3:-5}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export MLD_BENCH_CPU_SECONDS=2 MLD_BENCH_E2E_FRAMES=256
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err; tail -c 600 gpurun_out/scale_${TAG}_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 1 --warmup 3 --workload seq100k > gpurun_out/seq100k_${TAG}_n$N.json 2> gpurun_out/seq100k_${TAG}_n$N.err; tail -c 600 gpurun_out/seq100k_${TAG}_n$N.err
python - <<PY
import json
for f in ("scale_${TAG}_n$N", "seq100k_${TAG}_n$N"):
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        print(f, round(d["value"]), "f/s ms/step", round(d["ms_per_step"],3), d["scaling"], "e2e", round(d["e2e"]["value"]), "per-rank", d.get("per_rank_ms_per_step"))
    except Exception as e: print(f, "FAILED", e)
PY
