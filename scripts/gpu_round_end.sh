bash scripts/gpu_profile.sh
bash scripts/gpu_bench_lines.sh r1e
python scripts/bench_semantic.py 4096 2>&1 | tail -1 > gpurun_out/bench_r1e_semantic.json; cut -c1-200 gpurun_out/bench_r1e_semantic.json
