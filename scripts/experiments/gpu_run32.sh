python -m pytest tests/test_semantic_plane.py -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test32.log 2>&1; tail -3 gpurun_out/test32.log
python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:semantic -s 16 -c 8 --csv --log-file gpurun_out/launches_semantic_batch3.csv python scripts/bench_semantic.py 512 > gpurun_out/ncu_sem_batch3.log 2>&1
grep -o 'semantic_[a-z]*_kernel\|"[0-9.]*"$' gpurun_out/launches_semantic_batch3.csv | paste - - | sort | uniq -c | head
