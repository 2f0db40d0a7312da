python -m pytest tests/test_semantic_plane.py tests/test_shim_cpp.py -m gpu -q --no-header -rf --timeout 900 > gpurun_out/test37.log 2>&1; tail -3 gpurun_out/test37.log; grep -n "^E" gpurun_out/test37.log | head -5 | cut -c1-300
python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:semantic -s 16 -c 8 --csv --log-file gpurun_out/l_sem37.csv python scripts/bench_semantic.py 512 > /dev/null 2>&1
grep -o 'semantic_[a-z]*_kernel\|"[0-9.]*"$' gpurun_out/l_sem37.csv | paste - - | sort | uniq -c | head -8
