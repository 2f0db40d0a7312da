python -m pytest tests/test_semantic_plane.py -m gpu -q --no-header -rf --timeout 900 -k "labelled_set" > gpurun_out/test36.log 2>&1; grep -n "^E" gpurun_out/test36.log | head -12 | cut -c1-400
for v in s0q1 s0q0 s1q0; do
  export MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so
  echo "== $v"; python -m pytest tests/test_semantic_plane.py -m gpu -q --no-header --timeout 900 -k "labelled_set" 2>&1 | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:semantic_label -s 8 -c 2 --csv --log-file gpurun_out/l_$v.csv python scripts/bench_semantic.py 512 > /dev/null 2>&1
  grep -o 'semantic_[a-z]*_kernel\|"[0-9.]*"$' gpurun_out/l_$v.csv | paste - - | sort | uniq -c | head -2
done
