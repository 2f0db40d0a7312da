for v in sp8 sp8t128 sp4t128 sp2; do
  export MLD_CUDA_LIB=$PWD/build/variants/libmld_$v.so
  echo "== $v"; python scripts/bench_semantic.py 4096 2>&1 | tail -1 | cut -c1-140
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:semantic -s 16 -c 4 --csv --log-file gpurun_out/l_$v.csv python scripts/bench_semantic.py 512 > /dev/null 2>&1
  grep -o 'semantic_[a-z]*_kernel\|"[0-9.]*"$' gpurun_out/l_$v.csv | paste - - | sort | uniq -c | head -4
done
