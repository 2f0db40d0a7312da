python -m pytest tests/test_parity_gpu.py -m gpu -q --no-header -rf --timeout 900 -k "long_sequence" > gpurun_out/test34.log 2>&1; tail -3 gpurun_out/test34.log
ncu --set full --clock-control none --import-source on -k regex:semantic_label -s 4 -c 1 -f -o gpurun_out/prof_semantic_label_r1b python scripts/bench_semantic.py 512 > gpurun_out/ncu_semlabel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:semantic_select -s 4 -c 1 -f -o gpurun_out/prof_semantic_select_r1b python scripts/bench_semantic.py 512 > gpurun_out/ncu_semselect.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
