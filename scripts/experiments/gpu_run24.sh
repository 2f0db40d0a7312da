# state check after container re-creation: GPU parity suite, default bench (as the driver runs it), reference arm
python -m pytest tests -m gpu -q --no-header -rf -x --timeout 900 > gpurun_out/test24.log 2>&1; tail -3 gpurun_out/test24.log
( time python bench.py ) > gpurun_out/bench24.log 2>&1; tail -c 3000 gpurun_out/bench24.log
( time python bench.py --impl reference ) > gpurun_out/bench24_ref.log 2>&1; tail -c 1500 gpurun_out/bench24_ref.log
nproc; lscpu | grep -i "model name"
