ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_semantic_batch.csv python scripts/bench_semantic.py 512 > gpurun_out/ncu_sem_batch.log 2>&1
tail -2 gpurun_out/ncu_sem_batch.log | cut -c1-200
