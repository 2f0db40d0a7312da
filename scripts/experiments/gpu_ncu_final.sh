TAG=${1:-r1final}
export MLD_BENCH_FRAMES=512 MLD_BENCH_E2E_FRAMES=32 MLD_BENCH_CPU_SECONDS=1
# launch list of the default configuration (3 overlapping streams, chunk 128)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
for k in project_scatter feature_gather feature_solve; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
MLD_BENCH_FRAMES=256 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_road_$TAG.csv python bench.py --workload road --steps 2 --warmup 3 > /dev/null 2>&1
ls gpurun_out | grep $TAG
