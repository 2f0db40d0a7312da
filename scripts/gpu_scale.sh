N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.log 2>&1
grep '^{' gpurun_out/scale_n$N.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'fps',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],2),'clk',d['clocks'])"
tail -3 gpurun_out/scale_n$N.log | cut -c1-300
