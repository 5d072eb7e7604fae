mkdir -p gpurun_out
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 2>&1 | grep '^{"metric"' | tee gpurun_out/bench_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))"
done
