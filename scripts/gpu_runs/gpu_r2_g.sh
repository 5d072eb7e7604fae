# round 2: 2-GPU identity test + bench under torchrun (strong scaling, 2 ranks)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | grep -E "passed|failed|skipped|^E" | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
    print('N=2 value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'sha', d['results']['sha1'][:12], 'scaling', d['scaling'])
    print('tokens/s', d['relation_tokens_per_sec']['value'], 'cfg3', d['e2e_cfg3']['relation_tokens_per_sec'], 'cfg5', d['e2e_cfg5']['relation_tokens_per_sec'], d['e2e_cfg5']['object_pairs_per_sec'])
except Exception as e:
    print('parse failed', e)
PY
tail -5 gpurun_out/r2_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-budget-s 40 --no-llm 2>/dev/null | cut -c1-300
