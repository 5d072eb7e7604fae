# round 2, 2 GPUs of one box: 1-vs-2-rank identity test + strong-scaling bench under torchrun (LLM legs with grouped decode on every rank)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | grep -E "passed|failed|skipped|^E  |Error" | head -10 | tee gpurun_out/r2_2gpu_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -3 gpurun_out/r2_bench_2gpu.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_bench_2gpu.json') if x.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'sha1', d['results']['sha1'])
for k in ('e2e_cfg3','e2e_cfg5'):
    if k not in d: continue
    e=d[k]; print(k, e['ms_per_step'], e['relation_tokens_per_sec'], e['llm_batch'], '| per image:', e['llm_one_image_per_batch']['ms_per_step'], e['llm_one_image_per_batch']['relation_tokens_per_sec'])
r=d.get('relation_tokens_per_sec')
if r: print('stacked', r['value'], r['ms_per_batch'], r['sequences'])
P
