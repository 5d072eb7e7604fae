# mma.sync decode attention + row-CTA LayerNorm for <= 4096 rows: kernel tests, LLM tests, LLM bench legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_batching_gpu.py tests/test_llm_gpu.py tests/test_llama_gpu.py tests/test_parity_bench_sizes_gpu.py -x -q 2>&1 | grep -E "passed|failed|^E  |Error" | head -30 | tee gpurun_out/r2_db_tests.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2_db_bench.json 2> gpurun_out/r2_db_bench.err
tail -3 gpurun_out/r2_db_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_db_bench.json') if x.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'pruned', d['last_layer_selected_rows_only']['value'])
for k in ('e2e_cfg3','e2e_cfg5'):
    e=d[k]; print(k, e['ms_per_step'], e['relation_tokens_per_sec'], e['llm_batch'], '| per image:', e['llm_one_image_per_batch']['ms_per_step'], e['llm_one_image_per_batch']['relation_tokens_per_sec'])
r=d['relation_tokens_per_sec']
print('stacked', r['value'], r['ms_per_batch'], r['roofline']['bound'], r['roofline']['frac']); print(r['kernel_ms_per_batch'])
s=r['single_image_batch']; print('single', s['value'], s['ms_per_batch'], s['roofline']['bound'], s['roofline']['frac']); print(s['kernel_ms_per_batch'])
P
