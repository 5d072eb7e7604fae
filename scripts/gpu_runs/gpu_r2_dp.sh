# medium-M split-K path for the LLM's N = hidden Linears: tests, GEMM probe through the new dispatch, LLM parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or embed or patch" 2>&1 | grep -E "passed|failed|^E  |Error" | head -12 | tee gpurun_out/r2_dp_tests.log
OPSG_PROBE_MEDIUM=1 python scripts/gemm_medium_m.py 800 400 2>&1 | tee gpurun_out/r2_dp_gemm.log
timeout 1200 python -m pytest tests/test_parity_bench_sizes_gpu.py tests/test_batching_gpu.py tests/test_llm_gpu.py tests/test_llama_gpu.py tests/test_qformer_gpu.py -x -q -s 2>&1 | grep -E "passed|failed|^E  |Error|parity|checksum" | head -12 | tee -a gpurun_out/r2_dp_tests.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2_dp_bench.json 2> gpurun_out/r2_dp_bench.err
tail -2 gpurun_out/r2_dp_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_dp_bench.json') if x.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'pruned', d['last_layer_selected_rows_only']['value'])
for k in ('e2e_cfg3','e2e_cfg5'):
    e=d[k]; print(k, e['ms_per_step'], e['relation_tokens_per_sec'], e['llm_batch'], '| per image:', e['llm_one_image_per_batch']['ms_per_step'], e['llm_one_image_per_batch']['relation_tokens_per_sec'])
r=d['relation_tokens_per_sec']
print('stacked', r['value'], r['ms_per_batch'], r['roofline']['bound'], r['roofline']['frac']); print(r['kernel_ms_per_batch'])
P
