# after the decode attention rework: LLM parity at depth, bench with LLM legs, ncu of the product kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_bench_sizes_gpu.py -x -q -s 2>&1 | grep -E "passed|failed|^E  |Error|parity|checksum" | head -12 | tee gpurun_out/r2_du_tests.log
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/r2_du_bench.json 2> gpurun_out/r2_du_bench.err
tail -2 gpurun_out/r2_du_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_du_bench.json') if x.startswith('{')][-1])
print('value', d['value'], d['ms_per_step'], 'pruned', d['last_layer_selected_rows_only']['value'])
for k in ('e2e_cfg3','e2e_cfg5'):
    e=d[k]; print(k, e['ms_per_step'], e['relation_tokens_per_sec'], e['llm_batch'], '| per image:', e['llm_one_image_per_batch']['ms_per_step'], e['llm_one_image_per_batch']['relation_tokens_per_sec'])
r=d['relation_tokens_per_sec']
print('stacked', r['value'], r['ms_per_batch'], r['roofline']['bound'], r['roofline']['frac']); print(r['kernel_ms_per_batch'])
s=r['single_image_batch']; print('single', s['value'], s['ms_per_batch'], s['roofline']['bound'], s['roofline']['frac']); print(s['kernel_ms_per_batch'])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn_tma -s 4 -c 2 -o gpurun_out/r2_prof_decode_attn_tma2 -f python scripts/decode_attn_probe.py 800 65 > /dev/null 2>&1
ls -la gpurun_out/r2_prof_decode_attn_tma2* | awk '{print $5, $9}'
