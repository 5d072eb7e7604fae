# round 2: tcgen05 K4 (self_attn_pairs_kernel): parity, pipeline parity, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qformer_gpu.py -x -q -k "self_attn or qformer or relation or graph or stage" 2>&1 | grep -E "passed|failed|^E|Error|error" | head -20
timeout 600 python bench.py --steps 5 --no-llm --no-cpu-baseline > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_e.json').read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
print({k: round(v,3) for k,v in d['kernel_ms_per_step'].items()})
PY
tail -3 gpurun_out/r2_bench_e.err
