# bulk-copy ring LayerNorm: tests + Q-Former bench (no LLM legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_qformer_gpu.py -x -q -k "layernorm or golden or stage or cfg" 2>&1 | grep -E "passed|failed|^E  |Error" | head -12 | tee gpurun_out/r2_dj_tests.log
timeout 900 python bench.py --no-llm --no-cpu-baseline > gpurun_out/r2_dj_bench.json 2> gpurun_out/r2_dj_bench.err
tail -2 gpurun_out/r2_dj_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_dj_bench.json') if x.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step']['layernorm_bf16'], d['roofline_hbm_kernels']['layernorm_bf16']['achieved'], d['roofline']['achieved'], d['clocks'], d['results']['sha1'])
P
