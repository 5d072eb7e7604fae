# ring LayerNorm (register-resident gamma / beta) final check: kernel tests + Q-Former suites + bench without LLM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_qformer_gpu.py -x -q 2>&1 | grep -E "passed|failed|^E  |Error" | head -12 | tee gpurun_out/r2_dm_tests.log
timeout 900 python bench.py --no-llm --no-cpu-baseline > gpurun_out/r2_dm_bench.json 2> gpurun_out/r2_dm_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_dm_bench.json') if x.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step']['layernorm_bf16'], d['roofline_hbm_kernels']['layernorm_bf16']['achieved'], d['roofline']['achieved'], d['clocks'], d['results']['sha1'])
P
