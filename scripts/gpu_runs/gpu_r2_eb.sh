# head-of-image chains on three streams (mask chain / PatchEmbed / embeddings): Q-Former suites + bench without LLM legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_qformer_gpu.py tests/test_batching_gpu.py tests/test_multi_gpu.py -x -q 2>&1 | grep -E "passed|failed|^E  |Error" | head -12 | tee gpurun_out/r2_eb_tests.log
timeout 900 python bench.py --no-llm --no-cpu-baseline > gpurun_out/r2_eb_bench.json 2> gpurun_out/r2_eb_bench.err
tail -2 gpurun_out/r2_eb_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_eb_bench.json') if x.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ms_per_step_separate_calls'], d['roofline']['achieved'], d['clocks'], d['results']['sha1'])
print(sum(d['kernel_ms_per_step'].values()))
P
