timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "small_m or argmax or layernorm" 2>&1 | tail -3
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep skinny | cut -c1-230
timeout 900 python -m pytest tests/test_llm_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_bd.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bd.json'))
r=d['relation_tokens_per_sec']
print('value',d['value'],'e2e',d['e2e']['value'],'llm tokens/s',r['value'],'ms/image',r['ms_per_image'], r['kernel_ms_per_image'])
PY
