timeout 900 python -m pytest tests/test_llm_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 1 0; do OPSG_LLM_SMALL_M=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_ba_$v.json; done
python - <<'PY'
import json
for v in (1,0):
    d=json.load(open(f'gpurun_out/bench_ba_{v}.json'))
    r=d['relation_tokens_per_sec']
    print('small_m',v,'value',d['value'],'e2e',d['e2e']['value'],'llm tokens/s',r['value'],'ms/image',r['ms_per_image'], r['kernel_ms_per_image'], r.get('tokens_checksum'))
PY
