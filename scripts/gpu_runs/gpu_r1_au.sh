timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
OPSG_PDL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_pdl0.json
OPSG_PDL=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_pdl1.json
python - <<'PY'
import json
for f in ('gpurun_out/bench_pdl0.json','gpurun_out/bench_pdl1.json'):
    try:
        d=json.load(open(f))
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('value_one_call_per_step'), 'llm', d['relation_tokens_per_sec']['value'], d['relation_tokens_per_sec']['ms_per_image'])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-600:])
PY
