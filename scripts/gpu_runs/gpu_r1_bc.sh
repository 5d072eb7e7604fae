timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_bc.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bc.json'))
r=d['relation_tokens_per_sec']
print('value',d['value'],'sep_calls_ms',d.get('ms_per_step_separate_calls'),'e2e',d['e2e']['value'],'llm tokens/s',r['value'],'ms/image',r['ms_per_image'], r['kernel_ms_per_image'])
PY
