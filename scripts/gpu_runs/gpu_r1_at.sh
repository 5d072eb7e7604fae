timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_at.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_at.json'))
for k,v in d.items():
    if k in ('config',): continue
    print(k, json.dumps(v)[:1500])
PY
