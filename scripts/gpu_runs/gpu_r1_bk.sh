start=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "default bench rc=$? wall=$(( $(date +%s) - start ))s"
start=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench_default.err
echo "reference arm rc=$? wall=$(( $(date +%s) - start ))s"
tail -c 600 gpurun_out/bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
for k,v in d.items(): print(k, json.dumps(v)[:400])
PY
