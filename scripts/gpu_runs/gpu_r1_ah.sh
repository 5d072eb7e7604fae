set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-llm 2>&1 | tail -1 | tee gpurun_out/bench_ah.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}); print(d['e2e']['value']); print(d['roofline']['achieved'], d['kernel_ms_per_step'])"
