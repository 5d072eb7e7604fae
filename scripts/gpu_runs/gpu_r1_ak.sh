set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-llm 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'xattn', round(d['roofline_xattn']['achieved']), round(d['kernel_ms_per_step']['xattn_pairs'],3))"; }
run A=1
run A=2
