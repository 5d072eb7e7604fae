set -x
mkdir -p gpurun_out
OPSG_FOLD_LN=1 timeout 1500 python -m pytest tests/test_qformer_gpu.py -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
OPSG_FOLD_LN=1 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-llm 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm', round(d['roofline']['achieved']), d['kernel_ms_per_step'])"; }
run OPSG_FOLD_LN=1
run OPSG_FOLD_LN=0
