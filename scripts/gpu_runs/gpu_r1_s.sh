set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_qformer_gpu.py -q -x -k graph 2>&1 | grep -E "^E|Error|passed|failed|test_qformer_gpu.py:" | head -20
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['relation_tokens_per_sec']['value'], d['relation_tokens_per_sec']['ms_per_image'])"
