set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -k "streamk or gemm" 2>&1 | tail -6
timeout 600 python scripts/kbench.py streamk 2>&1 | tee gpurun_out/kbench_u.jsonl | cut -c1-200
timeout 1200 python -m pytest tests/test_llm_gpu.py -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_u.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}); print(d['relation_tokens_per_sec'])"
