timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "layernorm" 2>&1 | tail -1
for i in 1 2 3; do timeout 300 python bench.py --steps 16 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'ln ms',round(d['kernel_ms_per_step']['layernorm_bf16'],3))"; done
