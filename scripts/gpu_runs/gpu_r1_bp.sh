timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "small_m" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_qformer_gpu.py -m gpu -x -q -k "deterministic" 2>&1 | tail -1
for i in 1 2; do timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['relation_tokens_per_sec']; print('value',round(d['value']),'llm tokens/s',round(r['value']),'ms/image',round(r['ms_per_image'],2))"; done
