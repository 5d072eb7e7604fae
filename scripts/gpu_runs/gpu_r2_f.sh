# K4 tcgen05: 3 vs 4 TMEM buffers / softmax warpgroups
for lib in libopsg_b200.so libopsg_b200_sa4.so; do
  export OPSG_B200_LIB=$PWD/openpsg_b200/$lib
  timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "self_attn" 2>&1 | grep -E "passed|failed" 
  timeout 600 python bench.py --steps 5 --no-llm --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), round(d['ms_per_step'],2), {k: round(v,3) for k,v in list(d['kernel_ms_per_step'].items())[:4]})"
done
