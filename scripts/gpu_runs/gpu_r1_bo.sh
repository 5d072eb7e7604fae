for v in 0 1 0 1 0 1; do OPSG_GELU_TANH=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-llm --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gelu_tanh',$v,'value',round(d['value']),'ms',round(d['ms_per_step'],3),'gemm TF/s',round(d['roofline']['achieved']),'clk',d['clocks']['sm_mhz'],'W',d['clocks'].get('power_w_max'))"; done
