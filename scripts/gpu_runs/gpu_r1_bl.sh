for v in 0 1 0 1; do OPSG_FOLD_LN=$v timeout 300 python bench.py --steps 16 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fold',$v,'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'clk',d['clocks']['sm_mhz'],'W',d['clocks'].get('power_w_max'))"; done
OPSG_FOLD_LN=1 timeout 900 python -m pytest tests/test_qformer_gpu.py -m gpu -q 2>&1 | tail -3
