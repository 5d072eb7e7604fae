mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-llm 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'])"; }
run A=1
run OPSG_SELF_ATTN_PIPELINED=0
run OPSG_GEMM_2CTA=0
run A=2
