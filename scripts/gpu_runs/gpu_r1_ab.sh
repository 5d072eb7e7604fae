set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-llm 2>&1 | tail -1 | tee gpurun_out/bench_ab.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}); print(d['e2e']['value']); print(d['roofline']['achieved'], d['kernel_ms_per_step'])"
BENCH1="python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline --no-llm"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 8 -c 8 -o gpurun_out/prof_gemm_ab $BENCH1 > gpurun_out/ncu_gemm_ab.log 2>&1
tail -2 gpurun_out/ncu_gemm_ab.log
