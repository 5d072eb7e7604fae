# GPU call A (round 1): parity tests, smoke, bench, launch list, ncu --set full of the two top kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_a.json
BENCH1="python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs -s 6 -c 2 -o gpurun_out/prof_xattn_a $BENCH1 > gpurun_out/ncu_xattn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 60 -c 12 -o gpurun_out/prof_gemm_a $BENCH1 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out
