# GPU call C: evidence for GEMM v2 — ncu --set full of layer-0 GEMMs + launch list of one bench step
set -x
mkdir -p gpurun_out
BENCH1="python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 60 -c 8 -o gpurun_out/prof_gemm_c $BENCH1 > gpurun_out/ncu_gemm_c.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_c.csv $BENCH1 > gpurun_out/ncu_launches_c.log 2>&1
tail -2 gpurun_out/ncu_launches_c.log
