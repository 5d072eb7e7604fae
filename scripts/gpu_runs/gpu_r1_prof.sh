# final round-1 evidence: launch list of one eager bench step + ncu --set full of the tensor-core kernels
set -x
mkdir -p gpurun_out
export OPSG_CUDA_GRAPHS=0
BENCH1="python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline --no-llm"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv $BENCH1 > gpurun_out/ncu_launches_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs_kernel -s 6 -c 2 -o gpurun_out/prof_xattn_final $BENCH1 > gpurun_out/ncu_xattn_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 42 -c 8 -o gpurun_out/prof_gemm_final $BENCH1 > gpurun_out/ncu_gemm_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qformer_self_attn_kernel -s 6 -c 2 -o gpurun_out/prof_selfattn_final $BENCH1 > gpurun_out/ncu_selfattn_final.log 2>&1
tail -1 gpurun_out/ncu_selfattn_final.log
