# round-1 evidence after the epilogue rework / small-M GEMM: launch list of one eager bench step, ncu --set full of the
# CTA-pair GEMM (three epilogue kinds), the small-M GEMM and its finalize, and the launch list of decode steps
mkdir -p gpurun_out
export OPSG_CUDA_GRAPHS=0
BENCH1="python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline --no-llm"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final3.csv $BENCH1 > gpurun_out/ncu_launches_final3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 42 -c 8 -o gpurun_out/prof_gemm_final3 -f $BENCH1 > gpurun_out/ncu_gemm_final3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs_kernel -s 6 -c 1 -o gpurun_out/prof_xattn_final3 -f $BENCH1 > gpurun_out/ncu_xattn_final3.log 2>&1
for c in fc1 qkv; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny --launch-skip 6 -c 2 -o gpurun_out/prof_skinny_$c -f python scripts/skinny_one.py $c > gpurun_out/ncu_skinny_$c.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 800 --csv --log-file gpurun_out/llm_launches_final3.csv python scripts/llm_probe.py 5 > gpurun_out/llm_probe2.log 2>&1
tail -1 gpurun_out/llm_probe2.log
ls -la gpurun_out/*final3* gpurun_out/prof_skinny*
