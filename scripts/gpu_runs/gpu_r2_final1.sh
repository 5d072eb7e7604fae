# round 2, final evidence pass (1 GPU): full suite, both bench arms, launch list + ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_tests_final.log
tail -3 gpurun_out/r2_tests_final.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err
cut -c1-400 gpurun_out/r2_bench_reference_final.json
timeout 1200 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
cut -c1-300 gpurun_out/r2_bench_final.json; tail -2 gpurun_out/r2_bench_final.err
export OPSG_CUDA_GRAPHS=0
BENCH1="python bench.py --steps 1 --warmup 1 --total-images 1 --no-cpu-baseline --no-llm"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_final.csv $BENCH1 > /dev/null 2>&1
timeout 600 $NCU -k regex:xattn_pairs_kernel -s 6 -c 2 -o gpurun_out/r2_prof_xattn_final -f $BENCH1 > /dev/null 2>&1
timeout 900 $NCU -k regex:gemm2_bf16_kernel -s 42 -c 8 -o gpurun_out/r2_prof_gemm_final -f $BENCH1 > /dev/null 2>&1
timeout 600 $NCU -k "regex:decode_attn_smem_kernel|argmax_rows|layernorm_row_cta" --launch-skip 60 -c 6 -o gpurun_out/r2_prof_decode_small_final -f python scripts/llm_probe.py 4 > /dev/null 2>&1
timeout 600 $NCU -k "regex:llm_prefill_attn" -c 2 -o gpurun_out/r2_prof_prefill_final -f python scripts/llm_probe.py 2 > /dev/null 2>&1
ls -la gpurun_out/*final* | awk '{print $5, $9}'
