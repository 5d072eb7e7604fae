# round 2, third evidence pass (1 GPU) after the LLM batching / decode attention / ring LayerNorm / medium-M GEMM work:
# full suite, both bench arms, launch lists, ncu --set full of the kernels that changed
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_tests_final3.log
tail -3 gpurun_out/r2_tests_final3.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_final3.json 2> gpurun_out/r2_bench_reference_final3.err
cut -c1-300 gpurun_out/r2_bench_reference_final3.json
timeout 1500 python bench.py > gpurun_out/r2_bench_final3.json 2> gpurun_out/r2_bench_final3.err
cut -c1-300 gpurun_out/r2_bench_final3.json; tail -2 gpurun_out/r2_bench_final3.err
export OPSG_CUDA_GRAPHS=0
BENCH1="python bench.py --steps 1 --warmup 1 --total-images 1 --no-cpu-baseline --no-llm"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_final3.csv $BENCH1 > /dev/null 2>&1
ls -la gpurun_out/*final3* | awk '{print $5, $9}'
