# round 2: K5 v4 (split MMA warps, one-thread-per-row softmax, epilogue warps, keys in object order): tests, bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_b.log
tail -4 gpurun_out/r2_tests_b.log
timeout 600 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
cut -c1-600 gpurun_out/r2_bench_b.json
export OPSG_CUDA_GRAPHS=0
BENCH1="python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline --no-llm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs_kernel -s 6 -c 2 -o gpurun_out/r2_prof_xattn_b -f $BENCH1 > gpurun_out/r2_ncu_xattn_b.log 2>&1
tail -3 gpurun_out/r2_ncu_xattn_b.log
