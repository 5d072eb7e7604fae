# stacked-batch parity at depth 32, memcheck over the new kernels / grouped path, smoke()
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_bench_sizes_gpu.py -x -q -s 2>&1 | grep -E "passed|failed|^E  |Error|parity|checksum" | head -20 | tee gpurun_out/r2_di_tests.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_batching_gpu.py -x -q -k "llm_attn or layernorm or stacked or forward_batch or selected_rows" > gpurun_out/r2_di_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_di_tests.log
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_di_memcheck.log | head -10 | tee -a gpurun_out/r2_di_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r2_di_tests.log
