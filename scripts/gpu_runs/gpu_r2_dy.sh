# memcheck over the code paths added after gpu_r2_di.sh: K5 beyond 256 tokens, medium-M split-K, ring LayerNorm, reworked decode attention
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_qformer_gpu.py -x -q -k "more_than_256 or medium_m or layernorm or llm_attn or patch" > gpurun_out/r2_dy_memcheck.log 2>&1
echo "memcheck rc=$?" | tee gpurun_out/r2_dy.log
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r2_dy_memcheck.log | head -10 | tee -a gpurun_out/r2_dy.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -x -q -k "more_than_256 or llm_attn_decode_long or layernorm" > gpurun_out/r2_dy_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2_dy.log
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r2_dy_racecheck.log | head -10 | tee -a gpurun_out/r2_dy.log
