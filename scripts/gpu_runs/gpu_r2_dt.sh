# decode attention: one-item-ahead prefetch of the small loads + paired score tiles: tests + probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "llm_attn" 2>&1 | grep -E "passed|failed|^E  |Error" | head -20 | tee gpurun_out/r2_dt_tests.log
for a in "800 65" "800 81" "100 65" "400 65"; do python scripts/decode_attn_probe.py $a | tee -a gpurun_out/r2_dt_probe.log; done
timeout 900 python -m pytest tests/test_llm_gpu.py tests/test_llama_gpu.py tests/test_batching_gpu.py -x -q 2>&1 | grep -E "passed|failed|^E  |Error" | head -20 | tee -a gpurun_out/r2_dt_tests.log
