# TMA-staged decode attention: tests, A/B probe against the cp.async staging
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "llm_attn" 2>&1 | grep -E "passed|failed|^E  |Error" | head -20 | tee gpurun_out/r2_dd_tests.log
for a in "800 65" "800 81" "100 65" "400 65"; do
  python scripts/decode_attn_probe.py $a | tee -a gpurun_out/r2_dd_probe.log
  OPSG_DECODE_ATTN_TMA=0 python scripts/decode_attn_probe.py $a | sed 's/^/cp.async: /' | tee -a gpurun_out/r2_dd_probe.log
done
timeout 900 python -m pytest tests/test_llm_gpu.py tests/test_llama_gpu.py tests/test_batching_gpu.py -x -q 2>&1 | grep -E "passed|failed|^E  |Error" | head -20 | tee -a gpurun_out/r2_dd_tests.log
