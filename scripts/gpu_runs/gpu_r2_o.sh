# round 2: small-M GEMM with four activation-producer warps + 6-deep activation ring: parity, timeline, decode A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_llm_gpu.py -x -q -k "small_m or llm_decode or patch or determin" 2>&1 | grep -E "passed|failed|^E|Error" | head -20
timeout 300 python scripts/skinny_trace.py 2>&1 | tee gpurun_out/r2_skinny_trace_o.log | cut -c1-200
timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -3 | tee gpurun_out/r2_llm_decode_ab_o.log
