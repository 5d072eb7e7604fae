# round 2: tcgen05 prefill attention parity + decode L2-prefetch A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_llm_gpu.py tests/test_llama_gpu.py -x -q -k "llm_attn or prefill or llm_decode or llama" 2>&1 | grep -E "passed|failed|^E|Error" | head -20
timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -3
