# round 2: decode GEMMs stream constant weights ahead of the dependency (OPSG_GEMM_W_CONST) -- parity, A/B; TMA stream ubench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_llm_gpu.py tests/test_llama_gpu.py -x -q -k "small_m or llm_decode or llama or opt" 2>&1 | grep -E "passed|failed|^E|Error" | head -20
timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -3 | tee gpurun_out/r2_llm_decode_ab_j.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o /tmp/tma_stream scripts/ubench/tma_stream.cu && timeout 300 /tmp/tma_stream | tee gpurun_out/r2_tma_stream.log
