# round 2: small-M GEMM with three TMA producer warps + 208 KB of shared memory (room for the neighbours): parity, kbench, decode A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_llm_gpu.py -x -q -k "small_m or llm_decode or patch or determin" 2>&1 | grep -E "passed|failed|^E|Error" | head -20
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v tiled | tee gpurun_out/r2_kbench_streamk_k.log
timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -3 | tee gpurun_out/r2_llm_decode_ab_k.log
