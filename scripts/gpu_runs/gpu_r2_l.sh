# round 2: small-M GEMM: capped (all-resident) reduction grid so the next GEMM can stream ahead; 1 vs 3 producer warps; K11 2-D tiles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_llm_gpu.py -x -q -k "small_m or llm_decode or mask_pool" 2>&1 | grep -E "passed|failed|^E|Error" | head -20
for wp in 1 3; do
  OPSG_SKINNY_PRODUCERS=$wp timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled\|lm_head" | cut -c1-60,120-260 | sed "s/^/wp$wp /"
  OPSG_SKINNY_PRODUCERS=$wp timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -3 | sed "s/^/wp$wp /"
done 2>&1 | tee gpurun_out/r2_decode_ab_l.log
timeout 120 python scripts/mask_pool_probe.py 2>&1 | tee gpurun_out/r2_mask_pool_probe_l.log
