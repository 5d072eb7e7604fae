# chained small-M GEMM kernel (in-kernel reduction, phases, folded norms): tests, per-GEMM timing, layer chain, cfg3 decode
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "small_m or chain or row_stats" 2>&1 | grep -E "passed|failed|^E  |Error|timed out|^tests|opsg:" | head -20 | tee gpurun_out/r2_cf_tests.log
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/chain1 /" | tee gpurun_out/r2_cf_kbench.log
timeout 300 python scripts/kbench.py layers --iters 10 2>&1 | grep -v tiled | sed "s/^/chain1 /" | tee -a gpurun_out/r2_cf_kbench.log
timeout 600 python scripts/llm_decode_time.py 2>&1 | tail -1 | sed "s/^/chain1 /" | tee gpurun_out/r2_cf_decode.log
