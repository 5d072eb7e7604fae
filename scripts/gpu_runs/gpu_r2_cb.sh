# cluster-reduction small-M GEMM, warp-aggregated remote arrives: tests, per-GEMM timing (auto tile width / forced 64), cfg3 decode
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "small_m" 2>&1 | grep -E "passed|failed|^E  |Error|timed out|^tests" | head -12 | tee gpurun_out/r2_cb_tests.log
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/cluster /" | tee gpurun_out/r2_cb_kbench.log
OPSG_SKINNY_BN=64 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/cluster64 /" | tee -a gpurun_out/r2_cb_kbench.log
timeout 300 python scripts/kbench.py layers --iters 10 2>&1 | grep -v tiled | sed "s/^/cluster /" | tee -a gpurun_out/r2_cb_kbench.log
OPSG_SKINNY_BN=64 timeout 300 python scripts/kbench.py layers --iters 10 2>&1 | grep -v tiled | sed "s/^/cluster64 /" | tee -a gpurun_out/r2_cb_kbench.log
timeout 600 python scripts/llm_decode_time.py 2>&1 | tail -1 | sed "s/^/cluster /" | tee gpurun_out/r2_cb_decode.log
OPSG_SKINNY_CLUSTER=0 timeout 600 python scripts/llm_decode_time.py 2>&1 | tail -1 | sed "s/^/workspace /" | tee -a gpurun_out/r2_cb_decode.log
