# chained small-M GEMM: timeline of one qkv-shaped and one lm_head-shaped launch (trace build)
mkdir -p gpurun_out
export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_trace.so
timeout 300 python scripts/chain_trace.py 7680 2560 2>&1 | tail -40 | tee gpurun_out/r2_cg_trace.log
timeout 300 python scripts/chain_trace.py 2560 10240 2>&1 | tail -40 | tee -a gpurun_out/r2_cg_trace.log
