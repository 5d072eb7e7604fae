# producer loops without integer divisions: TMA stream ubench again; small-M GEMM 1 vs 3 weight-producer warps vs the round-start library
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o /tmp/tma_stream scripts/ubench/tma_stream.cu && timeout 300 /tmp/tma_stream 2>&1 | grep -v "^opsg" | tee gpurun_out/r2_tma_stream_u.log
export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_orig.so
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep "gemm_skinny\"" | cut -c1-60,120-260 | sed "s/^/orig /"
timeout 600 python scripts/llm_decode_ab.py 2>&1 | grep wait_for | sed "s/^/orig /"
unset OPSG_B200_LIB
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "small_m" 2>&1 | grep -E "passed|failed|^E|Error" | head -5
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/new /"
timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -2 | sed "s/^/new /"
