mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o /tmp/tma_stream scripts/ubench/tma_stream.cu && timeout 300 /tmp/tma_stream | tee gpurun_out/r2_tma_stream_q.log
