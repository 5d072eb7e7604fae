# round 2: timeline of the decode GEMM chain (globaltimer stamps), with / without early weight streaming
mkdir -p gpurun_out
timeout 300 python scripts/skinny_trace.py 2>&1 | tee gpurun_out/r2_skinny_trace_n.log
