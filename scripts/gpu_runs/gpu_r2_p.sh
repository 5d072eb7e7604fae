mkdir -p gpurun_out
timeout 300 python scripts/skinny_trace.py 2>&1 | tee gpurun_out/r2_skinny_trace_p.log | cut -c1-220
