timeout 200 python scripts/gemm_trace.py plain 2>&1 | tail -44
timeout 200 python scripts/gemm_trace.py res 2>&1 | tail -22
