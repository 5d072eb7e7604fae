timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -5
timeout 200 python scripts/gemm_trace.py plain 2>&1 | tail -6
timeout 200 python scripts/gemm_trace.py res 2>&1 | tail -6
timeout 200 python scripts/gemm_trace.py gelu 2>&1 | tail -6
timeout 300 python scripts/kbench.py gemm --iters 10 2>&1 | cut -c1-200
OPSG_GEMM2_BIAS_MMA=0 timeout 300 python scripts/kbench.py gemm --iters 10 2>&1 | cut -c1-200 | head -6
OPSG_FOLD_LN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
OPSG_FOLD_LN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
