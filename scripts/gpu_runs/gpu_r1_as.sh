OPSG_GEMM2_EPIW=16 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -3
OPSG_GEMM2_EPIW=16 timeout 200 python scripts/gemm_trace.py gelu 2>&1 | tail -6
OPSG_GEMM2_EPIW=16 timeout 300 python scripts/kbench.py gemm --iters 10 2>&1 | cut -c1-200 | head -8
OPSG_GEMM2_EPIW=8 timeout 300 python scripts/kbench.py gemm --iters 10 2>&1 | cut -c1-200 | head -8
OPSG_GEMM2_EPIW=16 timeout 300 python bench.py --steps 10 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
OPSG_GEMM2_EPIW=8 timeout 300 python bench.py --steps 10 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
