set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -3
timeout 600 python scripts/kbench.py gemm 2>&1 | cut -c1-175
timeout 900 python -m pytest tests/test_qformer_gpu.py -q 2>&1 | tail -2
