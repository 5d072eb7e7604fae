timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "small_m" 2>&1 | tail -5
timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | cut -c1-220
OPSG_SKINNY=0 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep streamk | cut -c1-220
