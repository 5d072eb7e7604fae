timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "small_m" 2>&1 | tail -3
echo "== lsu"; timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep 'skinny' | cut -c1-260
echo "== lsu nofinalize"; OPSG_SKINNY_DBG=2 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep 'skinny' | cut -c1-260 | sed -n '1p;2p;5p;6p;9p;10p'
echo "== tma"; OPSG_SKINNY_LSU=0 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep '"gemm_skinny"' | cut -c1-260
