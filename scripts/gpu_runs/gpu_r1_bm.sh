for m in 0 3 2 4; do
echo "== OPSG_SKINNY_LSU_MOD=$m"
OPSG_SKINNY_LSU_MOD=$m timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "small_m" 2>&1 | tail -1
OPSG_SKINNY_LSU_MOD=$m timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep '"gemm_skinny"' | cut -c1-230
done
