mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  OPSG_SKINNY_DEBUG=$dbg timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep "gemm_skinny\"" | cut -c1-60,120-260 | sed "s/^/dbg$dbg /"
done 2>&1 | tee gpurun_out/r2_skinny_debug_r.log
