# small-M GEMM back to one producer warp each (224 threads); shared-memory limit 208 KB vs 227 KB, w_const on/off
mkdir -p gpurun_out
for lim in 212992 232448; do
  export OPSG_SKINNY_SMEM=$lim
  timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled\|lm_head" | cut -c1-60,120-260 | sed "s/^/smem$lim /"
  timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -2 | sed "s/^/smem$lim /"
done 2>&1 | tee gpurun_out/r2_decode_ab_s.log
