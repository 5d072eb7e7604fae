set -x
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do
  echo "variant $v"
  OPSG_XATTN_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "test_xattn_pairs" 2>&1 | tail -1
  OPSG_XATTN_VARIANT=$v timeout 300 python scripts/kbench.py xattn --iters 10 2>&1 | cut -c1-140
done
