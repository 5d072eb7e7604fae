set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -8
timeout 600 python scripts/kbench.py gemm 2>&1 | tee gpurun_out/kbench_aa.jsonl | cut -c1-175
OPSG_GEMM_2CTA=0 timeout 600 python scripts/kbench.py gemm 2>&1 | cut -c1-175 | head -3
