set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "xattn" 2>&1 | tail -4
timeout 600 python scripts/kbench.py xattn 2>&1 | tee gpurun_out/kbench_l.jsonl
timeout 900 python -m pytest tests/test_qformer_gpu.py -q 2>&1 | tail -3
