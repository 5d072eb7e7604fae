# GPU call J: xattn v3b (bias tiles prebuilt, no-swizzle descriptors: try both LBO/SBO orders)
set -x
mkdir -p gpurun_out
OPSG_XATTN_DESC_SWAP=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "xattn" 2>&1 | tail -8
OPSG_XATTN_DESC_SWAP=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "xattn" 2>&1 | tail -8
timeout 600 python scripts/kbench.py xattn 2>&1 | tee gpurun_out/kbench_j.jsonl
timeout 900 python -m pytest tests/test_qformer_gpu.py -q --durations=4 2>&1 | tail -12
