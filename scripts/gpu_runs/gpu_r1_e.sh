# GPU call E: xattn v2 (double-buffered TMEM S, P in TMEM, ones-row sum) — parity under both sum modes + microbench
set -x
mkdir -p gpurun_out
OPSG_XATTN_FLAGS=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "xattn" 2>&1 | tail -6
OPSG_XATTN_FLAGS=0 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "xattn" 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python scripts/kbench.py xattn 2>&1 | tee gpurun_out/kbench_e.jsonl
OPSG_XATTN_IMPL=1 timeout 600 python scripts/kbench.py xattn 2>&1 | tail -1
