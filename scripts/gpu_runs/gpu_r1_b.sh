# GPU call B: new GEMM epilogue (TMA store, 8 epilogue warps) — parity + kernel micro-bench + bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python scripts/kbench.py gemm xattn 2>&1 | tee gpurun_out/kbench_b.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b.json
