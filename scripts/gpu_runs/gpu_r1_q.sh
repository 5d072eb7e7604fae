set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python scripts/kbench.py gemm xattn 2>&1 | tee gpurun_out/kbench_q.jsonl | cut -c1-160
timeout 300 python scripts/xattn_trace.py 80 2>&1 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_q.json | cut -c1-400
