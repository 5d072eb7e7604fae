set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_r.json | cut -c1-3000
