set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs_kernel -s 9 -c 1 -o gpurun_out/prof_xattn_k python scripts/kbench.py xattn --iters 3 > gpurun_out/ncu_xattn_k.log 2>&1
tail -2 gpurun_out/ncu_xattn_k.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_k.json
