# GPU call I: xattn v3 (16 softmax warps, MMA-side mask bias) — parity, microbench, ncu
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "xattn" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_qformer_gpu.py -q 2>&1 | tail -6
timeout 600 python scripts/kbench.py xattn 2>&1 | tee gpurun_out/kbench_i.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs -s 4 -c 1 -o gpurun_out/prof_xattn_i python scripts/kbench.py xattn --iters 3 > gpurun_out/ncu_xattn_i.log 2>&1
tail -2 gpurun_out/ncu_xattn_i.log
