set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs -s 4 -c 1 -o gpurun_out/prof_xattn_h python scripts/kbench.py xattn --iters 3 > gpurun_out/ncu_xattn_h.log 2>&1
tail -3 gpurun_out/ncu_xattn_h.log
