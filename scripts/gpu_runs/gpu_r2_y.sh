# K5: row sums from the softmax threads, PV MMA at N = 64: parity + kbench + ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qformer_gpu.py -x -q -k "xattn or qformer or stage or golden or cfg" 2>&1 | grep -E "passed|failed|^E|Error" | head -10
timeout 120 python scripts/kbench.py xattn 2>&1 | tail -7 | cut -c1-250 | tee gpurun_out/r2_kbench_xattn_y.log
export OPSG_CUDA_GRAPHS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xattn_pairs_kernel -s 6 -c 2 -o gpurun_out/r2_prof_xattn_y -f python bench.py --steps 1 --warmup 1 --total-images 1 --no-cpu-baseline --no-llm > gpurun_out/r2_ncu_xattn_y.log 2>&1
tail -2 gpurun_out/r2_ncu_xattn_y.log
