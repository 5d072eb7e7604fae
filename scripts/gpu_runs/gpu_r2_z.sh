# round 2: full GPU suite after pruning the experiment switches (7 left, each covered by a test), K11 tiles, trace hooks out
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_tests_z.log
tail -5 gpurun_out/r2_tests_z.log
timeout 300 ncu --set full --clock-control none -k regex:mask_pool_accum -s 2 -c 2 -o gpurun_out/r2_prof_maskpool_z -f python scripts/mask_pool_probe.py > gpurun_out/r2_ncu_maskpool_z.log 2>&1
tail -2 gpurun_out/r2_ncu_maskpool_z.log
