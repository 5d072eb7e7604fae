# K11: deferred reduction with 8 channels per CTA (two CTAs per SM): tests + probe + ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "mask_pool" 2>&1 | grep -E "passed|failed|^E  |Error" | head -5 | tee gpurun_out/r2_dx.log
python scripts/mask_pool_probe.py 2>&1 | tail -3 | tee -a gpurun_out/r2_dx.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_pool -s 4 -c 4 -o gpurun_out/r2_prof_maskpool3 -f python scripts/mask_pool_probe.py > /dev/null 2>&1
ls -la gpurun_out/r2_prof_maskpool3* | awk '{print $5, $9}'
