# K11 with two lane-private label accumulators: tests + A/B (no register cap / two CTAs per SM) + ncu of the accumulate kernel
mkdir -p gpurun_out
for v in nocap cap2; do
  export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_$v.so
  echo "== $v" | tee -a gpurun_out/r2_dw.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "mask_pool" 2>&1 | grep -E "passed|failed|^E  |Error" | head -5 | tee -a gpurun_out/r2_dw.log
  python scripts/mask_pool_probe.py 2>&1 | tail -3 | tee -a gpurun_out/r2_dw.log
done
export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_nocap.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_pool_accum -s 2 -c 2 -o gpurun_out/r2_prof_maskpool2 -f python scripts/mask_pool_probe.py > /dev/null 2>&1
ls -la gpurun_out/r2_prof_maskpool2* | awk '{print $5, $9}'
