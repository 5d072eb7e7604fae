mkdir -p gpurun_out
python scripts/e2e_cfg3_breakdown.py 2>&1 | grep -v Warning | tee gpurun_out/r2_ds_breakdown.log
