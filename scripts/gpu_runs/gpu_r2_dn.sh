mkdir -p gpurun_out
python scripts/gemm_medium_m.py 800 768 1024 512 400 2>&1 | tee gpurun_out/r2_dn_gemm.log
