mkdir -p gpurun_out
echo "pair kernel (product)"; OPSG_PROBE_MEDIUM=1 python scripts/gemm_medium_m.py 800 2>&1 | tee gpurun_out/r2_dq_gemm.log
echo "1-CTA kernel for <= 200 pair tiles"; OPSG_EXP_PAIR_MAX_TILES=200 OPSG_PROBE_MEDIUM=1 python scripts/gemm_medium_m.py 800 400 2>&1 | tee -a gpurun_out/r2_dq_gemm.log
