# round-2 evidence: launch list of one eager cfg2 image, ncu --set full of K5 (cfg2 and cfg5 shapes), K4, the CTA-pair GEMM,
# and the HBM-bound kernels the north star names (LayerNorm, pair-mask bits, existence filter + top-k, mask pooling, decode
# attention), plus K11 timings
mkdir -p gpurun_out
export OPSG_CUDA_GRAPHS=0
BENCH1="python bench.py --steps 1 --warmup 1 --total-images 1 --no-cpu-baseline --no-llm"
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv $BENCH1 > gpurun_out/r2_ncu_launches.log 2>&1
timeout 600 $NCU -k regex:xattn_pairs_kernel -s 6 -c 2 -o gpurun_out/r2_prof_xattn -f $BENCH1 > gpurun_out/r2_ncu_xattn.log 2>&1
timeout 600 $NCU -k regex:self_attn_pairs_kernel -s 6 -c 2 -o gpurun_out/r2_prof_selfattn -f $BENCH1 > gpurun_out/r2_ncu_selfattn.log 2>&1
timeout 900 $NCU -k regex:gemm2_bf16_kernel -s 42 -c 8 -o gpurun_out/r2_prof_gemm -f $BENCH1 > gpurun_out/r2_ncu_gemm.log 2>&1
timeout 600 $NCU -k "regex:layernorm_bf16_kernel|pair_mask_bits_kernel|exist_logits_kernel|topk_rank_kernel|token_order_kernel|splitk_reduce|patch_im2col_kernel|qformer_embed_ln_kernel" -s 30 -c 16 -o gpurun_out/r2_prof_hbm -f $BENCH1 > gpurun_out/r2_ncu_hbm.log 2>&1
timeout 300 $NCU -k regex:mask_pool -s 4 -c 4 -o gpurun_out/r2_prof_maskpool -f python scripts/mask_pool_probe.py > gpurun_out/r2_ncu_maskpool.log 2>&1
timeout 600 $NCU -k "regex:decode_attn_smem_kernel|skinny|argmax_rows" --launch-skip 40 -c 8 -o gpurun_out/r2_prof_decode -f python scripts/llm_probe.py 4 > gpurun_out/r2_ncu_decode.log 2>&1
unset OPSG_CUDA_GRAPHS
timeout 120 python scripts/mask_pool_probe.py > gpurun_out/r2_mask_pool_probe.log 2>&1
cat gpurun_out/r2_mask_pool_probe.log
timeout 120 python scripts/kbench.py xattn 2>&1 | tail -7 > gpurun_out/r2_kbench_xattn.log
# K5 at the cfg5 shape (80 objects): one eager image of cfg5 through the head
OPSG_CUDA_GRAPHS=0 timeout 600 $NCU -k regex:xattn_pairs_kernel -s 2 -c 1 -o gpurun_out/r2_prof_xattn_cfg5 -f python scripts/cfg5_probe.py > gpurun_out/r2_ncu_xattn_cfg5.log 2>&1
ls -la gpurun_out/r2_prof_*.ncu-rep | awk '{print $5, $9}'
