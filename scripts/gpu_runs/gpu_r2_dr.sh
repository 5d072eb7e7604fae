# experiment: contiguous item slices fetched by ONE 1-D bulk copy (heads = 1 emulates a head-major cache)
mkdir -p gpurun_out
for a in "800 65 32" "25600 65 1" "3200 65 1" "25600 81 1"; do python scripts/decode_attn_probe.py $a | tee -a gpurun_out/r2_dr_probe.log; done
