# does a head-major KV cache (every item's K / V slice contiguous) fix the decode attention's 3.7 TB/s? emulate with heads = 1
mkdir -p gpurun_out
for a in "800 65 32" "25600 65 1" "3200 65 1" "25600 81 1"; do
  python scripts/decode_attn_probe.py $a | tee -a gpurun_out/r2_de_probe.log
done
