# decode attention probe + ncu --set full of one launch
mkdir -p gpurun_out
python scripts/decode_attn_probe.py 800 65 | tee gpurun_out/r2_dc_probe.log
python scripts/decode_attn_probe.py 800 81 | tee -a gpurun_out/r2_dc_probe.log
python scripts/decode_attn_probe.py 100 65 | tee -a gpurun_out/r2_dc_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn -s 4 -c 2 -o gpurun_out/r2_prof_decode_attn -f python scripts/decode_attn_probe.py 800 65 > gpurun_out/r2_dc_ncu.log 2>&1
tail -2 gpurun_out/r2_dc_ncu.log
