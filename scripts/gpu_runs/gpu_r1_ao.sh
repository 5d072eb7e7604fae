for c in qkv res gelu; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2 --launch-skip 3 -c 1 -o gpurun_out/g2_$c -f python scripts/gemm_one.py $c > gpurun_out/g2_$c.log 2>&1
tail -2 gpurun_out/g2_$c.log
done
