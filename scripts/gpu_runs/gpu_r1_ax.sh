timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny --launch-skip 6 -c 4 -o gpurun_out/skinny_fc1 -f python scripts/skinny_one.py fc1 > gpurun_out/skinny_fc1.log 2>&1
tail -2 gpurun_out/skinny_fc1.log
