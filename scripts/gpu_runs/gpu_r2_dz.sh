# launch list of a stacked (800 sequences) prefill + 2 decode steps, eager: per-kernel durations of a decode step at M = 800
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 1400 --csv --log-file gpurun_out/r2_launches_llm_stacked.csv python scripts/llm_probe.py 3 8 > gpurun_out/r2_dz.log 2>&1
tail -2 gpurun_out/r2_dz.log; wc -l gpurun_out/r2_launches_llm_stacked.csv
