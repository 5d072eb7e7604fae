set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1400 --csv --log-file gpurun_out/llm_launches.csv python scripts/llm_probe.py 4 > gpurun_out/llm_probe.log 2>&1
tail -2 gpurun_out/llm_probe.log
