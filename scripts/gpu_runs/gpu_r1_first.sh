set -x
timeout 900 python -m pytest tests/test_qformer_gpu.py -q -x 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --images-per-step 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
