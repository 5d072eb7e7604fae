# round 2 (re-entry): full GPU suite, default bench, then the evidence pass (gpu_r2_prof.sh)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests_i.log
tail -4 gpurun_out/r2_tests_i.log
timeout 900 python bench.py > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err
cut -c1-1500 gpurun_out/r2_bench_i.json
tail -3 gpurun_out/r2_bench_i.err
bash scripts/gpu_runs/gpu_r2_prof.sh 2>&1 | tail -30
