# ring LayerNorm: tests again + ncu --set full of two launches inside a cfg2 image
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" 2>&1 | grep -E "passed|failed|^E  |Error" | head -12 | tee gpurun_out/r2_dl_tests.log
export OPSG_CUDA_GRAPHS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_ring -s 10 -c 3 -o gpurun_out/r2_prof_ln_ring2 -f python bench.py --steps 1 --warmup 1 --total-images 1 --no-cpu-baseline --no-llm > gpurun_out/r2_dl_ncu.log 2>&1
tail -2 gpurun_out/r2_dl_ncu.log
unset OPSG_CUDA_GRAPHS
timeout 900 python bench.py --no-llm --no-cpu-baseline > gpurun_out/r2_dl_bench.json 2> gpurun_out/r2_dl_bench.err
python - <<'P'
import json
d=json.loads([x for x in open('gpurun_out/r2_dl_bench.json') if x.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step']['layernorm_bf16'], d['roofline_hbm_kernels']['layernorm_bf16']['achieved'], d['roofline']['achieved'], d['clocks'], d['results']['sha1'])
P
