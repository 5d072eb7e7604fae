# trace hooks compiled out of the CTA-pair GEMM and K5: same-box A/B of the cfg4 step; new tests (NaN top-k, K11 tiles)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "topk or mask_pool or exist or xattn or gemm" 2>&1 | grep -E "passed|failed|^E|Error" | head -10
for lib in libopsg_b200_withtrace.so libopsg_b200.so libopsg_b200_withtrace.so libopsg_b200.so; do
  export OPSG_B200_LIB=$PWD/openpsg_b200/$lib
  timeout 600 python bench.py --steps 5 --no-llm --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']
print('$lib', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'gemm', round(k['gemm_bf16'],2), 'xattn', round(k['xattn_pairs'],3), 'roof', round(d['roofline']['frac'],3), round(d['roofline_xattn']['frac'],3), 'maskpool', d.get('roofline_mask_pool') and round(d['roofline_mask_pool']['achieved']))"
done 2>&1 | tee gpurun_out/r2_trace_ab_x.log
