# online-softmax K5 for more than 256 image tokens + ring LayerNorm specialisation: tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "xattn or layernorm" 2>&1 | grep -E "passed|failed|^E  |Error" | head -12 | tee gpurun_out/r2_dv_tests.log
timeout 900 python -m pytest tests/test_qformer_gpu.py -x -q -s -k "more_than_256 or golden or 80_objects" 2>&1 | grep -E "passed|failed|^E  |Error|320 tokens" | head -12 | tee -a gpurun_out/r2_dv_tests.log
