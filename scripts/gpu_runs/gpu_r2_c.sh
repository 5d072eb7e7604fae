# round 2: K5 v4 with one head per CTA; new K11; full-size parity tests; traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_bench_sizes_gpu.py tests/test_qformer_gpu.py -x -q -k "xattn or mask_pool or cfg3 or 80_objects or token_order" -s 2>&1 | grep -E "passed|failed|^E|cfg3|Error" | head -30
timeout 60 python scripts/xattn_trace.py 40 masks > gpurun_out/r2_xattn_trace_d.log 2>&1
head -14 gpurun_out/r2_xattn_trace_d.log | cut -c1-200; tail -4 gpurun_out/r2_xattn_trace_d.log
for v in 0 1 3; do OPSG_XATTN_VARIANT=$v timeout 120 python scripts/kbench.py xattn 2>&1 | grep object_order | sed "s/^/v$v /" ; done
