# chained small-M GEMM: a decoder layer's GEMMs as 1 / 2 / 4 launches against one kernel per op
mkdir -p gpurun_out
timeout 600 python scripts/chain_layers_bench.py 2>&1 | tail -6 | tee gpurun_out/r2_ci_layers.log
