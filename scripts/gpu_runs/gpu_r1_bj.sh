for v in 0 1; do
echo "== OPSG_GELU_TANH=$v"
OPSG_GELU_TANH=$v timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
OPSG_GELU_TANH=$v timeout 300 python scripts/kbench.py gemm --iters 10 2>&1 | grep gelu | cut -c1-200
OPSG_GELU_TANH=$v timeout 600 python -m pytest tests/test_qformer_gpu.py -m gpu -x -q -k "golden" 2>&1 | tail -2
OPSG_GELU_TANH=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-llm --no-cpu-baseline 2>&1 | tail -1 | cut -c1-160
done
