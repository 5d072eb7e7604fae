for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "test_xattn_pairs" 2>&1 | grep -E "^E  |passed|failed" | head -6
done
