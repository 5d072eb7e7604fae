set -x
timeout 900 python -m pytest tests/test_qformer_gpu.py -q -x --durations=5 2>&1 | grep -v Warning | tail -40
OPSG_XATTN_IMPL=1 timeout 900 python -m pytest tests/test_qformer_gpu.py -q -x 2>&1 | tail -3
