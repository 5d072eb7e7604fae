# GPU call D: LLM decode engine parity (a9-a10)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_llm_gpu.py -q -x -s 2>&1 | grep -v Warning | tail -25
