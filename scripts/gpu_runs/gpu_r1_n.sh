timeout 300 python scripts/xattn_trace.py 80 2>&1 | tail -45
