timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | cut -c1-260
