for a in 3 6 10; do echo "== a_stages $a"; OPSG_SKINNY_ASTAGES=$a timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep '"gemm_skinny"' | cut -c1-260 | sed -n '1p;3p;5p'; done
echo "== nostore"; OPSG_SKINNY_DBG=1 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep '"gemm_skinny"' | cut -c1-260 | sed -n '1p;3p;5p'
echo "== nofinalize"; OPSG_SKINNY_DBG=2 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep '"gemm_skinny"' | cut -c1-260 | sed -n '1p;3p;5p'
