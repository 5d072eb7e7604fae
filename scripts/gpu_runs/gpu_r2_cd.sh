# cluster-reduction small-M GEMM: where do 6 us per tile go?  (exchange off / epilogue memory off / 8 KB requests)
mkdir -p gpurun_out
for cfg in "1 0" "3 0" "0 1" "3 1"; do
set -- $cfg
if [ "$2" = "1" ]; then export OPSG_SKINNY_KREQ=1; else unset OPSG_SKINNY_KREQ; fi
OPSG_SKINNY_DBG=$1 timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/dbg$1 kreq1=$2 /" | tee -a gpurun_out/r2_cd_kbench.log
done
