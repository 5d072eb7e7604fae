# cluster-reduction small-M GEMM: what does the occupancy query return, and timing with the cluster count forced
mkdir -p gpurun_out
OPSG_SKINNY_DEBUG=1 timeout 300 python scripts/kbench.py streamk --iters 3 2>&1 | grep skinny_cluster | sort | uniq -c | tee gpurun_out/r2_cc_plan.log
for nc in 128 144 148; do
OPSG_SKINNY_NC=$nc timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/nc$nc /" | tee -a gpurun_out/r2_cc_kbench.log
done
