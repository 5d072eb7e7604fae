# bisect of the small-M GEMM slowdown: orig / current / original producer loop / no trace+debug code / both (+ full shared memory)
mkdir -p gpurun_out
for lib in orig v1 v2 v3 v3big cur; do
  unset OPSG_SKINNY_SMEM
  case $lib in
    cur) export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200.so ;;
    v3big) export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_v3.so; export OPSG_SKINNY_SMEM=232448 ;;
    *) export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_$lib.so ;;
  esac
  timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep "gemm_skinny\"" | cut -c1-60,120-260 | sed "s/^/$lib /"
done 2>&1 | tee gpurun_out/r2_skinny_bisect_v.log
