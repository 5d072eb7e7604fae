# clean (no trace hooks) small-M GEMM: producer-warp counts x constant-weight early streaming, same box as the round-start library
mkdir -p gpurun_out
export OPSG_B200_LIB=$PWD/openpsg_b200/libopsg_b200_orig.so
timeout 600 python scripts/llm_decode_ab.py 2>&1 | grep wait_for | sed "s/^/orig /"
unset OPSG_B200_LIB
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "small_m" 2>&1 | grep -E "passed|failed|^E|Error" | head -5
for cfg in "1 1 3" "3 1 3" "1 4 6" "3 4 6"; do
  set -- $cfg
  export OPSG_SKINNY_PRODUCERS=$1 OPSG_SKINNY_APRODUCERS=$2 OPSG_SKINNY_ASTAGES=$3
  timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep -v "tiled" | cut -c1-60,120-260 | sed "s/^/wp$1 ap$2 as$3 /"
  timeout 600 python scripts/llm_decode_ab.py 2>&1 | tail -2 | sed "s/^/wp$1 ap$2 as$3 /"
done 2>&1 | tee gpurun_out/r2_decode_ab_w.log
