# same-box A/B: library of commit f4245ae (round-2 start) vs the current one: kbench small-M GEMMs and the cfg3 decode leg
mkdir -p gpurun_out
for lib in libopsg_b200_orig.so libopsg_b200.so libopsg_b200_orig.so libopsg_b200.so; do
  export OPSG_B200_LIB=$PWD/openpsg_b200/$lib
  timeout 300 python scripts/kbench.py streamk --iters 10 2>&1 | grep "gemm_skinny\"" | grep -v lm_head | cut -c1-60,120-260 | sed "s/^/$lib /"
  timeout 600 python scripts/llm_decode_ab.py 2>&1 | grep wait_for | sed "s/^/$lib /"
done 2>&1 | tee gpurun_out/r2_decode_ab_t.log
