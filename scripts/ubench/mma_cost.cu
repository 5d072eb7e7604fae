// Micro-benchmark (development aid): cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and of the A
// operand source (shared memory descriptor vs TMEM), measured on one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_cost mma_cost.cu && ./mma_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../openpsg_b200/csrc/common.cuh"
using namespace opsg;

template <bool ELECT>
__global__ void __launch_bounds__(128, 1) bench(int N, int ts_mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  bool issuer;
  if (ELECT) issuer = (threadIdx.x < 32) && elect_one_sync();   // warp-uniform branch + elected lane
  else issuer = threadIdx.x == 0;
  if (issuer) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
    uint32_t phase = 0;
    // warm-up
    for (int k = 0; k < 16; ++k) {
      if (ts_mode) umma_ts(tb + 256, tb + (k & 7) * 8, umma_desc_k_sw128(b + (k & 3) * 32), idesc, k > 0);
      else umma_ss(tb + 256, umma_desc_k_sw128(a + (k & 3) * 32), umma_desc_k_sw128(b + (k & 3) * 32), idesc, k > 0);
    }
    tc_commit(&bar); mbar_wait(&bar, phase); phase ^= 1;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (ts_mode) umma_ts(tb + 256, tb + (k & 7) * 8, umma_desc_k_sw128(b + (k & 3) * 32), idesc, k > 0);
        else umma_ss(tb + 256, umma_desc_k_sw128(a + (k & 3) * 32), umma_desc_k_sw128(b + (k & 3) * 32), idesc, k > 0);
      }
    }
    tc_commit(&bar); mbar_wait(&bar, phase);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 2048);
  cudaFuncSetAttribute(bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 2048);
  const int reps = 64;
  for (int el = 0; el < 2; ++el)
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 64, 80, 128, 256}) {
      if (el) bench<true><<<148, 128, 16384 + 32768 + 2048>>>(N, ts, reps, out);
      else bench<false><<<148, 128, 16384 + 32768 + 2048>>>(N, ts, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d ts=%d: %s\n", N, ts, cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%s A from %s  N=%3d : %7.1f cycles per MMA (128 x N x 16), floor model 128*N/256 = %5.1f\n", el ? "elect.sync   " : "threadIdx==0 ", ts ? "TMEM" : "smem", N,
             (double)mx / (reps * 16), 128.0 * N / 256.0);
    }
  return 0;
}
