// Micro-benchmark (development aid): tcgen05.ld / tcgen05.st / MUFU.EX2 / FMNMX throughput per SM sub-partition on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../openpsg_b200/csrc/common.cuh"
using namespace opsg;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
    :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]) : "memory");
}

// mode 0: ld.x32 wait each; 1: ld.x32 x4 then wait; 2: ld.x16 wait each; 3: st.x16 (wait at end of 8); 4: ex2; 5: fmnmx; 6: ffma
__global__ void bench(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32]; uint32_t w[16];
  for (int j = 0; j < 16; ++j) w[j] = threadIdx.x + j;
  float acc = threadIdx.x * 1e-3f, acc2 = 1.f;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) { tmem_ld32(base + (i & 7) * 32, v); tmem_ld_wait(); acc += __uint_as_float(v[i & 31]); }
  } else if (mode == 1) {
    for (int i = 0; i < iters; i += 4) {
      uint32_t v1[32], v2[32], v3[32];
      tmem_ld32(base + 0, v); tmem_ld32(base + 32, v1); tmem_ld32(base + 64, v2); tmem_ld32(base + 96, v3); tmem_ld_wait();
      acc += __uint_as_float(v[i & 31] ^ v1[i & 31] ^ v2[i & 31] ^ v3[i & 31]);
    }
  } else if (mode == 2) {
    for (int i = 0; i < iters; ++i) { tmem_ld16(base + (i & 15) * 16, w); tmem_ld_wait(); acc += __uint_as_float(w[i & 15]); }
  } else if (mode == 3) {
    for (int i = 0; i < iters; ++i) { tmem_st16(base + (i & 15) * 16, w); if ((i & 7) == 7) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else if (mode == 4) {
    float x[8]; for (int j = 0; j < 8; ++j) x[j] = acc + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
    }
    for (int j = 0; j < 8; ++j) acc += x[j];
  } else if (mode == 5) {
    float x[8]; for (int j = 0; j < 8; ++j) x[j] = acc + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[j]) : "f"(acc2));
      acc2 += 1.f;
    }
    for (int j = 0; j < 8; ++j) acc += x[j];
  } else if (mode == 6) {
    float x[8]; for (int j = 0; j < 8; ++j) x[j] = acc + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(acc2), "f"(acc));
    }
    for (int j = 0; j < 8; ++j) acc += x[j];
  } else if (mode == 7) {   // 3-input max (sm_100)
    float x[8]; for (int j = 0; j < 8; ++j) x[j] = acc + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[j]) : "f"(acc2), "f"(acc));
      acc2 += 1.f;
    }
    for (int j = 0; j < 8; ++j) acc += x[j];
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  const char* names[] = {"ld.x32 (wait each)", "ld.x32 x4 then wait", "ld.x16 (wait each)", "st.x16", "ex2 (x8 per iter)", "max2 (x8)", "ffma (x8)", "max3 (x8)"};
  for (int threads : {128, 256, 384}) {
    for (int mode = 0; mode < 8; ++mode) {
      const int iters = 4096;
      bench<<<148, threads>>>(mode, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d threads %d: %s\n", mode, threads, cudaGetErrorString(e)); return 1; }
      long long h[16]; cudaMemcpy(h, out, sizeof(long long) * (threads / 32), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < threads / 32; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per = (double)mx / iters;
      const int ops = (mode >= 4) ? 8 : 1;
      printf("threads %3d  %-22s : %8.2f cycles per warp-iteration (%d op/iter) -> %.2f cycles per warp-op, warps/SMSP=%d\n",
             threads, names[mode], per, ops, per / ops, threads / 128);
    }
  }
  return 0;
}
