// Micro-benchmark (development aid): how fast can the SMs pull a bf16 weight matrix out of HBM through a TMA ring, as a
// function of the bytes per request (box rows), the ring depth and the number of producer lanes?  No MMA: one consumer
// warp waits for a stage and hands it straight back.  Answers the question behind csrc/gemm_skinny.cu's ~3.8 TB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o tma_stream tma_stream.cu && ./tma_stream
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../openpsg_b200/csrc/common.cuh"
using namespace opsg;

struct P {
  int box_rows;      // rows per TMA request (x 128 bytes)
  int stages;
  int n_boxes;       // requests per CTA
  int K;             // columns (elements) of the matrix
  int rows_total;
  long long unit0;   // first (row block, k block) unit of this launch
  int mode;          // 0: 2-D tensor TMA, 1: 1-D bulk copies of box_rows*128 contiguous bytes, 2: as 0 but two producer warps,
                     // 3: 2-D TMA in the small-M GEMM's traversal order (4 K slices x 37 CTAs, n-tiles g, g+37, ...; 10 K-blocks
                     //    per tile), three producer warps, 4: 1-D bulk (pre-tiled weights), three producer warps
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* base, const P p,
                                                        unsigned* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.box_rows * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + 64;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  // the matrix is walked as (row block, k block) pairs; CTA b takes pairs b, b + grid, ...
  const int kblocks = p.K / 64;
  const int nprod = p.mode >= 3 ? 3 : (p.mode == 2 ? 2 : 1);
  if (warp < nprod) {
    // incremental stage / phase / K-block counters: no integer divisions in the issue loop
    int stage = warp % p.stages;
    uint32_t phase = (warp / p.stages) & 1;
    int t10 = warp / 10, k10 = warp % 10;                     // mode 3: tile-in-CTA and K-block counters
    const int slice = blockIdx.x % 4, g = blockIdx.x / 4;
    const long long tile0 = (p.unit0 / (static_cast<long long>(p.n_boxes) * 148)) * 850 + g;
    long long unit = p.unit0 + static_cast<long long>(warp) * gridDim.x + blockIdx.x;
    int rb = static_cast<int>(unit / kblocks), kbm = static_cast<int>(unit % kblocks);      // modes 0 / 2
    const int d_rb = (nprod * static_cast<int>(gridDim.x)) / kblocks, d_kb = (nprod * static_cast<int>(gridDim.x)) % kblocks;
    for (int i = warp; i < p.n_boxes; i += nprod) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one_sync()) {
        mbar_expect_tx(&full[stage], stage_bytes);
        if (p.mode == 3) {
          tma_load_2d(smem + stage * stage_bytes, &tm, &full[stage], (slice * 10 + k10) * 64, static_cast<int>((tile0 + t10 * 37) * p.box_rows));
        } else if (p.mode == 1 || p.mode == 4) {
          bulk_load_1d(smem + stage * stage_bytes, base + unit * stage_bytes, stage_bytes, &full[stage]);
        } else {
          tma_load_2d(smem + stage * stage_bytes, &tm, &full[stage], kbm * 64, rb * p.box_rows);
        }
      }
      __syncwarp();
      stage += nprod;
      while (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
      k10 += nprod;
      while (k10 >= 10) { k10 -= 10; ++t10; }
      unit += static_cast<long long>(nprod) * gridDim.x;
      rb += d_rb; kbm += d_kb;
      if (kbm >= kblocks) { kbm -= kblocks; ++rb; }
    }
  } else if (warp == 3) {
    unsigned acc = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < p.n_boxes; ++i, ++stage) {
      if (stage == p.stages) { stage = 0; phase ^= 1; }
      mbar_wait(&full[stage], phase);
      acc += smem[stage * stage_bytes + (threadIdx.x & 31) * 4];
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage]);
    }
    if (acc == 0xffffffffu) *sink = acc;
  }
}

int main() {
  const int K = 2560;                   // OPT-2.7B hidden size: row pitch 5120 B
  const long long rows = 400000;        // 2.05 GB: far beyond L2
  uint8_t* w;
  cudaMalloc(&w, rows * K * 2);
  cudaMemset(w, 1, rows * K * 2);
  unsigned* sink;
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int box_rows_list[] = {64, 128, 256};
  for (int mode : {0, 2, 3})
    for (int br : box_rows_list)
      for (int smem_kb : {32, 64, 128, 192}) {
        P p;
        p.box_rows = br; p.K = K; p.rows_total = (int)rows; p.mode = mode;
        p.stages = smem_kb * 1024 / (br * 128);
        if (p.stages < 1 || p.stages > 64) continue;
        if (mode == 3 && br != 64) continue;
        if (p.stages < 3) continue;
        const long long total_boxes = rows / br * (K / 64);
        p.n_boxes = (int)(total_boxes / 148 / 8);          // 1/8 of the matrix per launch (~256 MB)
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)br};
        cuuint32_t es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); return 1; }
        const int smem_bytes = 1024 + p.stages * br * 128 + 1024;
        p.unit0 = 0;
        for (int it = 0; it < 2; ++it) stream_kernel<<<148, 128, smem_bytes>>>(tm, w, p, sink);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int it = 0; it < reps; ++it) { p.unit0 = (long long)(it + 2) * p.n_boxes * 148; stream_kernel<<<148, 128, smem_bytes>>>(tm, w, p, sink); }
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)p.n_boxes * 148 * br * 128 * reps;
        printf("mode %d (%s) box %3d rows (%2d KB) stages %2d (%3d KB in flight): %.2f TB/s\n", mode,
               mode == 0 ? "2-D TMA" : mode == 1 ? "1-D bulk" : mode == 2 ? "2-D TMA, 2 producer warps" : mode == 3 ? "2-D TMA, GEMM order, 3 warps" : "1-D bulk, 3 warps", br, br * 128 / 1024, p.stages,
               smem_kb, bytes / (ms * 1e-3) / 1e12);
      }
  return 0;
}
