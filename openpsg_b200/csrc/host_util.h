// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/opsg_b200.h"

namespace opsg {

int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = {box_cols, box_rows};
// 128-byte swizzle (box_cols * 2 bytes must be <= 128).  Out-of-bounds elements read as zero.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);
// the same without shared-memory swizzling: the box lands as dense rows of box_cols elements
int make_tmap_bf16_2d_plain(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

// LayerNorm-folding arguments of the CTA-pair GEMM (opsg_gemm_bf16_ln); any group may be null
struct GemmLnFold {
  const float* a_stats;    // [M, 2] (sum, sumsq) of the un-normalised A rows, or null
  const float* a_colsum;   // [N]
  const float* r_stats;    // [M, 2] of the un-normalised residual rows, or null
  const float* r_gamma;    // [N]
  const float* r_beta;     // [N]
  float* stats_out;        // [M, 2] accumulated (sum, sumsq) of the output rows, or null
  float eps;
};

// Slot of the current device in a per-device flag array: function attributes (cudaFuncSetAttribute) are per device, so the
// "already configured" flags of the launchers are kept per device ordinal.
inline int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
  return d;
}

#define OPSG_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) return ::opsg::set_error(OPSG_E_INVALID, __VA_ARGS__); \
  } while (0)

#define OPSG_CHECK_LAUNCH(what)                                     \
  do {                                                              \
    int _rc = ::opsg::check_cuda(cudaGetLastError(), what);         \
    if (_rc) return _rc;                                            \
  } while (0)

// Every kernel goes out through here.  With programmatic dependent launch (default; OPSG_PDL=0 turns it off) kernel
// N+1 may be scheduled as soon as every CTA of kernel N has executed griddepcontrol.launch_dependents, and blocks in
// griddepcontrol.wait until kernel N has completed and flushed: its launch latency, CTA rasterisation and prologue
// (barrier init, TMEM allocation, descriptor prefetch) overlap kernel N instead of following it.  Every kernel
// therefore calls pdl_wait() (common.cuh) before its first global-memory access.  Captured into CUDA graphs the
// attribute becomes a programmatic dependency edge.
bool pdl_enabled();
inline dim3 to_dim3(dim3 d) { return d; }
inline dim3 to_dim3(long long v) { return dim3(static_cast<unsigned>(v)); }
template <typename... KArgs, typename G, typename B, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), G grid, B block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = to_dim3(grid);
  cfg.blockDim = to_dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // the error is picked up by OPSG_CHECK_LAUNCH
}

}  // namespace opsg
