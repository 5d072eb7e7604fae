// Error reporting, device probing and TMA tensor-map creation for the C ABI.
#include "host_util.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace opsg {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  const char* e = getenv("OPSG_PDL");      // read per launch (host side, ~100 ns): tests flip it inside one process
  return e ? atoi(e) != 0 : true;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return OPSG_OK;
  return set_error(OPSG_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

static int make_tmap_bf16_2d_sw(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_bf16_2d_sw(map, base, rows, cols, ld, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_bf16_2d_plain(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                            uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_bf16_2d_sw(map, base, rows, cols, ld, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_NONE);
}

static int make_tmap_bf16_2d_sw(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(OPSG_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return set_error(OPSG_E_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(OPSG_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return OPSG_OK;
}

}  // namespace opsg

extern "C" int opsg_version(void) { return 100; }

extern "C" const char* opsg_last_error_string(void) { return opsg::g_err; }

static int g_dev_state = 0;  // 0 unknown, 1 ok, -1 bad
static int g_num_sms = 0;

extern "C" int opsg_device_check(void) {
  if (g_dev_state == 1) return OPSG_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return opsg::set_error(OPSG_E_NO_DEVICE, "no CUDA device: %s (libopsg_b200 has no CPU fallback)",
                           cudaGetErrorString(e));
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return opsg::set_error(OPSG_E_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  }
  if (prop.major != 10)
    return opsg::set_error(OPSG_E_NO_DEVICE, "device %s is sm_%d%d; libopsg_b200 is built for sm_100a only", prop.name,
                           prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  g_dev_state = 1;
  return OPSG_OK;
}

extern "C" int opsg_num_sms(void) {
  if (g_dev_state != 1 && opsg_device_check() != OPSG_OK) return 0;
  return g_num_sms;
}
