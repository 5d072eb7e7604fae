// Rows f1 / f2 of SURVEY.md §8 — the integer / byte passes that sit directly before and after the relation head in the
// reference's inference loop, as HBM-bound one-pass kernels (16-byte loads, id tables in shared memory):
//   * pan_relabel_kernel  (kings_sgg/models/detectors/openseed_relation_v2.py:112-128): the detector's panoptic segment ids
//     -> category + 1000 * instance ids.  The reference copies the map to the host and runs one np.where pass per segment
//     (D2H, S passes over H x W, H2D) right in front of the head on every image; here one pass on the device, no sync.
//   * pan_colorize_kernel (tools/infer.py:149-169): the RGB-encoded panoptic PNG of the submission format — every listed
//     object adds its colour to the pixels it owns (uint8 wrap-around like the reference's int sum cast to uint8).
#include "common.cuh"
#include "host_util.h"

namespace opsg {

constexpr int kGlueMaxSeg = 1024;

__global__ void __launch_bounds__(256)
pan_relabel_kernel(const int32_t* __restrict__ pan_in, long long n, const int32_t* __restrict__ seg_ids,
                   const int32_t* __restrict__ new_ids, int n_seg, int32_t* __restrict__ pan_out) {
  pdl_wait_then_trigger();
  __shared__ int32_t s_old[kGlueMaxSeg], s_new[kGlueMaxSeg];
  for (int i = threadIdx.x; i < n_seg; i += blockDim.x) { s_old[i] = seg_ids[i]; s_new[i] = new_ids[i]; }
  __syncthreads();
  auto map = [&](int32_t v) {
    int32_t out = 0;                                   // np.zeros_like: pixels of unlisted segments stay 0
    for (int s = 0; s < n_seg; ++s)
      if (s_old[s] == v) out = s_new[s];               // later segments overwrite earlier ones, as the sequential np.where does
    return out;
  };
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(pan_in) | reinterpret_cast<uintptr_t>(pan_out)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int4 v = __ldcs(reinterpret_cast<const int4*>(pan_in) + i);
    int4 o;
    o.x = map(v.x);
    o.y = (v.y == v.x) ? o.x : map(v.y);               // panoptic maps are piecewise constant: neighbours usually agree
    o.z = (v.z == v.y) ? o.y : map(v.z);
    o.w = (v.w == v.z) ? o.z : map(v.w);
    reinterpret_cast<int4*>(pan_out)[i] = o;
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    pan_out[i] = map(pan_in[i]);
}

__global__ void __launch_bounds__(256)
pan_colorize_kernel(const int32_t* __restrict__ pan, long long n, const int32_t* __restrict__ obj_ids,
                    const uint8_t* __restrict__ colors, int n_obj, uint8_t* __restrict__ out) {
  pdl_wait_then_trigger();
  __shared__ int32_t s_id[kGlueMaxSeg];
  __shared__ uint32_t s_col[kGlueMaxSeg];
  for (int i = threadIdx.x; i < n_obj; i += blockDim.x) {
    s_id[i] = obj_ids[i];
    s_col[i] = colors[3 * i] | (colors[3 * i + 1] << 8) | (colors[3 * i + 2] << 16);
  }
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t v = __ldcs(pan + i);
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    for (int s = 0; s < n_obj; ++s)
      if (s_id[s] == v) { c0 += s_col[s] & 255u; c1 += (s_col[s] >> 8) & 255u; c2 += (s_col[s] >> 16) & 255u; }
    out[3 * i] = static_cast<uint8_t>(c0);             // int sum cast to uint8 (tools/infer.py:163,168)
    out[3 * i + 1] = static_cast<uint8_t>(c1);
    out[3 * i + 2] = static_cast<uint8_t>(c2);
  }
}

}  // namespace opsg

using namespace opsg;

static int glue_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = 8LL * opsg_num_sms();
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int opsg_pan_relabel(const int32_t* pan_in, long long n_pixels, const int32_t* seg_ids, const int32_t* new_ids,
                                int n_segments, int32_t* pan_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(pan_in && pan_out && n_pixels > 0, "pan_relabel: bad arguments");
  OPSG_CHECK_ARG(n_segments >= 0 && n_segments <= kGlueMaxSeg && (n_segments == 0 || (seg_ids && new_ids)),
                 "pan_relabel: 0..%d segments", kGlueMaxSeg);
  launch_kernel(pan_relabel_kernel, glue_grid((n_pixels + 3) / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream), pan_in, n_pixels,
                seg_ids, new_ids, n_segments, pan_out);
  OPSG_CHECK_LAUNCH("pan_relabel_kernel");
  return OPSG_OK;
}

extern "C" int opsg_pan_colorize(const int32_t* pan, long long n_pixels, const int32_t* obj_ids, const uint8_t* colors,
                                 int n_objects, uint8_t* out_rgb, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(pan && out_rgb && n_pixels > 0, "pan_colorize: bad arguments");
  OPSG_CHECK_ARG(n_objects >= 0 && n_objects <= kGlueMaxSeg && (n_objects == 0 || (obj_ids && colors)),
                 "pan_colorize: 0..%d objects", kGlueMaxSeg);
  launch_kernel(pan_colorize_kernel, glue_grid(n_pixels), 256, 0, reinterpret_cast<cudaStream_t>(stream), pan, n_pixels, obj_ids,
                colors, n_objects, out_rgb);
  OPSG_CHECK_LAUNCH("pan_colorize_kernel");
  return OPSG_OK;
}
