// K11 — per-mask feature pooling + pair gather (row a11 of SURVEY.md §8; reference
// kings_sgg/models/detectors/openseed_relation.py:441-527, same code in mask2former_relation.py:275-295 and
// mask2former_relation_v2.py:392-465):
//
//   mask_o   = nearest(pad0(nearest(pan == id_o, img_shape), pad_shape), feature shape)            (:441-462)
//   obj_o    = sum(feat * mask_o) / (sum(mask_o) + 1e-8)      [N, C]                               (:466-468)
//            (+ class embedding, add or cat :469-474)  (+ background feature sum(feat * (1 - mask_o)) / ... :487-493)
//   pair_ij  = cat(obj_i, obj_j)                               [N*N, 2C']                          (:502-527)
//
// The reference materialises an [N, C, h, w] product (2.7 GB at N = 40).  Here the feature map is read ONCE, at HBM speed:
//   1. mask_pool_labels_kernel: the three-stage mask chain of ALL objects collapses into one label map at feature
//      resolution (label = first listed object whose id equals the source pixel of the panoptic map, N = nobody / padding);
//      objects that repeat an id share the label of the first one (`rep`).
//   2. mask_pool_accum_kernel: CTA = 16 channels x a strip of pixel tiles.  A warp takes a 32 x 4 pixel tile (8 lanes x 16
//      bytes per row = full 128-byte lines, 4 rows): 16-byte loads of the labels and of 16 channel rows (all in flight
//      together), then for every distinct label of the tile a recursive-halving shuffle reduction (16 values x 32 lanes
//      in 16 shuffles) into the warp's PRIVATE shared-memory accumulator [label][channel].  The tile is 2-D because the
//      kernel is instruction-bound on those reduction rounds (round 2 ncu: ALU pipe 54 %, 2.5 TB/s with 128 x 1 pixel
//      segments that cross ~3 panoptic regions each); a 32 x 4 tile mostly sees one label, which also takes a path
//      without per-pixel selects.  No atomics anywhere: the CTA sums its warps in a fixed order and writes one partial
//      per strip; the reduction over strips is a second small kernel, also in fixed order -> bit-reproducible.
//   3. mask_pool_finalize_kernel: division, class embedding, background feature; pair_concat_kernel: the N^2 gather.
// Algorithmic bytes: C*h*w*4 (features) + h*w*4 (labels, re-read per channel block from L2).
#include "common.cuh"
#include "host_util.h"

namespace opsg {

constexpr int kMpCB = 16;            // channels per CTA
constexpr int kMpSegPx = 128;        // pixels per warp step (32 lanes x 4)
constexpr int kMpWarps = 8;

__device__ __forceinline__ int mp_nearest_src(int dst, int in_size, int out_size) {
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  const int src = static_cast<int>(floorf(__fmul_rn(static_cast<float>(dst), scale)));
  return min(src, in_size - 1);
}

__global__ void __launch_bounds__(256)
mask_pool_labels_kernel(const int32_t* __restrict__ pan, int pan_h, int pan_w, int img_h, int img_w, int pad_h, int pad_w,
                        int feat_h, int feat_w, const int32_t* __restrict__ obj_ids, int num_objects,
                        int32_t* __restrict__ label, int32_t* __restrict__ rep) {
  pdl_wait_then_trigger();
  extern __shared__ int32_t s_ids[];
  for (int i = threadIdx.x; i < num_objects; i += blockDim.x) s_ids[i] = obj_ids[i];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < num_objects) {                       // objects that repeat an id share the first one's mask
    int r = idx;
    for (int o = 0; o < idx; ++o)
      if (s_ids[o] == s_ids[idx]) { r = o; break; }
    rep[idx] = r;
  }
  if (idx >= feat_h * feat_w) return;
  const int y = idx / feat_w, x = idx % feat_w;
  const int r2 = mp_nearest_src(y, pad_h, feat_h);     // feature row -> padded image row
  const int c2 = mp_nearest_src(x, pad_w, feat_w);
  int lab = num_objects;                                // zero padding: no object
  if (r2 < img_h && c2 < img_w) {
    const int r1 = mp_nearest_src(r2, pan_h, img_h);   // image row -> panoptic-map row
    const int c1 = mp_nearest_src(c2, pan_w, img_w);
    const int id = pan[static_cast<size_t>(r1) * pan_w + c1];
    for (int o = 0; o < num_objects; ++o)
      if (s_ids[o] == id) { lab = o; break; }
  }
  label[idx] = lab;
}

// 16 per-lane values -> sums over the 32 lanes; lane l (even) ends up with the total of value index (l >> 1).
__device__ __forceinline__ float mp_reduce16(float (&v)[16], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8];
    const float keep = b4 ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? v[i] : v[i + 4];
    const float keep = b3 ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? v[i] : v[i + 2];
    const float keep = b2 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const float send = b1 ? v[0] : v[1];
    const float keep = b1 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__global__ void __launch_bounds__(kMpWarps * 32)
mask_pool_accum_kernel(const float* __restrict__ feat, int C, int h, int w, int segs_per_strip, const int32_t* __restrict__ label,
                       int n_labels, float* __restrict__ partial, float* __restrict__ pcount) {
  pdl_wait_then_trigger();
  extern __shared__ float s_acc[];                      // [warp][label][16 channels] (+ [warp][label] counts)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x, cblk = blockIdx.y;
  const int c0 = cblk * kMpCB;
  float* acc = s_acc + static_cast<size_t>(warp) * n_labels * kMpCB;
  float* cnt = s_acc + static_cast<size_t>(kMpWarps) * n_labels * kMpCB + warp * n_labels;
  for (int i = lane; i < n_labels * kMpCB; i += 32) acc[i] = 0.f;
  for (int i = lane; i < n_labels; i += 32) cnt[i] = 0.f;
  __syncwarp();
  const int hw = h * w;
  // 2-D tiles need 16-byte aligned rows; otherwise a segment is 128 consecutive pixels of the flattened map
  const bool vec = (w & 3) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(label) & 15) == 0;
  const int tiles_x = (w + 31) >> 5;
  const int n_segs = vec ? tiles_x * ((h + 3) >> 2) : (hw + kMpSegPx - 1) / kMpSegPx;
  for (int seg = warp; seg < segs_per_strip; seg += kMpWarps) {
    const int sg = strip * segs_per_strip + seg;
    if (sg >= n_segs) break;                                                // warp-uniform
    int px = sg * kMpSegPx + lane * 4;
    bool inb = true;
    if (vec) {
      const int ty = sg / tiles_x, tx = sg - ty * tiles_x;
      const int y = ty * 4 + (lane >> 3), x = tx * 32 + (lane & 7) * 4;
      inb = y < h && x < w;
      px = y * w + x;
    }
    int lab[4];
    float v[kMpCB][4];
    if (vec && !inb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) lab[j] = -1;
#pragma unroll
      for (int cc = 0; cc < kMpCB; ++cc)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[cc][j] = 0.f;
    } else if (vec) {
      const int4 l4 = __ldg(reinterpret_cast<const int4*>(label + px));
      lab[0] = l4.x; lab[1] = l4.y; lab[2] = l4.z; lab[3] = l4.w;
#pragma unroll
      for (int cc = 0; cc < kMpCB; ++cc) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + cc < C) f = __ldcs(reinterpret_cast<const float4*>(feat + static_cast<size_t>(c0 + cc) * hw + px));   // streamed once
        v[cc][0] = f.x; v[cc][1] = f.y; v[cc][2] = f.z; v[cc][3] = f.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) lab[j] = (px + j < hw) ? label[px + j] : -1;
#pragma unroll
      for (int cc = 0; cc < kMpCB; ++cc)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[cc][j] = (px + j < hw && c0 + cc < C) ? __ldg(feat + static_cast<size_t>(c0 + cc) * hw + px + j) : 0.f;
    }
    uint32_t todo = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) todo |= (lab[j] >= 0 && lab[j] < n_labels) ? (1u << j) : 0u;
    // one round per distinct label of the segment (1-3 for panoptic regions; any pattern is handled)
    for (;;) {
      const uint32_t pending = __ballot_sync(0xffffffffu, todo != 0u);
      if (!pending) break;
      const int leader = __ffs(pending) - 1;
      int mine = 0;
#pragma unroll
      for (int j = 3; j >= 0; --j)
        if ((todo >> j) & 1u) mine = lab[j];                                // label of my first pending pixel
      const int cur = __shfl_sync(0xffffffffu, mine, leader);
      float part[kMpCB];
      uint32_t hit = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) hit |= (((todo >> j) & 1u) && lab[j] == cur) ? (1u << j) : 0u;
      todo &= ~hit;
      if (__all_sync(0xffffffffu, hit == 0xFu)) {                           // the whole tile is one label: no selects
#pragma unroll
        for (int cc = 0; cc < kMpCB; ++cc) part[cc] = (v[cc][0] + v[cc][1]) + (v[cc][2] + v[cc][3]);
      } else {
#pragma unroll
        for (int cc = 0; cc < kMpCB; ++cc) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) s += ((hit >> j) & 1u) ? v[cc][j] : 0.f;
          part[cc] = s;
        }
      }
      const float total = mp_reduce16(part, lane);
      if (!(lane & 1)) acc[cur * kMpCB + (lane >> 1)] += total;
      if (cblk == 0) {
        int n = __popc(hit);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) cnt[cur] += static_cast<float>(n);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // fixed-order sum over the CTA's warps -> this strip's partial
  for (int i = threadIdx.x; i < n_labels * kMpCB; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kMpWarps; ++w) s += s_acc[static_cast<size_t>(w) * n_labels * kMpCB + i];
    const int l = i / kMpCB, cc = i % kMpCB;
    if (c0 + cc < C) partial[(static_cast<size_t>(strip) * n_labels + l) * C + c0 + cc] = s;
  }
  if (cblk == 0)
    for (int l = threadIdx.x; l < n_labels; l += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kMpWarps; ++w) s += s_acc[static_cast<size_t>(kMpWarps) * n_labels * kMpCB + w * n_labels + l];
      pcount[static_cast<size_t>(strip) * n_labels + l] = s;
    }
}

// sums[l][c] = sum over strips (ascending) of partial[s][l][c]; counts likewise
__global__ void __launch_bounds__(256)
mask_pool_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ pcount, int n_strips, int n_labels, int C,
                        float* __restrict__ sums, float* __restrict__ counts) {
  pdl_wait_then_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n_labels * C;
  if (idx < total) {
    float s = 0.f;
    for (int st = 0; st < n_strips; ++st) s += partial[static_cast<size_t>(st) * total + idx];
    sums[idx] = s;
  }
  if (idx < n_labels) {
    float s = 0.f;
    for (int st = 0; st < n_strips; ++st) s += pcount[static_cast<size_t>(st) * n_labels + idx];
    counts[idx] = s;
  }
}

// obj[o, c] = sums[rep o][c] / (count + 1e-8)  (+ cls)  (+ background = (total - sums) / (hw - count + 1e-8))
__global__ void __launch_bounds__(256)
mask_pool_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ counts, const int32_t* __restrict__ rep,
                          int num_objects, int n_labels, int C, int hw, const float* __restrict__ cls_table,
                          const int32_t* __restrict__ cls_ids, int cls_dim, int cls_mode, int use_background,
                          float* __restrict__ obj, int ld_obj) {
  pdl_wait_then_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_objects * ld_obj) return;
  const int o = idx / ld_obj, c = idx % ld_obj;
  if (c >= C) {                                                        // cat mode: the class-embedding columns
    obj[idx] = cls_table[static_cast<size_t>(cls_ids[o]) * cls_dim + (c - C)];
    return;
  }
  const int l = rep ? rep[o] : o;
  const float s = sums[static_cast<size_t>(l) * C + c], n = counts[l];
  float e = s / (n + 1e-8f);
  if (cls_mode == 1) e += cls_table[static_cast<size_t>(cls_ids[o]) * cls_dim + c];
  if (use_background) {
    float tot = 0.f;
    for (int k = 0; k < n_labels; ++k) tot += sums[static_cast<size_t>(k) * C + c];
    e += (tot - s) / ((static_cast<float>(hw) - n) + 1e-8f);
  }
  obj[idx] = e;
}

__global__ void pair_concat_kernel(const float* __restrict__ obj, int N, int C, float* __restrict__ pair_out) {
  pdl_wait_then_trigger();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(N) * N * 2 * C) return;
  const int col = static_cast<int>(idx % (2 * C));
  const long long p = idx / (2 * C);
  const int i = static_cast<int>(p / N), j = static_cast<int>(p % N);
  pair_out[idx] = col < C ? obj[static_cast<size_t>(i) * C + col] : obj[static_cast<size_t>(j) * C + (col - C)];
}

static inline int mp_ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }
// Strip layout of the accumulation kernel: the map is cut into segments (32 x 4 pixel tiles when rows are 16-byte aligned,
// else 128 consecutive pixels); a CTA takes `per_strip` consecutive segments of 16 channels.  The strip count is chosen so
// that strips x channel blocks fills the SMs twice (two resident CTAs per SM at 116 registers) in ONE wave.
static inline void mp_layout(int channels, int h, int w, int* strips, int* per_strip) {
  const int segs = (w & 3) == 0 ? ((w + 31) >> 5) * ((h + 3) >> 2) : mp_ceil_div(static_cast<long long>(h) * w, kMpSegPx);
  const int cblks = mp_ceil_div(channels, kMpCB);
  int sms = opsg_num_sms();
  if (sms <= 0) sms = kNumSMsB200;
  int want = 2 * sms / cblks;
  if (want < 1) want = 1;
  if (want > segs) want = segs;
  *per_strip = mp_ceil_div(segs, want);
  if (*per_strip < kMpWarps && segs >= kMpWarps) *per_strip = kMpWarps;     // every warp of a CTA gets a segment
  *strips = mp_ceil_div(segs, *per_strip);
}

}  // namespace opsg

using namespace opsg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int opsg_mask_pool_labels(const int32_t* pan, int pan_h, int pan_w, int img_h, int img_w, int pad_h, int pad_w,
                                     int feat_h, int feat_w, const int32_t* obj_ids, int num_objects, int32_t* label_out,
                                     int32_t* rep_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(pan && obj_ids && label_out && rep_out, "mask_pool_labels: null pointer");
  OPSG_CHECK_ARG(pan_h > 0 && pan_w > 0 && img_h > 0 && img_w > 0 && feat_h > 0 && feat_w > 0 && num_objects > 0 &&
                 num_objects <= 4096, "mask_pool_labels: bad shape");
  OPSG_CHECK_ARG(pad_h >= img_h && pad_w >= img_w, "mask_pool_labels: pad_shape smaller than img_shape");
  const long long threads = static_cast<long long>(feat_h) * feat_w > num_objects ? static_cast<long long>(feat_h) * feat_w : num_objects;
  launch_kernel(mask_pool_labels_kernel, mp_ceil_div(threads, 256), 256, static_cast<size_t>(num_objects) * 4, ST(stream), pan, pan_h,
                pan_w, img_h, img_w, pad_h, pad_w, feat_h, feat_w, obj_ids, num_objects, label_out, rep_out);
  OPSG_CHECK_LAUNCH("mask_pool_labels_kernel");
  return OPSG_OK;
}

extern "C" size_t opsg_mask_pool_workspace_bytes(int channels, int h, int w, int num_objects) {
  if (channels <= 0 || h <= 0 || w <= 0 || num_objects <= 0) return 0;
  int n_strips, per_strip;
  mp_layout(channels, h, w, &n_strips, &per_strip);
  const size_t nl = static_cast<size_t>(num_objects) + 1, strips = static_cast<size_t>(n_strips);
  return (strips * nl * channels + strips * nl + nl * channels + nl) * sizeof(float);
}

extern "C" int opsg_mask_pool_pairs(const float* feat, int channels, int h, int w, const int32_t* label, const int32_t* rep,
                                    int num_objects, const float* cls_table, const int32_t* cls_ids, int cls_dim, int cls_mode,
                                    int use_background, float* workspace, size_t workspace_bytes, float* obj_out,
                                    float* pair_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(feat && label && obj_out && workspace, "mask_pool_pairs: null pointer");
  OPSG_CHECK_ARG(channels > 0 && h > 0 && w > 0 && num_objects > 0 && num_objects <= 255, "mask_pool_pairs: bad shape (1..255 objects)");
  OPSG_CHECK_ARG(cls_mode >= 0 && cls_mode <= 2, "mask_pool_pairs: cls_mode must be 0 (none), 1 (add) or 2 (cat)");
  OPSG_CHECK_ARG(cls_mode == 0 || (cls_table && cls_ids && cls_dim > 0), "mask_pool_pairs: class embedding table / ids missing");
  OPSG_CHECK_ARG(cls_mode != 1 || cls_dim == channels, "mask_pool_pairs: 'add' needs cls_dim == channels");
  OPSG_CHECK_ARG(!(cls_mode == 2 && use_background), "mask_pool_pairs: background feature cannot be added to a 'cat' embedding "
                                                     "(the reference's broadcast fails there too)");
  OPSG_CHECK_ARG(workspace_bytes >= opsg_mask_pool_workspace_bytes(channels, h, w, num_objects), "mask_pool_pairs: workspace too small");
  const int hw = h * w, nl = num_objects + 1;
  int strips, per_strip;
  mp_layout(channels, h, w, &strips, &per_strip);
  float* partial = workspace;
  float* pcount = partial + static_cast<size_t>(strips) * nl * channels;
  float* sums = pcount + static_cast<size_t>(strips) * nl;
  float* counts = sums + static_cast<size_t>(nl) * channels;
  const size_t smem = (static_cast<size_t>(kMpWarps) * nl * kMpCB + static_cast<size_t>(kMpWarps) * nl) * sizeof(float);
  static int configured_bytes_dev[64] = {};
  int& configured_bytes = configured_bytes_dev[device_slot()];
  if (static_cast<int>(smem) > configured_bytes && smem > 48 * 1024) {
    rc = check_cuda(cudaFuncSetAttribute(mask_pool_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024),
                    "cudaFuncSetAttribute(mask_pool_accum)");
    if (rc) return rc;
    configured_bytes = 160 * 1024;
  }
  launch_kernel(mask_pool_accum_kernel, dim3(strips, mp_ceil_div(channels, kMpCB)), kMpWarps * 32, smem, ST(stream), feat, channels,
                h, w, per_strip, label, nl, partial, pcount);
  OPSG_CHECK_LAUNCH("mask_pool_accum_kernel");
  launch_kernel(mask_pool_reduce_kernel, mp_ceil_div(static_cast<long long>(nl) * channels, 256), 256, 0, ST(stream), partial, pcount,
                strips, nl, channels, sums, counts);
  OPSG_CHECK_LAUNCH("mask_pool_reduce_kernel");
  const int ld_obj = channels + (cls_mode == 2 ? cls_dim : 0);
  launch_kernel(mask_pool_finalize_kernel, mp_ceil_div(static_cast<long long>(num_objects) * ld_obj, 256), 256, 0, ST(stream), sums,
                counts, rep, num_objects, nl, channels, hw, cls_table, cls_ids, cls_dim, cls_mode, use_background, obj_out, ld_obj);
  OPSG_CHECK_LAUNCH("mask_pool_finalize_kernel");
  if (pair_out) {
    const long long total = static_cast<long long>(num_objects) * num_objects * 2 * ld_obj;
    launch_kernel(pair_concat_kernel, mp_ceil_div(total, 256), 256, 0, ST(stream), obj_out, num_objects, ld_obj, pair_out);
    OPSG_CHECK_LAUNCH("pair_concat_kernel");
  }
  return OPSG_OK;
}
