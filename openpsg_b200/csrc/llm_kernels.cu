// HBM-bound elementwise kernels of the Llama-family decode path (row a10 of SURVEY.md §8 for the LLM the shipped config
// names, configs/psg/baseline_v4_ov.py:60-61 -> HF models/llama/modeling_llama.py): RMSNorm (:52-69), rotary position
// embedding on the fused q | k rows (:137-166), SwiGLU gate (:181-183).  16-byte accesses, fp32 arithmetic, bf16 storage.
#include "common.cuh"
#include "host_util.h"

namespace opsg {

static inline int ceil_div_ll(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// y = weight * (x * rsqrt(mean(x^2) + eps)); one CTA of 256 threads per row, row held in registers (cols <= 8192).
__global__ void __launch_bounds__(256) rmsnorm_bf16_kernel(const __nv_bfloat16* __restrict__ x, int ld_x,
                                                           const float* __restrict__ weight, float eps,
                                                           __nv_bfloat16* __restrict__ y, int ld_y, int cols) {
  pdl_wait_then_trigger();
  __shared__ float red[8];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const __nv_bfloat16* xr = x + static_cast<size_t>(blockIdx.x) * ld_x;
  constexpr int CH = 4;                                   // 4 x 256 x 8 = 8192 columns max
  float v[CH][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c0 = (i * 256 + t) * 8;
    if (c0 < cols) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + c0));
      v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
      v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(v[i][j], v[i][j], ss);
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];               // fixed order: bit-reproducible
  const float rstd = rsqrtf(tot / cols + eps);
  __nv_bfloat16* yr = y + static_cast<size_t>(blockIdx.x) * ld_y;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c0 = (i * 256 + t) * 8;
    if (c0 < cols) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(weight + c0));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(weight + c0) + 1);
      uint4 u;
      u.x = pack_bf16x2(g0.x * (v[i][0] * rstd), g0.y * (v[i][1] * rstd));
      u.y = pack_bf16x2(g0.z * (v[i][2] * rstd), g0.w * (v[i][3] * rstd));
      u.z = pack_bf16x2(g1.x * (v[i][4] * rstd), g1.y * (v[i][5] * rstd));
      u.w = pack_bf16x2(g1.z * (v[i][6] * rstd), g1.w * (v[i][7] * rstd));
      *reinterpret_cast<uint4*>(yr + c0) = u;
    }
  }
}

// Rotary embedding, HF "rotate_half" convention: for e < hd/2
//   out[e]        = x[e] cos[pos, e]        - x[e + hd/2] sin[pos, e]
//   out[e + hd/2] = x[e + hd/2] cos[pos, e] + x[e] sin[pos, e]
// applied in place to `n_parts` consecutive [num_heads * head_dim] column blocks of every row (q and k of a fused qkv row).
// Thread = 8 consecutive e of one (row, part, head).
__global__ void __launch_bounds__(256) rope_bf16_kernel(__nv_bfloat16* __restrict__ x, int ld, int rows, int n_parts,
                                                        int num_heads, int head_dim, const int32_t* __restrict__ pos,
                                                        const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                        int table_rows) {
  pdl_wait_then_trigger();
  const int half = head_dim / 2, vec = half / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(rows) * n_parts * num_heads * vec;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % vec);
  long long r = idx / vec;
  const int head = static_cast<int>(r % num_heads); r /= num_heads;
  const int part = static_cast<int>(r % n_parts);
  const int row = static_cast<int>(r / n_parts);
  int p = pos[row];
  p = min(max(p, 0), table_rows - 1);
  __nv_bfloat16* base = x + static_cast<size_t>(row) * ld + (static_cast<size_t>(part) * num_heads + head) * head_dim + c * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(base);
  const uint4 b = *reinterpret_cast<const uint4*>(base + half);
  const float* cp = cos_t + static_cast<size_t>(p) * half + c * 8;
  const float* sp = sin_t + static_cast<size_t>(p) * half + c * 8;
  const float4 c0 = __ldg(reinterpret_cast<const float4*>(cp)), c1 = __ldg(reinterpret_cast<const float4*>(cp) + 1);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp)), s1 = __ldg(reinterpret_cast<const float4*>(sp) + 1);
  const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float x1[8] = {bf16_lo(a.x), bf16_hi(a.x), bf16_lo(a.y), bf16_hi(a.y), bf16_lo(a.z), bf16_hi(a.z), bf16_lo(a.w), bf16_hi(a.w)};
  const float x2[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y), bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
  float o1[8], o2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o1[j] = x1[j] * cs[j] - x2[j] * sn[j];
    o2[j] = x2[j] * cs[j] + x1[j] * sn[j];
  }
  uint4 ua, ub;
  ua.x = pack_bf16x2(o1[0], o1[1]); ua.y = pack_bf16x2(o1[2], o1[3]); ua.z = pack_bf16x2(o1[4], o1[5]); ua.w = pack_bf16x2(o1[6], o1[7]);
  ub.x = pack_bf16x2(o2[0], o2[1]); ub.y = pack_bf16x2(o2[2], o2[3]); ub.z = pack_bf16x2(o2[4], o2[5]); ub.w = pack_bf16x2(o2[6], o2[7]);
  *reinterpret_cast<uint4*>(base) = ua;
  *reinterpret_cast<uint4*>(base + half) = ub;
}

__device__ __forceinline__ float silu_f(float g) { return g / (1.f + __expf(-g)); }

// out[r, c] = silu(gu[r, c]) * gu[r, ffn + c]   (gate | up halves of one fused projection)
__global__ void __launch_bounds__(256) swiglu_bf16_kernel(const __nv_bfloat16* __restrict__ gu, int ld_gu, int rows, int ffn,
                                                          __nv_bfloat16* __restrict__ out, int ld_out) {
  pdl_wait_then_trigger();
  const int vec = ffn / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * vec) return;
  const int c = static_cast<int>(idx % vec);
  const long long r = idx / vec;
  const uint4 g = __ldg(reinterpret_cast<const uint4*>(gu + r * ld_gu) + c);
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(gu + r * ld_gu + ffn) + c);
  uint4 o;
  o.x = pack_bf16x2(silu_f(bf16_lo(g.x)) * bf16_lo(u.x), silu_f(bf16_hi(g.x)) * bf16_hi(u.x));
  o.y = pack_bf16x2(silu_f(bf16_lo(g.y)) * bf16_lo(u.y), silu_f(bf16_hi(g.y)) * bf16_hi(u.y));
  o.z = pack_bf16x2(silu_f(bf16_lo(g.z)) * bf16_lo(u.z), silu_f(bf16_hi(g.z)) * bf16_hi(u.z));
  o.w = pack_bf16x2(silu_f(bf16_lo(g.w)) * bf16_lo(u.w), silu_f(bf16_hi(g.w)) * bf16_hi(u.w));
  *(reinterpret_cast<uint4*>(out + r * ld_out) + c) = o;
}


// Prompt bookkeeping of one batched generate() (v4:294-301; HF opt :64-70; HF generation/utils.py:707-729), one thread per
// sequence: the prompt is [n_prefix projected rows ; left-padded text], so with m[t] = 1 for t < n_prefix else
// text_mask[t - n_prefix] and c = inclusive cumsum(m):  pos[t] = m ? c - 1 + pos_offset : max(pos_offset - 1, 0);
// key_mask[t] = m for t < Tp and 1 for the generated positions; last_rows = index of the last prompt row; the token fed at
// decode step j + 1 (the j-th generated one) sits at position n_valid + j + pos_offset.
__global__ void llm_prompt_layout_kernel(const int32_t* __restrict__ text_mask, int nseq, int T, int n_prefix, int max_new,
                                         int pos_offset, int32_t* __restrict__ pos, uint8_t* __restrict__ key_mask,
                                         int32_t* __restrict__ last_rows, int32_t* __restrict__ dec_pos) {
  pdl_wait_then_trigger();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseq) return;
  const int Tp = n_prefix + T, max_ctx = Tp + max_new;
  const int pad_pos = pos_offset > 0 ? pos_offset - 1 : 0;
  int c = 0;
  for (int t = 0; t < Tp; ++t) {
    const int m = t < n_prefix ? 1 : (text_mask[static_cast<size_t>(s) * T + (t - n_prefix)] != 0);
    c += m;
    pos[static_cast<size_t>(s) * Tp + t] = m ? c - 1 + pos_offset : pad_pos;
    key_mask[static_cast<size_t>(s) * max_ctx + t] = static_cast<uint8_t>(m);
  }
  for (int t = Tp; t < max_ctx; ++t) key_mask[static_cast<size_t>(s) * max_ctx + t] = 1;
  last_rows[s] = s * Tp + Tp - 1;
  for (int j = 0; j + 1 < max_new; ++j) dec_pos[static_cast<size_t>(j) * nseq + s] = c + j + pos_offset;
}

// dst = src (device to device) as a KERNEL: staging copies into CUDA-graph inputs must not queue on a copy engine behind
// an in-flight host->device prefetch of the next image (openpsg_b200/relation_qformer.py).
__global__ void __launch_bounds__(256) copy_bytes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t nbytes,
                                                         int vec16) {
  pdl_wait_then_trigger();
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec16) {
    const size_t n16 = nbytes / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t j = i; j < n16; j += stride) d4[j] = __ldg(s4 + j);
    for (size_t j = n16 * 16 + i; j < nbytes; j += stride) dst[j] = src[j];
  } else {
    for (; i < nbytes; i += stride) dst[i] = src[i];
  }
}

// dst[c, r] = src[r, c]  (int32; token matrix [T_new, k] -> [k, T_new])
__global__ void transpose_i32_kernel(const int32_t* __restrict__ src, int rows, int cols, int32_t* __restrict__ dst) {
  pdl_wait_then_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int c = i / rows, r = i % rows;
  dst[i] = src[static_cast<size_t>(r) * cols + c];
}


// out[r, c] = bf16(bias[c] + sum_s partials[s][r][c] (+ residual[r, c])), slices summed in split order (fixed order ->
// bit-reproducible); thread = 4 consecutive columns, 16-byte loads, `splits` independent loads in flight per thread in groups
// of 8.  residual may alias out (every thread reads its own 4 elements before it writes them).
__global__ void __launch_bounds__(256) splitk_reduce_bf16_kernel(const float* __restrict__ partials, int splits, int rows, int cols,
                                                                 const float* __restrict__ bias, const __nv_bfloat16* residual,
                                                                 int ld_res, __nv_bfloat16* out, int ld_out) {
  pdl_wait_then_trigger();
  const int vec = cols / 4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * vec) return;
  const int c = static_cast<int>(idx % vec);
  const long long r = idx / vec;
  const size_t slice = static_cast<size_t>(rows) * cols;
  const float4* src = reinterpret_cast<const float4*>(partials + r * cols) + c;
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  for (; s + 8 <= splits; s += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + (static_cast<size_t>(s + u) * slice) / 4);
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  for (; s < splits; ++s) {
    const float4 v = __ldcs(src + (static_cast<size_t>(s) * slice) / 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (residual) {
    const uint2 rv = *reinterpret_cast<const uint2*>(residual + r * ld_res + c * 4);
    acc.x += bf16_lo(rv.x); acc.y += bf16_hi(rv.x); acc.z += bf16_lo(rv.y); acc.w += bf16_hi(rv.y);
  }
  uint2 o;
  o.x = pack_bf16x2(acc.x, acc.y);
  o.y = pack_bf16x2(acc.z, acc.w);
  *reinterpret_cast<uint2*>(out + r * ld_out + c * 4) = o;
}



}  // namespace opsg

using namespace opsg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int opsg_rmsnorm_bf16(const opsg_bf16* x, int ld_x, const float* weight, float eps, opsg_bf16* y, int ld_y,
                                 int rows, int cols, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(x && weight && y, "rmsnorm: null pointer");
  OPSG_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 8192, "rmsnorm: cols=%d unsupported (multiple of 8, <= 8192)", cols);
  OPSG_CHECK_ARG(ld_x % 8 == 0 && ld_y % 8 == 0 && ld_x >= cols && ld_y >= cols &&
                 (((uintptr_t)x | (uintptr_t)y | (uintptr_t)weight) & 15) == 0, "rmsnorm: bad layout");
  launch_kernel(rmsnorm_bf16_kernel, rows, 256, 0, ST(stream), reinterpret_cast<const __nv_bfloat16*>(x), ld_x, weight, eps,
                reinterpret_cast<__nv_bfloat16*>(y), ld_y, cols);
  OPSG_CHECK_LAUNCH("rmsnorm_bf16_kernel");
  return OPSG_OK;
}

extern "C" int opsg_rope_bf16(opsg_bf16* x, int ld, int rows, int n_parts, int num_heads, int head_dim, const int32_t* pos,
                              const float* cos_table, const float* sin_table, int table_rows, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(x && pos && cos_table && sin_table, "rope: null pointer");
  OPSG_CHECK_ARG(rows > 0 && n_parts > 0 && num_heads > 0 && head_dim % 16 == 0 && table_rows > 0, "rope: bad shape");
  OPSG_CHECK_ARG(ld % 8 == 0 && ld >= n_parts * num_heads * head_dim &&
                 (((uintptr_t)x | (uintptr_t)cos_table | (uintptr_t)sin_table) & 15) == 0, "rope: bad layout");
  const long long total = static_cast<long long>(rows) * n_parts * num_heads * (head_dim / 16);
  launch_kernel(rope_bf16_kernel, ceil_div_ll(total, 256), 256, 0, ST(stream), reinterpret_cast<__nv_bfloat16*>(x), ld, rows,
                n_parts, num_heads, head_dim, pos, cos_table, sin_table, table_rows);
  OPSG_CHECK_LAUNCH("rope_bf16_kernel");
  return OPSG_OK;
}

extern "C" int opsg_swiglu_bf16(const opsg_bf16* gate_up, int ld_gu, int rows, int ffn, opsg_bf16* out, int ld_out,
                                void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(gate_up && out, "swiglu: null pointer");
  OPSG_CHECK_ARG(rows > 0 && ffn > 0 && ffn % 8 == 0 && ld_gu % 8 == 0 && ld_out % 8 == 0 && ld_gu >= 2 * ffn && ld_out >= ffn &&
                 (((uintptr_t)gate_up | (uintptr_t)out) & 15) == 0, "swiglu: bad layout");
  launch_kernel(swiglu_bf16_kernel, ceil_div_ll(static_cast<long long>(rows) * (ffn / 8), 256), 256, 0, ST(stream),
                reinterpret_cast<const __nv_bfloat16*>(gate_up), ld_gu, rows, ffn, reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  OPSG_CHECK_LAUNCH("swiglu_bf16_kernel");
  return OPSG_OK;
}

extern "C" int opsg_llm_prompt_layout(const int32_t* text_mask, int nseq, int T, int n_prefix, int max_new_tokens,
                                      int pos_offset, int32_t* pos_out, uint8_t* key_mask_out, int32_t* last_rows_out,
                                      int32_t* dec_pos_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(text_mask && pos_out && key_mask_out && last_rows_out && (dec_pos_out || max_new_tokens <= 1),
                 "llm_prompt_layout: null pointer");
  OPSG_CHECK_ARG(nseq > 0 && T >= 0 && n_prefix >= 0 && n_prefix + T > 0 && max_new_tokens >= 1 && pos_offset >= 0,
                 "llm_prompt_layout: bad shape");
  launch_kernel(llm_prompt_layout_kernel, (nseq + 63) / 64, 64, 0, ST(stream), text_mask, nseq, T, n_prefix, max_new_tokens,
                pos_offset, pos_out, key_mask_out, last_rows_out, dec_pos_out);
  OPSG_CHECK_LAUNCH("llm_prompt_layout_kernel");
  return OPSG_OK;
}

extern "C" int opsg_copy_bytes(void* dst, const void* src, size_t nbytes, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  if (nbytes == 0) return OPSG_OK;
  OPSG_CHECK_ARG(dst && src, "copy_bytes: null pointer");
  const int vec16 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
  const size_t items = vec16 ? (nbytes + 15) / 16 : nbytes;
  long long blocks = static_cast<long long>((items + 255) / 256);
  const long long cap = 8LL * opsg_num_sms();
  if (blocks > cap) blocks = cap;
  launch_kernel(copy_bytes_kernel, blocks, 256, 0, ST(stream), static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst),
                nbytes, vec16);
  OPSG_CHECK_LAUNCH("copy_bytes_kernel");
  return OPSG_OK;
}

extern "C" int opsg_transpose_i32(const int32_t* src, int rows, int cols, int32_t* dst, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(src && dst && rows > 0 && cols > 0, "transpose_i32: bad arguments");
  launch_kernel(transpose_i32_kernel, (rows * cols + 255) / 256, 256, 0, ST(stream), src, rows, cols, dst);
  OPSG_CHECK_LAUNCH("transpose_i32_kernel");
  return OPSG_OK;
}

extern "C" int opsg_splitk_reduce_bf16(const float* partials, int splits, int rows, int cols, const float* bias,
                                       const opsg_bf16* residual, int ld_res, opsg_bf16* out, int ld_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(partials && out, "splitk_reduce: null pointer");
  OPSG_CHECK_ARG(splits > 0 && rows > 0 && cols > 0 && cols % 4 == 0 && ld_out % 4 == 0 && ld_out >= cols, "splitk_reduce: bad shape");
  OPSG_CHECK_ARG((((uintptr_t)partials | (uintptr_t)bias) & 15) == 0 && ((uintptr_t)out & 7) == 0, "splitk_reduce: bad alignment");
  OPSG_CHECK_ARG(!residual || (ld_res % 4 == 0 && ld_res >= cols && ((uintptr_t)residual & 7) == 0), "splitk_reduce: bad residual layout");
  launch_kernel(splitk_reduce_bf16_kernel, ceil_div_ll(static_cast<long long>(rows) * (cols / 4), 256), 256, 0, ST(stream),
                partials, splits, rows, cols, bias, reinterpret_cast<const __nv_bfloat16*>(residual), ld_res,
                reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  OPSG_CHECK_LAUNCH("splitk_reduce_bf16_kernel");
  return OPSG_OK;
}

