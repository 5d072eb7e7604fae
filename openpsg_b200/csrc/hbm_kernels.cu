// HBM-bound integer / byte / elementwise kernels of the relation-head path: pair-mask bits (K2), PatchEmbed
// operand re-layout (K1 input side), Q-Former embeddings + LayerNorm (K7), LayerNorm, existence filter +
// exact top-k (K8), row/embedding gathers, argmax.  (K11, mask mean-pool + pair gather: mask_pool.cu.)
// Coalesced 16-byte accesses, warp-shuffle reductions, no tensor cores (none of this is GEMM-shaped).
#include "common.cuh"
#include "host_util.h"

namespace opsg {

static inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// K2: pan id map -> object token bitmasks.  One warp per (object, 32-token word).
// nearest index = min(int(floorf(dst * (float(in) / float(out)))), in - 1)  (ATen legacy 'nearest').
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  const int src = static_cast<int>(floorf(__fmul_rn(static_cast<float>(dst), scale)));
  return min(src, in_size - 1);
}

__global__ void pair_mask_bits_kernel(const int32_t* __restrict__ pan, int pan_h, int pan_w, int img_h, int img_w,
                                      int pad_h, int pad_w, int tok_h, int tok_w, const int32_t* __restrict__ obj_ids,
                                      int num_objects, uint32_t* __restrict__ bits, int words) {
  pdl_wait_then_trigger();
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_global >= num_objects * words) return;
  const int obj = warp_global / words;
  const int word = warp_global % words;
  const int l = word * 32 + lane;
  const int L = tok_h * tok_w;
  bool hit = false;
  if (l < L) {
    const int ty = l / tok_w, tx = l % tok_w;
    const int r2 = nearest_src(ty, pad_h, tok_h);   // token row -> padded image row   (v4:422-423)
    const int c2 = nearest_src(tx, pad_w, tok_w);
    float value = 0.f;                              // F.pad(value=0)                   (v4:420-421)
    if (r2 < img_h && c2 < img_w) {
      const int r1 = nearest_src(r2, pan_h, img_h); // image row -> pan row             (v4:417-418)
      const int c1 = nearest_src(c2, pan_w, img_w);
      value = static_cast<float>(pan[static_cast<size_t>(r1) * pan_w + c1]);   // .float() round trip
    }
    hit = (value == static_cast<float>(obj_ids[obj]));
  }
  const uint32_t w = __ballot_sync(0xffffffffu, hit);
  if (lane == 0) bits[static_cast<size_t>(obj) * words + word] = w;
}

// ------------------------------------------------------------------------------------------------
// K1 input side: fp32 [C,h,w] -> bf16 [L, C*p*p], K order (c, py, px).  Thread = 8 consecutive px.
// ------------------------------------------------------------------------------------------------
__global__ void patch_im2col_kernel(const float* __restrict__ feat, int C, int h, int w, int patch, int th, int tw,
                                    __nv_bfloat16* __restrict__ out) {
  pdl_wait_then_trigger();
  const int x8_per_row = (tw * patch) / 8;
  const long long total = static_cast<long long>(C) * (th * patch) * x8_per_row;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x8 = static_cast<int>(idx % x8_per_row);
  const long long t = idx / x8_per_row;
  const int y = static_cast<int>(t % (th * patch));
  const int c = static_cast<int>(t / (th * patch));
  const int x = x8 * 8;
  const float* src = feat + (static_cast<size_t>(c) * h + y) * w + x;
  float f[8];
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __ldg(src + j);
  }
  const int ty = y / patch, py = y % patch, tx = x / patch, px = x % patch;
  const size_t K = static_cast<size_t>(C) * patch * patch;
  __nv_bfloat16* dst = out + (static_cast<size_t>(ty) * tw + tx) * K + (static_cast<size_t>(c) * patch + py) * patch + px;
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, int ld_in, __nv_bfloat16* __restrict__ out, int ld_out,
                                     int rows, int cols) {
  pdl_wait_then_trigger();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(idx / cols), c = static_cast<int>(idx % cols);
  out[static_cast<size_t>(r) * ld_out + c] = __float2bfloat16(in[static_cast<size_t>(r) * ld_in + c]);
}

__global__ void init_rows_f32_kernel(float* __restrict__ out, int ld_out, const float* __restrict__ row, int rows, int cols) {
  pdl_wait_then_trigger();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(idx / cols), c = static_cast<int>(idx % cols);
  out[static_cast<size_t>(r) * ld_out + c] = row ? row[c] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm helpers: one warp per row; row kept in registers (cols <= 32 * 8 * kMaxChunks).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxChunks = 12;   // 12 * 256 = 3072 columns

template <int CHUNKS>
__device__ __forceinline__ void ln_normalize_store(float (&v)[CHUNKS][8], int cols, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i)
    if ((i * 32 + lane) * 8 < cols) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  const float mean = warp_sum(s) / cols;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i)
    if ((i * 32 + lane) * 8 < cols) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
    }
  const float rstd = rsqrtf(warp_sum(ss) / cols + eps);
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    if (c0 < cols) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(y + c0) = u;
    }
  }
}

template <int CHUNKS>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             __nv_bfloat16* __restrict__ y, int rows, int cols) {
  pdl_wait_then_trigger();
  // Rows are walked from the END of the tensor: the producer (a GEMM epilogue) wrote them in ascending order, so the last
  // ~100 MB are still in L2 when this kernel starts; reading those first saves their HBM round trip.
  const int row = rows - 1 - static_cast<int>((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row < 0) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(row) * cols;
  float v[CHUNKS][8];
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    if (c0 < cols) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + c0));
      v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
      v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
    }
  }
  ln_normalize_store<CHUNKS>(v, cols, gamma, beta, eps, y + static_cast<size_t>(row) * cols);
}

// LayerNorm over many rows of <= 1024 columns (the Q-Former's 8 LayerNorm passes per image: 11 % of the step).  The
// one-row-per-warp kernel above holds its loads in registers: ~27 warps x 1.5 KB in flight per SM = 4 TB/s at HBM latency
// (ncu: DRAM 38-45 %, issue slots 52 %).  Here the bytes in flight live in shared memory instead: a CTA walks slabs of 8
// consecutive rows (ONE 1-D bulk copy of 12 KB each) through a 4-stage ring, warp w normalises row w of the slab out of shared
// memory with gamma / beta held in registers (111 registers, two CTAs per SM), and stores straight from registers.  Same lane
// <-> column assignment and reduction order as ln_normalize_store.  Measured per 32-image step: 8.4 ms against 9.4 for the
// kernel above (4.3 against 3.9 TB/s); the variant that reloads gamma / beta per row at four CTAs per SM: 11.0 ms.  ncu
// (profiles/r2_ncu_layernorm_ring.md): 34.5 us per launch, issue slots 55 %, DRAM 40 % -- still bound by the per-row chain.  Slabs are walked from the END of the tensor
// (what the producing GEMM wrote last is still in L2).
constexpr int kLnRingStages = 4, kLnRingRows = 8;

template <int CHUNKS, bool FULL>                             // FULL: cols == CHUNKS * 256 exactly (768: no column guards)
__global__ void __launch_bounds__(256, 2) layernorm_ring_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps,
                                                                 __nv_bfloat16* __restrict__ y, int rows, int cols, int slabs_per_cta) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ __align__(8) uint64_t full[kLnRingStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_bytes = cols * 2, stage_bytes = kLnRingRows * row_bytes;
  const int total_slabs = (rows + kLnRingRows - 1) / kLnRingRows;
  const int j0 = blockIdx.x * slabs_per_cta;
  const int n_slabs = min(slabs_per_cta, total_slabs - j0);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kLnRingStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  // gamma / beta of this lane's columns stay in registers for all rows of the CTA (constant inputs: loaded before the wait)
  float g[CHUNKS][8], b[CHUNKS][8];
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    if (FULL || c0 < cols) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
      g[i][0] = g0.x; g[i][1] = g0.y; g[i][2] = g0.z; g[i][3] = g0.w; g[i][4] = g1.x; g[i][5] = g1.y; g[i][6] = g1.z; g[i][7] = g1.w;
      b[i][0] = b0.x; b[i][1] = b0.y; b[i][2] = b0.z; b[i][3] = b0.w; b[i][4] = b1.x; b[i][5] = b1.y; b[i][6] = b1.z; b[i][7] = b1.w;
    }
  }
  __syncthreads();
  pdl_wait_then_trigger();
  if (n_slabs <= 0) return;
  // slab j (0 = the LAST 8 rows of the tensor) covers rows [max(0, rows - 8 (j + 1)), rows - 8 j)
  auto issue = [&](int t) {                                   // thread 0: t-th slab of this CTA into stage t % stages
    const int j = j0 + t;
    const int hi = rows - kLnRingRows * j, lo = max(0, hi - kLnRingRows);
    const uint32_t bytes = static_cast<uint32_t>(hi - lo) * row_bytes;
    uint64_t* bar = &full[t % kLnRingStages];
    mbar_expect_tx(bar, bytes);
    bulk_load_1d(ln_smem + (t % kLnRingStages) * stage_bytes, x + static_cast<size_t>(lo) * cols, bytes, bar);
  };
  if (threadIdx.x == 0)
    for (int t = 0; t < min(n_slabs, kLnRingStages); ++t) issue(t);
  for (int t = 0; t < n_slabs; ++t) {
    const int st = t % kLnRingStages;
    const int j = j0 + t;
    const int hi = rows - kLnRingRows * j, lo = max(0, hi - kLnRingRows);
    mbar_wait(&full[st], (t / kLnRingStages) & 1);
    if (lo + warp < hi) {
      const uint8_t* src = ln_smem + st * stage_bytes + warp * row_bytes;
      float v[CHUNKS][8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) {
        const int c0 = (i * 32 + lane) * 8;
        if (FULL || c0 < cols) {
          const uint4 u = *reinterpret_cast<const uint4*>(src + c0 * 2);
          v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
          v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
#pragma unroll
          for (int k = 0; k < 8; ++k) s += v[i][k];
        }
      }
      const float mean = warp_sum(s) / cols;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i)
        if (FULL || (i * 32 + lane) * 8 < cols) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; ss += d * d; }
        }
      const float rstd = rsqrtf(warp_sum(ss) / cols + eps);
      __nv_bfloat16* yr = y + static_cast<size_t>(lo + warp) * cols;
#pragma unroll
      for (int i = 0; i < CHUNKS; ++i) {
        const int c0 = (i * 32 + lane) * 8;
        if (FULL || c0 < cols) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * g[i][k] + b[i][k];
          uint4 u;
          u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
          u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(yr + c0) = u;
        }
      }
    }
    if (t + kLnRingStages < n_slabs) {                        // the stage is free once every warp has read its row
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) issue(t + kLnRingStages);
    }
  }
}

// K7: row p*nq+q = LN(query[q]); row B*nq + p*T + t = LN(word_emb[ids[p,t]] + pos_emb[t]).  d <= 1024.
__global__ void __launch_bounds__(256) qformer_embed_ln_kernel(const float* __restrict__ query, int nq,
                                                               const int32_t* __restrict__ ids, int B, int T,
                                                               const float* __restrict__ word_emb, int vocab,
                                                               const float* __restrict__ pos_emb,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float eps, int d, __nv_bfloat16* __restrict__ out) {
  pdl_wait_then_trigger();
  // work item w: 0 .. nq-1 = the query rows of pair 0 (every other pair gets a copy, qformer_broadcast_rows_kernel),
  // nq .. nq + B*T - 1 = the text rows
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int n_query_rows = B * nq;
  if (w >= nq + B * T) return;
  const int row = w < nq ? w : n_query_rows + (w - nq);
  const float* src;
  const float* pos = nullptr;
  if (row < n_query_rows) {
    src = query + static_cast<size_t>(row % nq) * d;
  } else {
    const int r = row - n_query_rows;
    int id = ids[r];
    id = min(max(id, 0), vocab - 1);
    src = word_emb + static_cast<size_t>(id) * d;
    pos = pos_emb + static_cast<size_t>(r % T) * d;
  }
  float v[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c0 = (i * 32 + lane) * 8;
    if (c0 < d) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + c0) + 1);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
      if (pos) {
        const float4 pa = __ldg(reinterpret_cast<const float4*>(pos + c0));
        const float4 pb = __ldg(reinterpret_cast<const float4*>(pos + c0) + 1);
        v[i][0] += pa.x; v[i][1] += pa.y; v[i][2] += pa.z; v[i][3] += pa.w;
        v[i][4] += pb.x; v[i][5] += pb.y; v[i][6] += pb.z; v[i][7] += pb.w;
      }
    }
  }
  ln_normalize_store<4>(v, d, gamma, beta, eps, out + static_cast<size_t>(row) * d);
}

// ------------------------------------------------------------------------------------------------
// K8: existence logits (warp per pair) + exact stable top-k by rank counting.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exist_logits_kernel(const __nv_bfloat16* __restrict__ x, int ld_x, int B, int d,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float logit_threshold, float* __restrict__ logits,
                                                           float* __restrict__ probs, uint8_t* __restrict__ mask) {
  pdl_wait_then_trigger();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const __nv_bfloat16* xr = x + static_cast<size_t>(row) * ld_x;
  float acc = 0.f;
  for (int c0 = lane * 8; c0 < d; c0 += 256) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + c0));
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + c0) + 1);
    acc += bf16_lo(u.x) * w0.x + bf16_hi(u.x) * w0.y + bf16_lo(u.y) * w0.z + bf16_hi(u.y) * w0.w +
           bf16_lo(u.z) * w1.x + bf16_hi(u.z) * w1.y + bf16_lo(u.w) * w1.z + bf16_hi(u.w) * w1.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float z = acc + b[0];
    logits[row] = z;
    if (probs) probs[row] = 1.f / (1.f + expf(-z));
    if (mask) mask[row] = z > logit_threshold ? 1 : 0;
  }
}

// rank(i) = #{j : z_j > z_i or (z_j == z_i and j < i)};  rank < k  ->  topk[rank] = i.
// NaN logits (never produced by finite weights) are ordered as -inf so that the order stays total: every rank is taken
// exactly once and all k slots of topk are written.
// Exact top-k by rank counting (ties -> lower index, as torch.topk / sort of the reference, v4:236-237): candidate i's
// rank = #{j : z_j > z_i or (z_j == z_i and j < i)}.  A CTA ranks 32 candidates; each of its 8 warps counts over one
// eighth of the list (lane = candidate), so a thread walks B / 8 values instead of B (25 us -> ~4 us at B = 1600).
__global__ void __launch_bounds__(256) topk_rank_kernel(const float* __restrict__ logits, int B, int k,
                                                        int32_t* __restrict__ topk) {
  pdl_wait_then_trigger();
  __shared__ int s_rank[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 32 + lane;
  if (threadIdx.x < 32) s_rank[threadIdx.x] = 0;
  __syncthreads();
  float zi = i < B ? __ldg(logits + i) : 0.f;
  if (zi != zi) zi = -INFINITY;
  const int per = (B + 7) / 8;
  const int j0 = warp * per, j1 = min(B, j0 + per);
  int rank = 0;
  for (int j = j0; j < j1; ++j) {
    float zj = __ldg(logits + j);                        // same address across the warp: one broadcast load
    if (zj != zj) zj = -INFINITY;
    rank += (zj > zi || (zj == zi && j < i)) ? 1 : 0;
  }
  if (i < B) atomicAdd(&s_rank[lane], rank);
  __syncthreads();
  if (warp == 0 && i < B && s_rank[lane] < k) topk[s_rank[lane]] = i;
}

// ------------------------------------------------------------------------------------------------
// gathers / argmax
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_bf16_kernel(const uint4* __restrict__ src, int row_vec, const int32_t* __restrict__ idx, int n_rows,
                                        uint4* __restrict__ out) {
  pdl_wait_then_trigger();
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(n_rows) * row_vec) return;
  const int r = static_cast<int>(t / row_vec), c = static_cast<int>(t % row_vec);
  out[t] = __ldg(src + static_cast<size_t>(idx[r]) * row_vec + c);
}

__global__ void embed_gather_kernel(const __nv_bfloat16* __restrict__ table, int d, const int32_t* __restrict__ ids,
                                    const __nv_bfloat16* __restrict__ pos_table, const int32_t* __restrict__ pos, int n_rows,
                                    __nv_bfloat16* __restrict__ out, int ld_out) {
  pdl_wait_then_trigger();
  const int vec = d / 8;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(n_rows) * vec) return;
  const int r = static_cast<int>(t / vec), c = static_cast<int>(t % vec);
  uint4 u = __ldg(reinterpret_cast<const uint4*>(table + static_cast<size_t>(ids[r]) * d) + c);
  if (pos_table) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(pos_table + static_cast<size_t>(pos[r]) * d) + c);
    u.x = pack_bf16x2(bf16_lo(u.x) + bf16_lo(q.x), bf16_hi(u.x) + bf16_hi(q.x));
    u.y = pack_bf16x2(bf16_lo(u.y) + bf16_lo(q.y), bf16_hi(u.y) + bf16_hi(q.y));
    u.z = pack_bf16x2(bf16_lo(u.z) + bf16_lo(q.z), bf16_hi(u.z) + bf16_hi(q.z));
    u.w = pack_bf16x2(bf16_lo(u.w) + bf16_lo(q.w), bf16_hi(u.w) + bf16_hi(q.w));
  }
  *(reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * ld_out) + c) = u;
}

// a9: out[s, t] = (t < n_prefix ? proj[s * rows_per_seq + row0 + t] : table[ids[s, t - n_prefix]]) + pos_table[pos[s, t]]
__global__ void llm_build_prefix_kernel(const __nv_bfloat16* __restrict__ proj, int rows_per_seq, int row0, int n_prefix,
                                        const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ ids, int T,
                                        const __nv_bfloat16* __restrict__ pos_table, const int32_t* __restrict__ pos,
                                        int nseq, int d, __nv_bfloat16* __restrict__ out) {
  pdl_wait_then_trigger();
  const int vec = d / 8;
  const int Tp = n_prefix + T;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(nseq) * Tp * vec) return;
  const int c = static_cast<int>(idx % vec);
  const long long r = idx / vec;
  const int t = static_cast<int>(r % Tp), s = static_cast<int>(r / Tp);
  const __nv_bfloat16* src = t < n_prefix ? proj + (static_cast<size_t>(s) * rows_per_seq + row0 + t) * d
                                          : table + static_cast<size_t>(ids[static_cast<size_t>(s) * T + (t - n_prefix)]) * d;
  uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + c);
  if (pos_table) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(pos_table + static_cast<size_t>(pos[r]) * d) + c);
    u.x = pack_bf16x2(bf16_lo(u.x) + bf16_lo(q.x), bf16_hi(u.x) + bf16_hi(q.x));
    u.y = pack_bf16x2(bf16_lo(u.y) + bf16_lo(q.y), bf16_hi(u.y) + bf16_hi(q.y));
    u.z = pack_bf16x2(bf16_lo(u.z) + bf16_lo(q.z), bf16_hi(u.z) + bf16_hi(q.z));
    u.w = pack_bf16x2(bf16_lo(u.w) + bf16_lo(q.w), bf16_hi(u.w) + bf16_hi(q.w));
  }
  *(reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * d) + c) = u;
}

// First maximum of every row (ties -> lower index).  One CTA of 1024 threads per row; 16-byte loads, four of them in
// flight per thread (the first cut walked the row with dependent 4-byte loads: 94 us for 100 x 50272 logits).
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const float* __restrict__ logits, int ld, int rows, int cols,
                                                           int32_t* __restrict__ out) {
  pdl_wait_then_trigger();
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const int row = blockIdx.x;
  const float* x = logits + static_cast<size_t>(row) * ld;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  auto take = [&](float v, int c) {
    if (v > best || (v == best && c < best_i)) { best = v; best_i = c; }
  };
  const bool vec = (ld % 4) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  const int n4 = vec ? cols / 4 : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  int i = threadIdx.x;
  for (; i + 3 * static_cast<int>(blockDim.x) < n4; i += 4 * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x4 + i + u * blockDim.x);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = (i + u * blockDim.x) * 4;
      take(v[u].x, c); take(v[u].y, c + 1); take(v[u].z, c + 2); take(v[u].w, c + 3);
    }
  }
  for (; i < n4; i += blockDim.x) {
    const float4 v = __ldg(x4 + i);
    take(v.x, i * 4); take(v.y, i * 4 + 1); take(v.z, i * 4 + 2); take(v.w, i * 4 + 3);
  }
  for (int c = n4 * 4 + threadIdx.x; c < cols; c += blockDim.x) take(x[c], c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = best_i; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? s_val[threadIdx.x] : -INFINITY;
    best_i = threadIdx.x < nw ? s_idx[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (threadIdx.x == 0) out[row] = best_i;
  }
}

}  // namespace opsg

using namespace opsg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int opsg_pair_mask_bits(const int32_t* pan, int pan_h, int pan_w, int img_h, int img_w, int pad_h, int pad_w,
                                   int tok_h, int tok_w, const int32_t* obj_ids, int num_objects, uint32_t* bits_out,
                                   int words, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(pan && obj_ids && bits_out, "pair_mask_bits: null pointer");
  OPSG_CHECK_ARG(pan_h > 0 && pan_w > 0 && img_h > 0 && img_w > 0 && tok_h > 0 && tok_w > 0 && num_objects > 0,
                 "pair_mask_bits: bad shape");
  OPSG_CHECK_ARG(pad_h >= img_h && pad_w >= img_w, "pair_mask_bits: pad_shape smaller than img_shape");
  OPSG_CHECK_ARG(words * 32 >= tok_h * tok_w, "pair_mask_bits: words too small for %d tokens", tok_h * tok_w);
  const long long threads = static_cast<long long>(num_objects) * words * 32;
  launch_kernel(pair_mask_bits_kernel, ceil_div(threads, 256), 256, 0, ST(stream), pan, pan_h, pan_w, img_h, img_w, pad_h, pad_w,
                                                                        tok_h, tok_w, obj_ids, num_objects, bits_out, words);
  OPSG_CHECK_LAUNCH("pair_mask_bits_kernel");
  return OPSG_OK;
}

extern "C" int opsg_patch_im2col(const float* feat, int channels, int h, int w, int patch, opsg_bf16* out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(feat && out, "patch_im2col: null pointer");
  OPSG_CHECK_ARG(patch > 0 && patch % 8 == 0 && h >= patch && w >= patch, "patch_im2col: patch must be a multiple of 8");
  const int th = h / patch, tw = w / patch;
  const long long total = static_cast<long long>(channels) * (th * patch) * ((tw * patch) / 8);
  launch_kernel(patch_im2col_kernel, ceil_div(total, 256), 256, 0, ST(stream), feat, channels, h, w, patch, th, tw,
                                                                    reinterpret_cast<__nv_bfloat16*>(out));
  OPSG_CHECK_LAUNCH("patch_im2col_kernel");
  return OPSG_OK;
}

extern "C" int opsg_cast_f32_bf16(const float* in, int ld_in, opsg_bf16* out, int ld_out, int rows, int cols, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(in && out && rows > 0 && cols > 0, "cast_f32_bf16: bad argument");
  launch_kernel(cast_f32_bf16_kernel, ceil_div(static_cast<long long>(rows) * cols, 256), 256, 0, ST(stream), 
      in, ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  OPSG_CHECK_LAUNCH("cast_f32_bf16_kernel");
  return OPSG_OK;
}

extern "C" int opsg_init_rows_f32(float* out, int ld_out, const float* row, int rows, int cols, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(out && rows > 0 && cols > 0, "init_rows_f32: bad argument");
  launch_kernel(init_rows_f32_kernel, ceil_div(static_cast<long long>(rows) * cols, 256), 256, 0, ST(stream), out, ld_out, row, rows, cols);
  OPSG_CHECK_LAUNCH("init_rows_f32_kernel");
  return OPSG_OK;
}

// rows [nq, B*nq) of h = copies of rows [0, nq): LN(query tokens) is the same for every pair (v4:158-159 expands it)
__global__ void __launch_bounds__(256) qformer_broadcast_rows_kernel(__nv_bfloat16* __restrict__ h, int nq, int B, int d) {
  pdl_wait_then_trigger();
  const int vec = d / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B - 1) * nq * vec;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % vec);
  const long long r = idx / vec + nq;                    // destination row
  const int q = static_cast<int>(r % nq);
  reinterpret_cast<uint4*>(h + r * d)[c] = __ldg(reinterpret_cast<const uint4*>(h + static_cast<size_t>(q) * d) + c);
}

extern "C" int opsg_qformer_embed_ln(const float* query, int n_query, const int32_t* input_ids, int B, int T,
                                     const float* word_emb, int vocab, const float* pos_emb, const float* gamma,
                                     const float* beta, float eps, int d, opsg_bf16* h_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(query && word_emb && pos_emb && gamma && beta && h_out, "qformer_embed_ln: null pointer");
  OPSG_CHECK_ARG(T == 0 || input_ids, "qformer_embed_ln: null input_ids");
  OPSG_CHECK_ARG(d % 8 == 0 && d <= 1024 && B > 0 && n_query > 0 && T >= 0, "qformer_embed_ln: bad shape (d=%d)", d);
  const long long items = n_query + static_cast<long long>(B) * T;
  launch_kernel(qformer_embed_ln_kernel, ceil_div(items * 32, 256), 256, 0, ST(stream), query, n_query, input_ids, B, T, word_emb, vocab,
                                                                           pos_emb, gamma, beta, eps, d,
                                                                           reinterpret_cast<__nv_bfloat16*>(h_out));
  OPSG_CHECK_LAUNCH("qformer_embed_ln_kernel");
  if (B > 1) {
    const long long total = static_cast<long long>(B - 1) * n_query * (d / 8);
    launch_kernel(qformer_broadcast_rows_kernel, ceil_div(total, 256), 256, 0, ST(stream),
                  reinterpret_cast<__nv_bfloat16*>(h_out), n_query, B, d);
    OPSG_CHECK_LAUNCH("qformer_broadcast_rows_kernel");
  }
  return OPSG_OK;
}

// Few rows (LLM decode: 100 rows of 2560 per image, 800 with eight images stacked): one CTA per row instead of one warp per row -- 256 threads issue the row's
// loads (x, gamma, beta) at once, so the kernel is one memory round trip deep instead of ten chunks per lane.
template <int CH>                                         // CH x 256 x 8 columns max
__global__ void __launch_bounds__(256) layernorm_row_cta_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps,
                                                                __nv_bfloat16* __restrict__ y, int cols) {
  pdl_wait_then_trigger();
  __shared__ float red[2][8];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const __nv_bfloat16* xr = x + static_cast<size_t>(blockIdx.x) * cols;
  float v[CH][8], g[CH][8], b[CH][8];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c0 = (i * 256 + t) * 8;
    if (c0 < cols) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + c0));
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
      v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
      v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
      g[i][0] = g0.x; g[i][1] = g0.y; g[i][2] = g0.z; g[i][3] = g0.w; g[i][4] = g1.x; g[i][5] = g1.y; g[i][6] = g1.z; g[i][7] = g1.w;
      b[i][0] = b0.x; b[i][1] = b0.y; b[i][2] = b0.z; b[i][3] = b0.w; b[i][4] = b1.x; b[i][5] = b1.y; b[i][6] = b1.z; b[i][7] = b1.w;
    }
  }
  auto block_sum = [&](float s, int slot) {
    s = warp_sum(s);
    if (lane == 0) red[slot][warp] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[slot][w];
    return tot;
  };
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
    if ((i * 256 + t) * 8 < cols) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  const float mean = block_sum(s, 0) / cols;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
    if ((i * 256 + t) * 8 < cols) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
    }
  const float rstd = rsqrtf(block_sum(ss, 1) / cols + eps);
  __nv_bfloat16* yr = y + static_cast<size_t>(blockIdx.x) * cols;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c0 = (i * 256 + t) * 8;
    if (c0 < cols) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[i][j] + b[i][j];
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + c0) = u;
    }
  }
}

extern "C" int opsg_layernorm_bf16(const opsg_bf16* x, const float* gamma, const float* beta, float eps, opsg_bf16* y,
                                   int rows, int cols, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(x && gamma && beta && y, "layernorm: null pointer");
  OPSG_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 8192, "layernorm: cols=%d unsupported (multiple of 8, <= 8192)", cols);
  const int grid = ceil_div(static_cast<long long>(rows) * 32, 256);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  // many rows of <= 1024 columns (the Q-Former): shared-memory ring fed by bulk copies
  if (rows >= 4096 && cols <= 1024 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0) {
    const int total_slabs = ceil_div(rows, kLnRingRows);
    const int smem = kLnRingStages * kLnRingRows * cols * 2;
    int spc = ceil_div(total_slabs, opsg_num_sms() * 2);      // two CTAs per SM (48 KB of ring each at 768 columns), one wave
    if (spc < kLnRingStages) spc = kLnRingStages;
    const int ctas = ceil_div(total_slabs, spc);
    static bool configured_dev[64] = {};
    bool& configured = configured_dev[device_slot()];
    if (!configured) {
      rc = check_cuda(cudaFuncSetAttribute(layernorm_ring_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute(layernorm_ring<3, true>)");
      if (rc) return rc;
      rc = check_cuda(cudaFuncSetAttribute(layernorm_ring_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute(layernorm_ring<3, false>)");
      if (rc) return rc;
      rc = check_cuda(cudaFuncSetAttribute(layernorm_ring_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024), "cudaFuncSetAttribute(layernorm_ring<4, false>)");
      if (rc) return rc;
      configured = true;
    }
    if (cols == 768) launch_kernel(layernorm_ring_kernel<3, true>, ctas, 256, smem, ST(stream), xp, gamma, beta, eps, yp, rows, cols, spc);
    else if (cols < 768) launch_kernel(layernorm_ring_kernel<3, false>, ctas, 256, smem, ST(stream), xp, gamma, beta, eps, yp, rows, cols, spc);
    else launch_kernel(layernorm_ring_kernel<4, false>, ctas, 256, smem, ST(stream), xp, gamma, beta, eps, yp, rows, cols, spc);
    OPSG_CHECK_LAUNCH("layernorm_ring_kernel");
    return OPSG_OK;
  }
  if (cols > 4096)
    launch_kernel(layernorm_row_cta_kernel<4>, rows, 256, 0, ST(stream), xp, gamma, beta, eps, yp, cols);
  else if ((rows <= 4096 && cols >= 1024) || cols > 256 * kMaxChunks)
    launch_kernel(layernorm_row_cta_kernel<2>, rows, 256, 0, ST(stream), xp, gamma, beta, eps, yp, cols);
  else if (cols <= 1024) launch_kernel(layernorm_bf16_kernel<4>, grid, 256, 0, ST(stream), xp, gamma, beta, eps, yp, rows, cols);
  else launch_kernel(layernorm_bf16_kernel<kMaxChunks>, grid, 256, 0, ST(stream), xp, gamma, beta, eps, yp, rows, cols);
  OPSG_CHECK_LAUNCH("layernorm_bf16_kernel");
  return OPSG_OK;
}

extern "C" int opsg_exist_filter_topk(const opsg_bf16* x, int ld_x, int B, int d, const float* w, const float* b,
                                      float threshold, int k, float* logits_out, float* probs_out, uint8_t* mask_out,
                                      int32_t* topk_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(x && w && b && logits_out, "exist_filter_topk: null pointer");
  OPSG_CHECK_ARG(B > 0 && B <= 65536 && d % 8 == 0 && ld_x % 8 == 0, "exist_filter_topk: bad shape");
  OPSG_CHECK_ARG(k >= 0 && k <= B && (k == 0 || topk_out), "exist_filter_topk: bad k");
  OPSG_CHECK_ARG(threshold > 0.f && threshold < 1.f, "exist_filter_topk: threshold must be in (0,1)");
  const float logit_thr = logf(threshold / (1.f - threshold));
  launch_kernel(exist_logits_kernel, ceil_div(static_cast<long long>(B) * 32, 256), 256, 0, ST(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), ld_x, B, d, w, b, logit_thr, logits_out, probs_out, mask_out);
  OPSG_CHECK_LAUNCH("exist_logits_kernel");
  if (k > 0) {
    launch_kernel(topk_rank_kernel, ceil_div(B, 32), 256, 0, ST(stream), logits_out, B, k, topk_out);
    OPSG_CHECK_LAUNCH("topk_rank_kernel");
  }
  return OPSG_OK;
}

extern "C" int opsg_gather_rows_bf16(const opsg_bf16* src, int row_elems, const int32_t* idx, int n_rows, opsg_bf16* out,
                                     void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(src && idx && out && n_rows > 0 && row_elems > 0 && row_elems % 8 == 0, "gather_rows: bad argument");
  const int row_vec = row_elems / 8;
  launch_kernel(gather_rows_bf16_kernel, ceil_div(static_cast<long long>(n_rows) * row_vec, 256), 256, 0, ST(stream), 
      reinterpret_cast<const uint4*>(src), row_vec, idx, n_rows, reinterpret_cast<uint4*>(out));
  OPSG_CHECK_LAUNCH("gather_rows_bf16_kernel");
  return OPSG_OK;
}

extern "C" int opsg_embed_gather(const opsg_bf16* table, int d, const int32_t* ids, const opsg_bf16* pos_table,
                                 const int32_t* pos, int n_rows, opsg_bf16* out, int ld_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(table && ids && out && n_rows > 0 && d % 8 == 0 && ld_out % 8 == 0, "embed_gather: bad argument");
  OPSG_CHECK_ARG(!pos_table || pos, "embed_gather: pos_table without pos");
  launch_kernel(embed_gather_kernel, ceil_div(static_cast<long long>(n_rows) * (d / 8), 256), 256, 0, ST(stream), 
      reinterpret_cast<const __nv_bfloat16*>(table), d, ids, reinterpret_cast<const __nv_bfloat16*>(pos_table), pos, n_rows,
      reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  OPSG_CHECK_LAUNCH("embed_gather_kernel");
  return OPSG_OK;
}

extern "C" int opsg_llm_build_prefix(const opsg_bf16* proj, int proj_rows_per_seq, int proj_row0, int n_prefix,
                                     const opsg_bf16* table, const int32_t* ids, int T, const opsg_bf16* pos_table,
                                     const int32_t* pos, int nseq, int d, opsg_bf16* out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(out && nseq > 0 && d > 0 && d % 8 == 0 && n_prefix >= 0 && T >= 0 && n_prefix + T > 0,
                 "llm_build_prefix: bad shape");
  OPSG_CHECK_ARG((n_prefix == 0 || proj) && (T == 0 || (table && ids)), "llm_build_prefix: null pointer");
  OPSG_CHECK_ARG(proj_row0 >= 0 && proj_row0 + n_prefix <= proj_rows_per_seq, "llm_build_prefix: prefix rows out of range");
  OPSG_CHECK_ARG(!pos_table || pos, "llm_build_prefix: pos_table without pos");
  const long long total = static_cast<long long>(nseq) * (n_prefix + T) * (d / 8);
  launch_kernel(llm_build_prefix_kernel, ceil_div(total, 256), 256, 0, ST(stream), 
      reinterpret_cast<const __nv_bfloat16*>(proj), proj_rows_per_seq, proj_row0, n_prefix,
      reinterpret_cast<const __nv_bfloat16*>(table), ids, T, reinterpret_cast<const __nv_bfloat16*>(pos_table), pos, nseq, d,
      reinterpret_cast<__nv_bfloat16*>(out));
  OPSG_CHECK_LAUNCH("llm_build_prefix_kernel");
  return OPSG_OK;
}

extern "C" int opsg_argmax_rows(const float* logits, int ld, int rows, int cols, int32_t* out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(logits && out && rows > 0 && cols > 0 && ld >= cols, "argmax_rows: bad argument");
  launch_kernel(argmax_rows_kernel, rows, 1024, 0, ST(stream), logits, ld, rows, cols, out);
  OPSG_CHECK_LAUNCH("argmax_rows_kernel");
  return OPSG_OK;
}
