// K4 / K10a fallbacks and K10b — small-sequence multi-head attention on warp-level mma.sync.
//
//  * `qformer_self_attn_kernel`, `small_attn_kernel`: the FALLBACKS of the tcgen05 + TMA kernels (self_attn_pairs.cu: Q-Former
//    self-attention over S = 33 + T <= 64 rows per pair; llm_prefill_attn.cu: LLM prefill) for the shapes those do not cover
//    (head_dim != 64 / more than 31 text tokens; prompts longer than 64 tokens or a non-zero first position).  One CTA per
//    (sequence, head), whole K / V / Q tile in shared memory, FlashAttention-2 style register-resident softmax.
//  * `decode_attn_tma_kernel`: K10b, the LLM decode step (ONE query row per sequence: not a tensor-core tile; HBM-bound on the
//    K / V cache) — TMA-staged head slices, persistent warps, 5.7 TB/s at 800 sequences (see its own header below);
//    `decode_attn_kernel` is its register-staged fallback for unaligned operands / contexts beyond 256 keys.
#include "common.cuh"
#include "host_util.h"

namespace opsg {

struct SmallAttnParams {
  int mode;                       // 0 = Q-Former split layout, 1 = LLM static KV cache
  // mode 0
  const __nv_bfloat16* qkv;       // [R, 3*d_model]
  const __nv_bfloat16* qkv_shared;// [n_query, 3*d_model] or NULL: q/k/v of the query rows when they are the same for every pair
  const int32_t* text_mask;       // [B, T]
  int B, n_query, T, text_queries;
  // mode 1
  const __nv_bfloat16* q;         // [nseq*q_len, ld_q]
  const __nv_bfloat16* k_cache;   // [nseq, max_ctx, d_model]
  const __nv_bfloat16* v_cache;
  const uint8_t* key_mask;        // [nseq, max_ctx]
  int ld_q, max_ctx, q_len, q_pos0;
  int append_kv;                  // decode: q points at fused [q | k | v] rows; the new token's k / v are taken from there
                                  // (not from the cache) and written to the caches at position q_pos0 by this kernel
  // common
  __nv_bfloat16* out;
  int ld_out, num_heads, d_model;
  float scale_log2e;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// four transposed 8x8 b16 matrices: lanes 0-7 / 8-15 / 16-23 / 24-31 give the row addresses of matrix 0 / 1 / 2 / 3
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// HD = head dim (multiple of 16), NK = padded key capacity (multiple of 16), 4 warps x 16 query rows per pass.
template <int HD, int NK>
__global__ void __launch_bounds__(128) small_attn_kernel(const SmallAttnParams p) {
  pdl_wait_then_trigger();
  constexpr int QS = HD + 8;      // padded row strides (elements) -> conflict-free fragment loads
  constexpr int KS = HD + 8;
  constexpr int VS = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_dyn);      // [64][QS]
  __nv_bfloat16* sK = sQ + 64 * QS;                                     // [NK][KS]
  __nv_bfloat16* sV = sK + NK * KS;                                     // [NK][VS] row-major (keys x head dims)
  uint8_t* sValid = reinterpret_cast<uint8_t*>(sV + NK * VS);           // [NK]

  const int seq = blockIdx.x, head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  int n_keys, n_q;
  if (p.mode == 0) { n_keys = p.n_query + p.T; n_q = p.text_queries ? n_keys : p.n_query; }
  else             { n_keys = p.q_pos0 + p.q_len; n_q = p.q_len; }

  auto qformer_row = [&](int i) -> size_t {
    return i < p.n_query ? static_cast<size_t>(seq) * p.n_query + i
                         : static_cast<size_t>(p.B) * p.n_query + static_cast<size_t>(seq) * p.T + (i - p.n_query);
  };

  auto qformer_qkv_row = [&](int i) -> const __nv_bfloat16* {
    return (p.qkv_shared && i < p.n_query) ? p.qkv_shared + static_cast<size_t>(i) * (3 * p.d_model)
                                           : p.qkv + qformer_row(i) * (3 * p.d_model);
  };
  // ---- stage K, V^T and key validity -------------------------------------------------------------
  constexpr int VEC = HD / 8;
  for (int idx = threadIdx.x; idx < NK * VEC; idx += blockDim.x) {
    const int key = idx / VEC, v8 = idx % VEC;
    uint4 ku = make_uint4(0, 0, 0, 0), vu = make_uint4(0, 0, 0, 0);
    if (key < n_keys) {
      if (p.mode == 0) {
        const __nv_bfloat16* base = qformer_qkv_row(key) + head * HD + v8 * 8;
        ku = __ldg(reinterpret_cast<const uint4*>(base + p.d_model));
        vu = __ldg(reinterpret_cast<const uint4*>(base + 2 * p.d_model));
      } else {
        const size_t off = (static_cast<size_t>(seq) * p.max_ctx + key) * p.d_model + head * HD + v8 * 8;
        ku = __ldg(reinterpret_cast<const uint4*>(p.k_cache + off));
        vu = __ldg(reinterpret_cast<const uint4*>(p.v_cache + off));
      }
    }
    *reinterpret_cast<uint4*>(sK + key * KS + v8 * 8) = ku;
    *reinterpret_cast<uint4*>(sV + key * VS + v8 * 8) = vu;
  }
  for (int key = threadIdx.x; key < NK; key += blockDim.x) {
    bool ok = key < n_keys;
    if (ok) {
      if (p.mode == 0) ok = key < p.n_query || p.text_mask[static_cast<size_t>(seq) * p.T + (key - p.n_query)] != 0;
      else ok = p.key_mask[static_cast<size_t>(seq) * p.max_ctx + key] != 0;
    }
    sValid[key] = ok ? 1 : 0;
  }

  auto stage_q = [&](int q0) {
    for (int idx = threadIdx.x; idx < 64 * VEC; idx += blockDim.x) {
      const int r = idx / VEC, v8 = idx % VEC;
      const int qi = q0 + r;
      uint4 qu = make_uint4(0, 0, 0, 0);
      if (qi < n_q) {
        if (p.mode == 0) qu = __ldg(reinterpret_cast<const uint4*>(qformer_qkv_row(qi) + head * HD + v8 * 8));
        else qu = __ldg(reinterpret_cast<const uint4*>(p.q + (static_cast<size_t>(seq) * p.q_len + qi) * p.ld_q + head * HD + v8 * 8));
      }
      *reinterpret_cast<uint4*>(sQ + r * QS + v8 * 8) = qu;
    }
  };
  stage_q(0);       // the first 64 query rows travel with K / V: one global round trip, one barrier

  for (int q0 = 0; q0 < n_q; q0 += 64) {
    if (q0 > 0) {
      __syncthreads();
      stage_q(q0);
    }
    __syncthreads();
    if (q0 + warp * 16 >= n_q) continue;   // warp-uniform; no further block-wide barriers below in this pass

    // ---- S = Q K^T ---------------------------------------------------------------------------------
    float s[NK / 8][4];
#pragma unroll
    for (int n = 0; n < NK / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
    const __nv_bfloat16* qw = sQ + (warp * 16) * QS;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(qw + g * QS + kk * 16 + 2 * t);
      a[1] = *reinterpret_cast<const uint32_t*>(qw + (g + 8) * QS + kk * 16 + 2 * t);
      a[2] = *reinterpret_cast<const uint32_t*>(qw + g * QS + kk * 16 + 8 + 2 * t);
      a[3] = *reinterpret_cast<const uint32_t*>(qw + (g + 8) * QS + kk * 16 + 8 + 2 * t);
#pragma unroll
      for (int n = 0; n < NK / 8; ++n) {
        const __nv_bfloat16* kr = sK + (n * 8 + g) * KS + kk * 16 + 2 * t;
        mma_bf16_16816(s[n], a, *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
      }
    }
    // ---- mask + softmax (rows g and g+8 of this warp's 16) -------------------------------------------
    const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;
    const int lim0 = (p.mode == 1) ? p.q_pos0 + qi0 : 0x7fffffff;   // causal: key <= absolute query position
    const int lim1 = (p.mode == 1) ? p.q_pos0 + qi1 : 0x7fffffff;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NK / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = n * 8 + 2 * t + e;
        const bool ok = sValid[key] != 0;
        if (!(ok && key <= lim0)) s[n][e] = -INFINITY;
        if (!(ok && key <= lim1)) s[n][2 + e] = -INFINITY;
        m0 = fmaxf(m0, s[n][e]);
        m1 = fmaxf(m1, s[n][2 + e]);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    if (m0 == -INFINITY) m0 = 0.f;
    if (m1 == -INFINITY) m1 = 0.f;
    const float ms0 = m0 * p.scale_log2e, ms1 = m1 * p.scale_log2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int n = 0; n < NK / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[n][e] = ex2f(fmaf(s[n][e], p.scale_log2e, -ms0));
        s[n][2 + e] = ex2f(fmaf(s[n][2 + e], p.scale_log2e, -ms1));
        l0 += s[n][e];
        l1 += s[n][2 + e];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V --------------------------------------------------------------------------------------
    float o[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < NK / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      // B fragments of V[kk*16 .. +16][n*8 .. +8] for two adjacent n-tiles per ldmatrix.x4.trans:
      // lanes 0-15 address the 16 key rows at column n0, lanes 16-31 the same rows at column n0 + 8
      const __nv_bfloat16* vrow = sV + (kk * 16 + (lane & 15)) * VS + ((lane >> 4) << 3);
#pragma unroll
      for (int n = 0; n < HD / 8; n += 2) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, vrow + n * 8);
        mma_bf16_16816(o[n], a, b[0], b[1]);
        mma_bf16_16816(o[n + 1], a, b[2], b[3]);
      }
    }
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    auto out_row = [&](int qi) -> __nv_bfloat16* {
      const size_t row = (p.mode == 0) ? qformer_row(qi) : static_cast<size_t>(seq) * p.q_len + qi;
      return p.out + row * p.ld_out + head * HD;
    };
    if (qi0 < n_q) {
      __nv_bfloat16* d = out_row(qi0);
#pragma unroll
      for (int n = 0; n < HD / 8; ++n)
        *reinterpret_cast<uint32_t*>(d + n * 8 + 2 * t) = pack_bf16x2(o[n][0] * inv0, o[n][1] * inv0);
    }
    if (qi1 < n_q) {
      __nv_bfloat16* d = out_row(qi1);
#pragma unroll
      for (int n = 0; n < HD / 8; ++n)
        *reinterpret_cast<uint32_t*>(d + n * 8 + 2 * t) = pack_bf16x2(o[n][2] * inv1, o[n][3] * inv1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 (Q-Former shape: head_dim 64, <= 64 rows per pair) — persistent, software-pipelined variant.
// Profile of the one-CTA-per-(pair, head) kernel above (profiles/r1_launches_*.md): 355 us per cfg2 image for 480 MB of
// qkv-in / ctx-out traffic (74 us at HBM speed) — every CTA pays two dependent global round trips before it computes.
// Here a CTA walks a contiguous range of (pair, head) problems (consecutive heads of one pair = the same 49 rows of qkv,
// so DRAM pages are fully used), cp.async fills stage i+1 while stage i is computed, and the context rows leave through
// shared memory as full 128-byte rows.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;           // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kQfHD = 64, kQfNK = 64, kQfRS = kQfHD + 8;                 // padded row stride (elements)
constexpr int kQfStageElems = 3 * 64 * kQfRS;                             // Q | K | V tiles of one problem
constexpr int kQfStageBytes = kQfStageElems * 2 + 64;                     // + key validity bytes
constexpr int kQfStages = 2;

__global__ void __launch_bounds__(128) qformer_self_attn_kernel(const SmallAttnParams p, int num_problems) {
  pdl_wait_then_trigger();
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_keys = p.n_query + p.T;
  const int n_q = p.text_queries ? n_keys : p.n_query;
  const long long pb = static_cast<long long>(num_problems) * blockIdx.x / gridDim.x;
  const long long pe = static_cast<long long>(num_problems) * (blockIdx.x + 1) / gridDim.x;

  auto row_of = [&](int pair, int i) -> size_t {
    return i < p.n_query ? static_cast<size_t>(pair) * p.n_query + i
                         : static_cast<size_t>(p.B) * p.n_query + static_cast<size_t>(pair) * p.T + (i - p.n_query);
  };
  auto stage_ptr = [&](int st) { return reinterpret_cast<__nv_bfloat16*>(smem_dyn + st * kQfStageBytes); };
  auto issue_loads = [&](long long prob, int st) {
    const int pair = static_cast<int>(prob / p.num_heads), head = static_cast<int>(prob % p.num_heads);
    __nv_bfloat16* sQ = stage_ptr(st);
    __nv_bfloat16* sK = sQ + 64 * kQfRS;
    __nv_bfloat16* sV = sK + 64 * kQfRS;
    uint8_t* sValid = reinterpret_cast<uint8_t*>(sV + 64 * kQfRS);
    for (int idx = threadIdx.x; idx < 64 * 8; idx += 128) {
      const int r = idx >> 3, v8 = idx & 7;
      const bool ok = r < n_keys;
      const int rr = ok ? r : 0;
      const __nv_bfloat16* base = ((p.qkv_shared && rr < p.n_query) ? p.qkv_shared + static_cast<size_t>(rr) * (3 * p.d_model)
                                                                    : p.qkv + row_of(pair, rr) * (3 * p.d_model)) +
                                  head * kQfHD + v8 * 8;
      cp_async16(sQ + r * kQfRS + v8 * 8, base, ok && r < n_q);
      cp_async16(sK + r * kQfRS + v8 * 8, base + p.d_model, ok);
      cp_async16(sV + r * kQfRS + v8 * 8, base + 2 * p.d_model, ok);
    }
    if (threadIdx.x < 64) {
      const int key = threadIdx.x;
      bool ok = key < n_keys;
      if (ok && key >= p.n_query) ok = p.text_mask[static_cast<size_t>(pair) * p.T + (key - p.n_query)] != 0;
      sValid[key] = ok ? 1 : 0;
    }
    cp_async_commit();
  };

  if (pb < pe) issue_loads(pb, 0);
  for (long long prob = pb; prob < pe; ++prob) {
    const int st = static_cast<int>((prob - pb) & 1);
    if (prob + 1 < pe) {
      issue_loads(prob + 1, st ^ 1);          // stage st^1 was released by the barrier at the end of the previous iteration
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int pair = static_cast<int>(prob / p.num_heads), head = static_cast<int>(prob % p.num_heads);
    __nv_bfloat16* sQ = stage_ptr(st);
    const __nv_bfloat16* sK = sQ + 64 * kQfRS;
    const __nv_bfloat16* sV = sK + 64 * kQfRS;
    const uint8_t* sValid = reinterpret_cast<const uint8_t*>(sV + 64 * kQfRS);

    if (warp * 16 < n_q) {
      // ---- S = Q K^T ----
      float sc[kQfNK / 8][4];
#pragma unroll
      for (int n = 0; n < kQfNK / 8; ++n) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f; }
      const __nv_bfloat16* qw = sQ + (warp * 16) * kQfRS;
#pragma unroll
      for (int kk = 0; kk < kQfHD / 16; ++kk) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(qw + g * kQfRS + kk * 16 + 2 * t);
        a[1] = *reinterpret_cast<const uint32_t*>(qw + (g + 8) * kQfRS + kk * 16 + 2 * t);
        a[2] = *reinterpret_cast<const uint32_t*>(qw + g * kQfRS + kk * 16 + 8 + 2 * t);
        a[3] = *reinterpret_cast<const uint32_t*>(qw + (g + 8) * kQfRS + kk * 16 + 8 + 2 * t);
#pragma unroll
        for (int n = 0; n < kQfNK / 8; ++n) {
          const __nv_bfloat16* kr = sK + (n * 8 + g) * kQfRS + kk * 16 + 2 * t;
          mma_bf16_16816(sc[n], a, *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
        }
      }
      // ---- mask + softmax (rows g and g+8 of this warp's 16) ----
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < kQfNK / 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (sValid[n * 8 + 2 * t + e] == 0) { sc[n][e] = -INFINITY; sc[n][2 + e] = -INFINITY; }
          m0 = fmaxf(m0, sc[n][e]);
          m1 = fmaxf(m1, sc[n][2 + e]);
        }
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      if (m0 == -INFINITY) m0 = 0.f;
      if (m1 == -INFINITY) m1 = 0.f;
      const float ms0 = m0 * p.scale_log2e, ms1 = m1 * p.scale_log2e;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int n = 0; n < kQfNK / 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[n][e] = ex2f(fmaf(sc[n][e], p.scale_log2e, -ms0));
          sc[n][2 + e] = ex2f(fmaf(sc[n][2 + e], p.scale_log2e, -ms1));
          l0 += sc[n][e];
          l1 += sc[n][2 + e];
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // ---- O = P V ----
      float o[kQfHD / 8][4];
#pragma unroll
      for (int n = 0; n < kQfHD / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < kQfNK / 16; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
        a[1] = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
        a[2] = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
        a[3] = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
        const __nv_bfloat16* vrow = sV + (kk * 16 + (lane & 15)) * kQfRS + ((lane >> 4) << 3);
#pragma unroll
        for (int n = 0; n < kQfHD / 8; n += 2) {
          uint32_t b[4];
          ldmatrix_x4_trans(b, vrow + n * 8);
          mma_bf16_16816(o[n], a, b[0], b[1]);
          mma_bf16_16816(o[n + 1], a, b[2], b[3]);
        }
      }
      // ---- context rows -> this warp's own Q rows in shared memory -> 128-byte row stores ----
      const float inv0 = 1.f / l0, inv1 = 1.f / l1;
      __syncwarp();                               // all lanes are done reading this warp's Q rows
      __nv_bfloat16* ow = sQ + (warp * 16) * kQfRS;
#pragma unroll
      for (int n = 0; n < kQfHD / 8; ++n) {
        *reinterpret_cast<uint32_t*>(ow + g * kQfRS + n * 8 + 2 * t) = pack_bf16x2(o[n][0] * inv0, o[n][1] * inv0);
        *reinterpret_cast<uint32_t*>(ow + (g + 8) * kQfRS + n * 8 + 2 * t) = pack_bf16x2(o[n][2] * inv1, o[n][3] * inv1);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = i * 32 + lane, r = idx >> 3, v8 = idx & 7;
        const int qi = warp * 16 + r;
        if (qi < n_q)
          *reinterpret_cast<uint4*>(p.out + row_of(pair, qi) * p.ld_out + head * kQfHD + v8 * 8) =
              *reinterpret_cast<const uint4*>(ow + r * kQfRS + v8 * 8);
      }
    }
    __syncthreads();                              // stage st may be refilled by the next iteration's prefetch
  }
}

template <int HD, int NK>
static int launch_small_attn(const SmallAttnParams& p, int nseq, cudaStream_t st) {
  constexpr int smem = (64 * (HD + 8) + 2 * NK * (HD + 8)) * 2 + NK;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    int rc = check_cuda(cudaFuncSetAttribute(small_attn_kernel<HD, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                        "cudaFuncSetAttribute(small_attn)");
    if (rc) return rc;
    configured = true;
  }
  launch_kernel(small_attn_kernel<HD, NK>, dim3(nseq, p.num_heads), 128, smem, st, p);
  OPSG_CHECK_LAUNCH("small_attn_kernel");
  return OPSG_OK;
}

template <int HD>
static int dispatch_nk(const SmallAttnParams& p, int nseq, int n_keys, cudaStream_t st) {
  if (n_keys <= 64) return launch_small_attn<HD, 64>(p, nseq, st);
  if (n_keys <= 128) return launch_small_attn<HD, 128>(p, nseq, st);
  if (n_keys <= 256) return launch_small_attn<HD, 256>(p, nseq, st);
  return set_error(OPSG_E_UNSUPPORTED, "small attention: %d keys > 256 unsupported", n_keys);
}

static int dispatch_hd(const SmallAttnParams& p, int nseq, int n_keys, int head_dim, cudaStream_t st) {
  switch (head_dim) {
    case 64: return dispatch_nk<64>(p, nseq, n_keys, st);
    case 80: return dispatch_nk<80>(p, nseq, n_keys, st);
    case 128: return dispatch_nk<128>(p, nseq, n_keys, st);
    default: return set_error(OPSG_E_UNSUPPORTED, "small attention: head_dim %d unsupported (64, 80, 128)", head_dim);
  }
}

// ------------------------------------------------------------------------------------------------
// K10b — decode attention (q_len == 1): one warp per (sequence, head), no tensor cores (a 1 x ctx x head_dim problem is
// a pair of GEMVs).  HBM-bound: reads the sequence's K / V head slices once (ctx x head_dim x 2 x 2 bytes).
// A key's head slice (HD bf16) is covered by HD/8 lanes with one 16-byte load each; LPK = 8 or 16 lanes form a key
// group, so a warp handles 32/LPK keys per iteration and 16 iterations' loads are in flight at once (the first cut of
// this kernel walked the keys one at a time in the PV phase and took 37 us per launch, latency-bound).
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) decode_attn_kernel(const SmallAttnParams p) {
  pdl_wait_then_trigger();
  constexpr int kWarps = 8;
  constexpr int kMaxCtx = 256;
  constexpr int LPK = (HD / 8 <= 8) ? 8 : 16;              // lanes per key (power of two >= HD / 8)
  constexpr int KPW = 32 / LPK;                             // keys per warp iteration
  constexpr int CH = 16;                                    // iterations whose loads are issued back to back
  __shared__ float s_p[kWarps][kMaxCtx];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarps + warp;             // (sequence, head)
  const int nseq_heads = p.B;                               // field reused by the launcher: nseq * num_heads
  if (item >= nseq_heads) return;
  const int seq = item / p.num_heads, head = item % p.num_heads;
  const int ctx = p.q_pos0 + 1;                             // keys 0 .. q_pos0 (causal, the new token included)
  const int sub = lane % LPK, grp = lane / LPK;
  const bool active = sub < HD / 8;
  const __nv_bfloat16* kbase = p.k_cache + static_cast<size_t>(seq) * p.max_ctx * p.d_model + head * HD + sub * 8;
  const __nv_bfloat16* vbase = p.v_cache + static_cast<size_t>(seq) * p.max_ctx * p.d_model + head * HD + sub * 8;
  const uint8_t* kmask = p.key_mask + static_cast<size_t>(seq) * p.max_ctx;
  float q8[8];
  {
    uint4 qv = make_uint4(0, 0, 0, 0);
    if (active) qv = __ldg(reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(seq) * p.ld_q + head * HD + sub * 8));
    q8[0] = bf16_lo(qv.x); q8[1] = bf16_hi(qv.x); q8[2] = bf16_lo(qv.y); q8[3] = bf16_hi(qv.y);
    q8[4] = bf16_lo(qv.z); q8[5] = bf16_hi(qv.z); q8[6] = bf16_lo(qv.w); q8[7] = bf16_hi(qv.w);
  }
  // ---- scores ----------------------------------------------------------------------------------------
  for (int k0 = 0; k0 < ctx; k0 += CH * KPW) {
    uint4 kv[CH];
    bool ok[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int key = k0 + i * KPW + grp;
      // the K load does not wait for the mask byte: both are issued together, the mask only gates the score
      kv[i] = (key < ctx && active) ? __ldg(reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(key) * p.d_model))
                                    : make_uint4(0, 0, 0, 0);
      ok[i] = key < ctx && __ldg(kmask + min(key, p.max_ctx - 1)) != 0;
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float d = bf16_lo(kv[i].x) * q8[0] + bf16_hi(kv[i].x) * q8[1] + bf16_lo(kv[i].y) * q8[2] + bf16_hi(kv[i].y) * q8[3] +
                bf16_lo(kv[i].z) * q8[4] + bf16_hi(kv[i].z) * q8[5] + bf16_lo(kv[i].w) * q8[6] + bf16_hi(kv[i].w) * q8[7];
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      const int key = k0 + i * KPW + grp;
      if (sub == 0 && key < kMaxCtx) s_p[warp][key] = ok[i] ? d : -INFINITY;
    }
  }
  __syncwarp();
  // ---- softmax over the ctx scores (lane-strided) --------------------------------------------------------
  float sc[kMaxCtx / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < kMaxCtx / 32; ++r) {
    const int key = r * 32 + lane;
    sc[r] = key < ctx ? s_p[warp][key] : -INFINITY;
    mx = fmaxf(mx, sc[r]);
  }
  mx = warp_max(mx);
  if (mx == -INFINITY) mx = 0.f;
  float sum = 0.f;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < kMaxCtx / 32; ++r) {
    const float e = ex2f((sc[r] - mx) * p.scale_log2e);      // exp2(-inf) = 0 for masked / out-of-range keys
    // P is rounded to bf16 like the tensor-core path (and HF's bf16 softmax output) before multiplying V
    s_p[warp][r * 32 + lane] = __bfloat162float(__float2bfloat16(e));
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // ---- context ---------------------------------------------------------------------------------------
  float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < ctx; k0 += CH * KPW) {
    uint4 vv[CH];
    float pj[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int key = k0 + i * KPW + grp;
      pj[i] = key < ctx ? s_p[warp][key] : 0.f;
      vv[i] = (pj[i] != 0.f && active) ? __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(key) * p.d_model))
                                       : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      o8[0] = fmaf(pj[i], bf16_lo(vv[i].x), o8[0]); o8[1] = fmaf(pj[i], bf16_hi(vv[i].x), o8[1]);
      o8[2] = fmaf(pj[i], bf16_lo(vv[i].y), o8[2]); o8[3] = fmaf(pj[i], bf16_hi(vv[i].y), o8[3]);
      o8[4] = fmaf(pj[i], bf16_lo(vv[i].z), o8[4]); o8[5] = fmaf(pj[i], bf16_hi(vv[i].z), o8[5]);
      o8[6] = fmaf(pj[i], bf16_lo(vv[i].w), o8[6]); o8[7] = fmaf(pj[i], bf16_hi(vv[i].w), o8[7]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) o8[e] += __shfl_xor_sync(0xffffffffu, o8[e], o);
  }
  if (grp == 0 && active) {
    const float inv = 1.f / sum;
    uint4 u;
    u.x = pack_bf16x2(o8[0] * inv, o8[1] * inv); u.y = pack_bf16x2(o8[2] * inv, o8[3] * inv);
    u.z = pack_bf16x2(o8[4] * inv, o8[5] * inv); u.w = pack_bf16x2(o8[6] * inv, o8[7] * inv);
    *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(seq) * p.ld_out + head * HD + sub * 8) = u;
  }
}

// ------------------------------------------------------------------------------------------------
// K10b -- decode attention (one query token per sequence) over the static KV cache: one warp per (sequence, head) item.
// History (all measured on B200, 32 heads x 80): register-held scalar version 27 us per launch at 100 sequences; K / V head
// slices staged in shared memory by cp.async with scalar FMAs 15 us (166 us at 800 sequences: ~2700 warp instructions per item,
// issue-bound); the same staging with warp-level mma.sync arithmetic 145 us; the kernel below (TMA staging, persistent warps)
// 142 us = 3.7 TB/s of K / V.
//   S  = Q K^T : A = the query row broadcast to all 16 rows, B = K rows straight from ldmatrix (keys = n, dims = k); the
//                accumulator layout of S (keys along columns) is the A-fragment layout the PV product needs, so P never
//                leaves registers (FlashAttention-2's trick);
//   O  = P V   : B = V through ldmatrix.trans.  Every row of S / O is the same row; row-group g of the warp writes the
//                output n-tiles g and g + 8.
// Rows past the context (up to the next multiple of 16) are zero-filled so that 0 x garbage cannot produce NaN.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

// ------------------------------------------------------------------------------------------------
// Staging by TMA: one elected lane issues ONE tensor copy for the item's K slice and one for its V slice (box = [cached rows x head_dim] of the cache viewed as
// [nseq * max_ctx, d_model]; the tensor maps are built per launch because the box height is the context length).  ncu of the
// cp.async version at 800 sequences (profiles/r2_ncu_decode_attn.md): ~2000 warp instructions per item, 40 % of them the copy
// loop, 0.32 IPC with two warps per scheduler -> the LOAD ISSUE itself took ~3 us per item and DRAM sat at 46 %.
// head_dim 80: dense 160-byte rows (ldmatrix phases are 2-way bank conflicted, harmless); head_dim 64 / 128: 128-byte
// swizzled boxes of 64 dims (a dense 128 / 256-byte pitch would put all 8 rows of an ldmatrix phase on the same banks).
// ------------------------------------------------------------------------------------------------
template <int HD>
struct KvSmem {
  static constexpr bool kSwizzled = (HD % 64) == 0;
  // heads per warp item.  Two adjacent 160-byte head rows (head_dim 80) as one 320-byte piece were measured SLOWER (203 against
  // 142 us at 800 sequences): half as many warps fit, and a warp's serial mma.sync chain per head (~2.8 us), not the row-request
  // rate, is what bounds the kernel (profiles/r2_ncu_decode_attn.md)
  static constexpr int kHeads = 1;
  static constexpr int kC = HD / 8;                         // 16-byte chunks per head row
  static constexpr int kRowBytes = kHeads * HD * 2;
  // byte offset of chunk c (of the item's kHeads * kC) of row r inside a K (or V) buffer of `rows16` rows
  __device__ static __forceinline__ uint32_t at(int r, int c, int rows16) {
    if constexpr (kSwizzled) return (c >> 3) * rows16 * 128 + r * 128 + (((c & 7) ^ (r & 7)) << 4);
    else return r * kRowBytes + c * 16;
  }
};

// Persistent warps: the shared memory of an SM bounds the bytes in flight (8 warps x 26 KB), so a buffer should spend its time
// LOADING, not waiting for its warp's arithmetic: a warp walks items warp, warp + W, ... and re-issues the K copy of its NEXT
// item as soon as the scores of the current one are in registers, the V copy as soon as the context is -- single buffers, the
// next item's loads overlap the rest of the current item's work.
template <int HD, int NT>                                   // NT = key tiles of 16 the registers are sized for
__global__ void __launch_bounds__(128) decode_attn_tma_kernel(const __grid_constant__ CUtensorMap tm_k,
                                                              const __grid_constant__ CUtensorMap tm_v,
                                                              const SmallAttnParams p, int warps_per_cta, int per_warp) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bars[4][2];
  using L = KvSmem<HD>;
  constexpr int C = HD / 8, HW = L::kHeads, RB = L::kRowBytes;   // an item = HW adjacent heads of one sequence
  const int groups = p.num_heads / HW;                      // (host-checked: divisible)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = blockIdx.x * warps_per_cta + warp;     // items are (sequence, head group)
  const int stride = gridDim.x * warps_per_cta;
  const bool live = warp < warps_per_cta && first < p.B;    // p.B = nseq * num_heads (set by the launcher)
  if (live && lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  __syncwarp();
  pdl_wait_then_trigger();
  if (!live) return;
  const int ctx = p.q_pos0 + 1;                             // keys 0 .. q_pos0 (causal, the new token included)
  const int ctx16 = (ctx + 15) & ~15;
  const int cached = p.append_kv ? ctx - 1 : ctx;           // rows that come from the caches
  const uint32_t bytes = static_cast<uint32_t>(cached) * RB;
  uint8_t* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u) + static_cast<size_t>(warp) * per_warp;
  uint8_t* sK = base;
  uint8_t* sV = sK + ctx16 * RB;
  uint8_t* sQ = sV + ctx16 * RB;
  uint64_t* bar_k = &bars[warp][0];
  uint64_t* bar_v = &bars[warp][1];
  auto issue = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int item) {      // lane 0 only
    const int seq = item / groups, head = (item - seq * groups) * HW;
    mbar_expect_tx(bar, bytes);
    if constexpr (L::kSwizzled) {
#pragma unroll
      for (int r = 0; r < HD / 64; ++r) tma_load_2d(dst + r * ctx16 * 128, tm, bar, head * HD + r * 64, seq * p.max_ctx);
    } else {
      tma_load_2d(dst, tm, bar, head * HD, seq * p.max_ctx);
    }
  };
  if (lane == 0 && cached > 0) {
    issue(sK, &tm_k, bar_k, first);
    issue(sV, &tm_v, bar_v, first);
  }
  // V rows past the context stay zero for the whole launch (P is zero there, 0 x garbage must not be NaN)
  for (int i = lane; i < (ctx16 - ctx) * C * HW; i += 32) {
    const int r = ctx + i / (C * HW), c = i % (C * HW);
    *reinterpret_cast<uint4*>(sV + L::at(r, c, ctx16)) = make_uint4(0u, 0u, 0u, 0u);
  }
  const int quad = lane & 3, grp = lane >> 2;
  // ldmatrix rows of this lane.  K: matrix lane/8 = (keys +0 dims +0) (keys +0 dims +8) (keys +8 dims +0) (keys +8 dims +8);
  // V (.trans): (keys +0 dims +0) (keys +8 dims +0) (keys +0 dims +8) (keys +8 dims +8)
  const int k_r = ((lane >> 4) & 1) * 8 + (lane & 7), k_c = (lane >> 3) & 1;
  const int v_r = ((lane >> 3) & 1) * 8 + (lane & 7), v_c = (lane >> 4) & 1;
  uint32_t phase = 0;
  // The item's own small loads (query row, new token's k / v row, key-mask bytes: lane l holds keys l, l + 32, ...) are fetched
  // ONE ITEM AHEAD into registers and consumed at the top of the next iteration: their ~1 us of global latency then runs under
  // the current item's arithmetic instead of in front of it (ncu: long-scoreboard stalls were the largest stall class).
  uint4 rq = make_uint4(0u, 0u, 0u, 0u), rk = rq, rv = rq;
  uint32_t rmask[NT / 2];
  auto prefetch = [&](int it) {
    const int sq = it / groups, hd = (it - sq * groups) * HW;
    const __nv_bfloat16* nk = p.q + static_cast<size_t>(sq) * p.ld_q + hd * HD;          // + d_model: k, + 2 d_model: v
    if (lane < C * HW) {
      rq = __ldg(reinterpret_cast<const uint4*>(nk + lane * 8));
      if (p.append_kv) {
        rk = __ldg(reinterpret_cast<const uint4*>(nk + p.d_model + lane * 8));
        rv = __ldg(reinterpret_cast<const uint4*>(nk + 2 * p.d_model + lane * 8));
      }
    }
    const uint8_t* km = p.key_mask + static_cast<size_t>(sq) * p.max_ctx;
#pragma unroll
    for (int r = 0; r < NT / 2; ++r) {
      const int key = r * 32 + lane;
      rmask[r] = key < ctx ? static_cast<uint32_t>(__ldg(km + key)) : 0u;
    }
  };
  prefetch(first);

  for (int item = first; item < p.B; item += stride, phase ^= 1u) {
    const int seq = item / groups, head = (item - seq * groups) * HW;
    const int next = item + stride;
    __nv_bfloat16* kbase = const_cast<__nv_bfloat16*>(p.k_cache) + static_cast<size_t>(seq) * p.max_ctx * p.d_model + head * HD;
    __nv_bfloat16* vbase = const_cast<__nv_bfloat16*>(p.v_cache) + static_cast<size_t>(seq) * p.max_ctx * p.d_model + head * HD;
    // the query row and the new token's k / v (prefetched): into shared memory (rows the tensor copies do not touch) AND into
    // the caches (replaces a kv_append launch)
    if (lane < C * HW) {
      *reinterpret_cast<uint4*>(sQ + lane * 16) = rq;
      if (p.append_kv) {
        *reinterpret_cast<uint4*>(sK + L::at(ctx - 1, lane, ctx16)) = rk;
        *reinterpret_cast<uint4*>(sV + L::at(ctx - 1, lane, ctx16)) = rv;
        const size_t off = static_cast<size_t>(ctx - 1) * p.d_model + lane * 8;
        *reinterpret_cast<uint4*>(kbase + off) = rk;
        *reinterpret_cast<uint4*>(vbase + off) = rv;
      }
    }
    // key validity: 32 keys per ballot word, then this lane's score columns (keys t*16 + quad*2 + {0, 1} and + 8)
    uint32_t vw[NT / 2];
#pragma unroll
    for (int r = 0; r < NT / 2; ++r) vw[r] = __ballot_sync(0xffffffffu, rmask[r] != 0u);
    uint32_t valid[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      valid[t] = 0;
      if (t * 16 < ctx) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int bit = (t & 1) * 16 + (j >> 1) * 8 + quad * 2 + (j & 1);     // key = t * 16 + ... = (t >> 1) * 32 + bit
          if ((vw[t >> 1] >> bit) & 1u) valid[t] |= 1u << j;
        }
      }
    }
    if (next < p.B) prefetch(next);

    // ---- S = q K^T --------------------------------------------------------------------------------------
    __syncwarp();
    if (cached > 0) mbar_wait(bar_k, phase);
    uint32_t pa[HW][NT][2];                                 // P as bf16 A fragments (rounded like the tensor-core prefill path)
    float inv[HW];
#pragma unroll
    for (int hh = 0; hh < HW; ++hh) {
    uint32_t qa[HD / 16][2];                                // A fragments of the broadcast query row: a0 = a1, a2 = a3
#pragma unroll
    for (int kb = 0; kb < HD / 16; ++kb) {
      qa[kb][0] = *reinterpret_cast<const uint32_t*>(sQ + (hh * HD + kb * 16 + quad * 2) * 2);
      qa[kb][1] = *reinterpret_cast<const uint32_t*>(sQ + (hh * HD + kb * 16 + 8 + quad * 2) * 2);
    }
    float sc[NT][4];                                        // [t][0..1]: keys t*16 + quad*2 + {0,1}; [t][2..3]: the same + 8
    float mx = -INFINITY;
    // two key tiles per step: four independent accumulator chains in flight (a lone tile leaves two chains of depth HD / 16 and
    // the warp waits out every mma's latency).  The second tile of the last pair may lie past the context: its rows are then
    // whatever follows in the warp's own buffer, and its scores are masked (valid bits are zero there).
#pragma unroll
    for (int t = 0; t < NT; t += 2) {
      if (t * 16 < ctx) {
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f}, c3[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kb = 0; kb < HD / 16; ++kb) {
          uint32_t b[4], b2[4];
          ldmatrix_x4(b, sK + L::at(t * 16 + k_r, hh * C + kb * 2 + k_c, ctx16));
          ldmatrix_x4(b2, sK + L::at(t * 16 + 16 + k_r, hh * C + kb * 2 + k_c, ctx16));
          const uint32_t a[4] = {qa[kb][0], qa[kb][0], qa[kb][1], qa[kb][1]};
          mma_bf16_16816(c0, a, b[0], b[1]);
          mma_bf16_16816(c1, a, b[2], b[3]);
          mma_bf16_16816(c2, a, b2[0], b2[1]);
          mma_bf16_16816(c3, a, b2[2], b2[3]);
        }
        sc[t][0] = (valid[t] & 1u) ? c0[0] : -INFINITY; sc[t][1] = (valid[t] & 2u) ? c0[1] : -INFINITY;
        sc[t][2] = (valid[t] & 4u) ? c1[0] : -INFINITY; sc[t][3] = (valid[t] & 8u) ? c1[1] : -INFINITY;
        sc[t + 1][0] = (valid[t + 1] & 1u) ? c2[0] : -INFINITY; sc[t + 1][1] = (valid[t + 1] & 2u) ? c2[1] : -INFINITY;
        sc[t + 1][2] = (valid[t + 1] & 4u) ? c3[0] : -INFINITY; sc[t + 1][3] = (valid[t + 1] & 8u) ? c3[1] : -INFINITY;
        mx = fmaxf(fmaxf(mx, fmaxf(sc[t][0], sc[t][1])), fmaxf(sc[t][2], sc[t][3]));
        mx = fmaxf(fmaxf(mx, fmaxf(sc[t + 1][0], sc[t + 1][1])), fmaxf(sc[t + 1][2], sc[t + 1][3]));
      }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (t * 16 < ctx) {
        float e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          e[j] = ex2f((sc[t][j] - mx) * p.scale_log2e);     // exp2(-inf) = 0 for masked / out-of-range keys
          sum += e[j];
        }
        pa[hh][t][0] = pack_bf16x2(e[0], e[1]);
        pa[hh][t][1] = pack_bf16x2(e[2], e[3]);
      }
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    inv[hh] = 1.f / sum;
    }
    // the scores are in registers (the shuffles above consumed them): K of the warp's next item may land in the buffer
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && cached > 0 && next < p.B) issue(sK, &tm_k, bar_k, next);
    // ---- O = P V ----------------------------------------------------------------------------------------
    if (cached > 0) mbar_wait(bar_v, phase);
    uint32_t packed[HW][(HD / 8 + 7) / 8];
#pragma unroll
    for (int hh = 0; hh < HW; ++hh) {
    float o[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (t * 16 < ctx) {
        const uint32_t a[4] = {pa[hh][t][0], pa[hh][t][0], pa[hh][t][1], pa[hh][t][1]};
#pragma unroll
        for (int n = 0; n < HD / 8; n += 2) {
          uint32_t b[4];
          ldmatrix_x4_trans(b, sV + L::at(t * 16 + v_r, hh * C + n + v_c, ctx16));
          mma_bf16_16816(o[n], a, b[0], b[1]);
          mma_bf16_16816(o[n + 1], a, b[2], b[3]);
        }
      }
    }
    // every row group holds the same output row: group g writes n-tiles g and g + 8 (4 bytes per lane, 16 contiguous per quad)
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
      if ((n & 7) == grp) packed[hh][n >> 3] = pack_bf16x2(o[n][0] * inv[hh], o[n][1] * inv[hh]);
    }
    // the context is in registers: V of the next item may land
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && cached > 0 && next < p.B) issue(sV, &tm_v, bar_v, next);
    __nv_bfloat16* orow = p.out + static_cast<size_t>(seq) * p.ld_out + head * HD;
#pragma unroll
    for (int hh = 0; hh < HW; ++hh)
#pragma unroll
      for (int n = 0; n < HD / 8; ++n)
        if ((n & 7) == grp) *reinterpret_cast<uint32_t*>(orow + hh * HD + n * 8 + quad * 2) = packed[hh][n >> 3];
  }
}

__global__ void kv_append_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int nseq, int q_len, int pos0, int d_model,
                                 __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache, int max_ctx);

template <int HD>
static int launch_decode_attn(SmallAttnParams p, int nseq, cudaStream_t st) {
  p.B = nseq * p.num_heads;
  const int ctx = p.q_pos0 + 1;
  const int ctx16 = (ctx + 15) & ~15;
  // contexts <= 128 keys: up to 110 KB of shared memory per CTA; up to 256 keys: one CTA per SM
  const int budget = ctx <= 128 ? 110 * 1024 : 220 * 1024;
  constexpr int HW = KvSmem<HD>::kHeads, RB = KvSmem<HD>::kRowBytes;
  const int pw = (2 * ctx16 * RB + RB + 1023) & ~1023;                   // per warp: K | V | q, 1024-byte aligned (swizzle atoms)
  // the TMA kernel needs 16-byte aligned caches / qkv rows; anything else takes the register kernel below
  if (ctx <= 256 && pw + 1024 <= budget && p.num_heads % HW == 0 && (p.d_model % 8) == 0 && (p.ld_q % 8) == 0 &&
      ((reinterpret_cast<uintptr_t>(p.k_cache) | reinterpret_cast<uintptr_t>(p.v_cache) | reinterpret_cast<uintptr_t>(p.q)) & 15) == 0) {
    // warps per CTA: whatever packs most warps into an SM's 227 KB (1 KB per CTA is reserved by the driver); ties -> larger CTAs
    int wpc = 1, best = 0;
    for (int w = 1; w <= 4 && w * pw + 1024 <= budget; ++w) {
      const int resident = (227 * 1024) / (w * pw + 2048) * w;
      if (resident >= best) { best = resident; wpc = w; }
    }
    const int cached = p.append_kv ? ctx - 1 : ctx;
    CUtensorMap tmk, tmv;
    const uint32_t box_rows = static_cast<uint32_t>(cached > 0 ? cached : 1);
    const uint64_t rows = static_cast<uint64_t>(nseq) * p.max_ctx;
    int rc;
    if (KvSmem<HD>::kSwizzled) {
      rc = make_tmap_bf16_2d(&tmk, p.k_cache, rows, p.d_model, p.d_model, box_rows, 64);
      if (!rc) rc = make_tmap_bf16_2d(&tmv, p.v_cache, rows, p.d_model, p.d_model, box_rows, 64);
    } else {
      rc = make_tmap_bf16_2d_plain(&tmk, p.k_cache, rows, p.d_model, p.d_model, box_rows, HW * HD);
      if (!rc) rc = make_tmap_bf16_2d_plain(&tmv, p.v_cache, rows, p.d_model, p.d_model, box_rows, HW * HD);
    }
    if (rc) return rc;
    static bool configured_dev[64] = {};
    bool& configured = configured_dev[device_slot()];
    if (!configured) {
      rc = check_cuda(cudaFuncSetAttribute(decode_attn_tma_kernel<HD, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024),
                      "cudaFuncSetAttribute(decode_attn_tma)");
      if (rc) return rc;
      rc = check_cuda(cudaFuncSetAttribute(decode_attn_tma_kernel<HD, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                      "cudaFuncSetAttribute(decode_attn_tma)");
      if (rc) return rc;
      configured = true;
    }
    const size_t smem = static_cast<size_t>(wpc) * pw + 1024;
    // persistent: as many CTAs as fit on the machine at once, every warp walks items warp, warp + W, ...
    p.B /= HW;                                                           // items = (sequence, group of HW heads)
    int grid = (p.B + wpc - 1) / wpc;
    const int resident_ctas = opsg_num_sms() * (best / wpc);
    if (grid > resident_ctas) grid = resident_ctas;
    if (ctx <= 128)
      launch_kernel(decode_attn_tma_kernel<HD, 8>, grid, wpc * 32, smem, st, tmk, tmv, p, wpc, pw);
    else
      launch_kernel(decode_attn_tma_kernel<HD, 16>, grid, wpc * 32, smem, st, tmk, tmv, p, wpc, pw);
    OPSG_CHECK_LAUNCH("decode_attn_tma_kernel");
    return OPSG_OK;
  }
  if (p.append_kv) {                         // register version reads the cache only: append with the copy kernel first
    const long long total = static_cast<long long>(nseq) * (p.d_model / 8);
    launch_kernel(kv_append_kernel, static_cast<int>((total + 255) / 256), 256, 0, st, p.q, p.ld_q, nseq, 1, p.q_pos0,
                  p.d_model, const_cast<__nv_bfloat16*>(p.k_cache), const_cast<__nv_bfloat16*>(p.v_cache), p.max_ctx);
    OPSG_CHECK_LAUNCH("kv_append_kernel");
  }
  launch_kernel(decode_attn_kernel<HD>, (p.B + 7) / 8, 256, 0, st, p);
  OPSG_CHECK_LAUNCH("decode_attn_kernel");
  return OPSG_OK;
}

__global__ void kv_append_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int nseq, int q_len, int pos0, int d_model,
                                 __nv_bfloat16* __restrict__ k_cache, __nv_bfloat16* __restrict__ v_cache, int max_ctx) {
  pdl_wait_then_trigger();
  const int vec = d_model / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(nseq) * q_len * vec) return;
  const int c = static_cast<int>(idx % vec);
  const long long r = idx / vec;
  const int t = static_cast<int>(r % q_len), s = static_cast<int>(r / q_len);
  const __nv_bfloat16* src = qkv + static_cast<size_t>(r) * ld_qkv + c * 8;
  const size_t dst = (static_cast<size_t>(s) * max_ctx + pos0 + t) * d_model + c * 8;
  *reinterpret_cast<uint4*>(k_cache + dst) = __ldg(reinterpret_cast<const uint4*>(src + d_model));
  *reinterpret_cast<uint4*>(v_cache + dst) = __ldg(reinterpret_cast<const uint4*>(src + 2 * d_model));
}

}  // namespace opsg

using namespace opsg;

int launch_llm_prefill_attn_tc(const opsg_bf16* q, int ld_q, const opsg_bf16* k_cache, const opsg_bf16* v_cache, int max_ctx,
                               const uint8_t* key_mask, int nseq, int q_len, int q_pos0, int num_heads, int head_dim, float scale,
                               opsg_bf16* out, int ld_out, cudaStream_t stream);
int launch_self_attn_pairs(const opsg_bf16* qkv, const opsg_bf16* shared_query_qkv, const int32_t* text_mask, int B, int n_query,
                           int T, int num_heads, int head_dim, int text_queries, opsg_bf16* ctx_out, cudaStream_t stream);

extern "C" int opsg_self_attn_small(const opsg_bf16* qkv, const opsg_bf16* shared_query_qkv, const int32_t* text_mask, int B,
                                    int n_query, int T, int num_heads, int head_dim, int text_queries, opsg_bf16* ctx_out,
                                    void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(qkv && ctx_out && (T == 0 || text_mask), "self_attn_small: null pointer");
  OPSG_CHECK_ARG(B > 0 && n_query > 0 && T >= 0 && num_heads > 0, "self_attn_small: bad shape");
  SmallAttnParams p{};
  p.mode = 0;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.qkv_shared = reinterpret_cast<const __nv_bfloat16*>(shared_query_qkv);
  OPSG_CHECK_ARG(!shared_query_qkv || ((uintptr_t)shared_query_qkv & 15) == 0, "self_attn_small: shared_query_qkv must be 16-byte aligned");
  p.text_mask = text_mask;
  p.B = B; p.n_query = n_query; p.T = T; p.text_queries = text_queries;
  p.out = reinterpret_cast<__nv_bfloat16*>(ctx_out);
  p.num_heads = num_heads; p.d_model = num_heads * head_dim; p.ld_out = p.d_model;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  // tcgen05 + TMA tiles (self_attn_pairs.cu: two pairs stacked per 128-row tile) for the head's shape; the warp-level
  // mma.sync kernels below remain the fallback for other shapes (head_dim != 64, more than 64 rows per pair)
  {
    rc = launch_self_attn_pairs(qkv, shared_query_qkv, text_mask, B, n_query, T, num_heads, head_dim, text_queries, ctx_out,
                                reinterpret_cast<cudaStream_t>(stream));
    if (rc != OPSG_E_UNSUPPORTED) return rc;
  }
  if (head_dim == kQfHD && n_query + T <= kQfNK && (((uintptr_t)qkv | (uintptr_t)ctx_out) & 15) == 0) {
    constexpr int smem = kQfStages * kQfStageBytes;
    static bool configured_dev[64] = {};
  bool& configured = configured_dev[device_slot()];
    if (!configured) {
      rc = check_cuda(cudaFuncSetAttribute(qformer_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                      "cudaFuncSetAttribute(qformer_self_attn)");
      if (rc) return rc;
      configured = true;
    }
    const int problems = B * num_heads;
    int grid = opsg_num_sms() * 4;                 // 4 resident CTAs per SM (55 KB of smem each)
    if (grid > problems) grid = problems;
    launch_kernel(qformer_self_attn_kernel, grid, 128, smem, reinterpret_cast<cudaStream_t>(stream), p, problems);
    OPSG_CHECK_LAUNCH("qformer_self_attn_kernel");
    return OPSG_OK;
  }
  return dispatch_hd(p, B, n_query + T, head_dim, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int opsg_llm_attn(const opsg_bf16* q, int ld_q, const opsg_bf16* k_cache, const opsg_bf16* v_cache, int max_ctx,
                             const uint8_t* key_mask, int nseq, int q_len, int q_pos0, int num_heads, int head_dim,
                             float scale, opsg_bf16* out, int ld_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(q && k_cache && v_cache && key_mask && out, "llm_attn: null pointer");
  OPSG_CHECK_ARG(nseq > 0 && q_len > 0 && q_pos0 >= 0 && q_pos0 + q_len <= max_ctx, "llm_attn: bad shape");
  OPSG_CHECK_ARG(ld_q % 8 == 0 && ld_out % 2 == 0, "llm_attn: bad leading dims");
  SmallAttnParams p{};
  p.mode = 1;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k_cache = reinterpret_cast<const __nv_bfloat16*>(k_cache);
  p.v_cache = reinterpret_cast<const __nv_bfloat16*>(v_cache);
  p.key_mask = key_mask;
  p.ld_q = ld_q; p.max_ctx = max_ctx; p.q_len = q_len; p.q_pos0 = q_pos0;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out; p.num_heads = num_heads; p.d_model = num_heads * head_dim;
  p.scale_log2e = 1.4426950408889634f * scale;
  if (q_len == 1 && q_pos0 + 1 <= 256 && (head_dim == 64 || head_dim == 80 || head_dim == 128) && (ld_q % 8) == 0 &&
      (ld_out % 8) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)q & 15) == 0) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);     // decode step: GEMV-style kernel, one warp per (seq, head)
    if (head_dim == 64) return launch_decode_attn<64>(p, nseq, st);
    if (head_dim == 80) return launch_decode_attn<80>(p, nseq, st);
    return launch_decode_attn<128>(p, nseq, st);
  }
  // prefill of prompts up to 64 tokens: tcgen05 + TMA tiles, two sequences stacked per 128-row tile (llm_prefill_attn.cu)
  if (q_len > 1) {
    rc = launch_llm_prefill_attn_tc(q, ld_q, k_cache, v_cache, max_ctx, key_mask, nseq, q_len, q_pos0, num_heads, head_dim, scale, out,
                                    ld_out, reinterpret_cast<cudaStream_t>(stream));
    if (rc != OPSG_E_UNSUPPORTED) return rc;
  }
  return dispatch_hd(p, nseq, q_pos0 + q_len, head_dim, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int opsg_llm_attn_append(const opsg_bf16* qkv, int ld_qkv, opsg_bf16* k_cache, opsg_bf16* v_cache, int max_ctx,
                                    const uint8_t* key_mask, int nseq, int q_pos0, int num_heads, int head_dim, float scale,
                                    opsg_bf16* out, int ld_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(qkv && k_cache && v_cache && key_mask && out, "llm_attn_append: null pointer");
  OPSG_CHECK_ARG(nseq > 0 && q_pos0 >= 0 && q_pos0 < max_ctx && q_pos0 + 1 <= 256, "llm_attn_append: bad shape");
  OPSG_CHECK_ARG(head_dim == 64 || head_dim == 80 || head_dim == 128, "llm_attn_append: head_dim %d unsupported", head_dim);
  OPSG_CHECK_ARG(ld_qkv >= 3 * num_heads * head_dim && (ld_qkv % 8) == 0 && (ld_out % 8) == 0 &&
                 (((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)k_cache | (uintptr_t)v_cache) & 15) == 0,
                 "llm_attn_append: bad layout");
  SmallAttnParams p{};
  p.mode = 1;
  p.q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.k_cache = reinterpret_cast<const __nv_bfloat16*>(k_cache);
  p.v_cache = reinterpret_cast<const __nv_bfloat16*>(v_cache);
  p.key_mask = key_mask;
  p.ld_q = ld_qkv; p.max_ctx = max_ctx; p.q_len = 1; p.q_pos0 = q_pos0; p.append_kv = 1;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out; p.num_heads = num_heads; p.d_model = num_heads * head_dim;
  p.scale_log2e = 1.4426950408889634f * scale;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (head_dim == 64) return launch_decode_attn<64>(p, nseq, st);
  if (head_dim == 80) return launch_decode_attn<80>(p, nseq, st);
  return launch_decode_attn<128>(p, nseq, st);
}

extern "C" int opsg_kv_append(const opsg_bf16* qkv, int ld_qkv, int nseq, int q_len, int pos0, int d_model, opsg_bf16* k_cache,
                              opsg_bf16* v_cache, int max_ctx, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(qkv && k_cache && v_cache, "kv_append: null pointer");
  OPSG_CHECK_ARG(nseq > 0 && q_len > 0 && pos0 >= 0 && pos0 + q_len <= max_ctx && d_model % 8 == 0 && ld_qkv % 8 == 0,
                 "kv_append: bad shape");
  const long long total = static_cast<long long>(nseq) * q_len * (d_model / 8);
  launch_kernel(kv_append_kernel, static_cast<int>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(qkv), ld_qkv, nseq, q_len, pos0, d_model,
      reinterpret_cast<__nv_bfloat16*>(k_cache), reinterpret_cast<__nv_bfloat16*>(v_cache), max_ctx);
  OPSG_CHECK_LAUNCH("kv_append_kernel");
  return OPSG_OK;
}
