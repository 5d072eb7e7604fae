// K5 (v1, kept as the A/B baseline of round 1: select with OPSG_XATTN_IMPL=1) — pair-query x image-feature masked
// cross-attention on tcgen05 tensor cores.  Superseded by xattn_pairs.cu (v2).
//
// All pairs' query rows are stacked along M (row = pair * n_query + r); K and V are projected once per image
// and shared by every pair, so per head the whole thing is  softmax(Q[M x 64] . K^T[64 x L] + mask(pair)) . V.
// Work unit = (128-row tile, head); a persistent CTA walks a contiguous, head-major range of units so the
// head's K [L x 64] and V^T [64 x L] stay resident in shared memory.
//   warp 0 (1 thread)  TMA producer : K / V^T per head change, Q tile per unit (2 stages)
//   warp 1 (1 thread)  MMA issuer   : S = Q.K^T (128 x 256 x 64, 4 MMAs) -> TMEM cols [0,256)
//                                     O = P.V   (128 x 64 x 256, 16 MMAs) -> TMEM cols [256,320)
//   warp 2             TMEM allocator
//   warps 4-7          softmax + epilogue: thread = score row; the pair's 256-bit key mask is
//                      bits[i] | bits[j] (never materialised in HBM); P goes to smem as the bf16 A operand in
//                      the 128B-swizzled K-major layout; O is normalised by 1/rowsum and stored as bf16.
// Mask semantics follow HF's `(1 - m) * finfo.min` additive bias: masked keys get weight exactly 0 and a pair
// whose union mask is empty attends uniformly to all L keys.
#include "common.cuh"
#include "host_util.h"

namespace opsg {
namespace xa1 {

constexpr int kXaThreads = 256;
constexpr int kXaKeys = 256;       // max keys (one N=256 MMA)
constexpr int kXaHd = 64;

struct XattnParams {
  const uint32_t* bits;
  const int32_t* pair_index;
  __nv_bfloat16* out;
  int words, num_objects, n_query, L, num_heads, d_model;
  int rows;          // B * n_query
  int m_tiles;
  int total_units;
  float scale_log2e;
};

struct XattnSmem {
  static constexpr int kK = kXaKeys * 128;        // 32768
  static constexpr int kVt = 4 * 64 * 128;        // 32768
  static constexpr int kQ = 128 * 128;            // 16384 per stage
  static constexpr int kP = 4 * 128 * 128;        // 65536
  static constexpr int kOffK = 0;
  static constexpr int kOffVt = kOffK + kK;
  static constexpr int kOffQ = kOffVt + kVt;
  static constexpr int kOffP = kOffQ + 2 * kQ;
  static constexpr int kOffBar = kOffP + kP;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kXaThreads, 1)
xattn_pairs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmVt, const XattnParams p) {
  pdl_wait_then_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + XattnSmem::kOffK;
  uint8_t* sVt = smem + XattnSmem::kOffVt;
  uint8_t* sQ = smem + XattnSmem::kOffQ;
  uint8_t* sP = smem + XattnSmem::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XattnSmem::kOffBar);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVt);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(&q_empty[0], 1); mbar_init(&q_empty[1], 1);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 256;

  // contiguous, head-major unit range of this CTA
  const int per = (p.total_units + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per;
  const int u_end = min(p.total_units, u_begin + per);
  const int n_units = max(0, u_end - u_begin);

  if (threadIdx.x == 0) {
    // ===================== TMA producer =====================
    int cur_head = -1, kv_loads = 0;
    for (int i = 0; i < n_units; ++i) {
      const int u = u_begin + i;
      const int head = u / p.m_tiles, mt = u % p.m_tiles;
      if (head != cur_head) {
        if (kv_loads > 0) mbar_wait(kv_empty, (kv_loads - 1) & 1);
        mbar_expect_tx(kv_full, XattnSmem::kK + XattnSmem::kVt);
        tma_load_2d(sK, &tmK, kv_full, head * kXaHd, 0);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sVt + kb * 8192, &tmVt, kv_full, kb * 64, head * kXaHd);
        cur_head = head;
        ++kv_loads;
      }
      const int st = i & 1;
      mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
      mbar_expect_tx(&q_full[st], XattnSmem::kQ);
      tma_load_2d(sQ + st * XattnSmem::kQ, &tmQ, &q_full[st], head * kXaHd, mt * 128);
    }
  } else if (threadIdx.x == 32) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = umma_idesc_bf16(128, kXaKeys);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kXaHd);
    int kv_waits = 0;
    int prev_head = -1;
    auto issue_qk = [&](int j) {
      const int head = (u_begin + j) / p.m_tiles;
      if (head != prev_head) {
        mbar_wait(kv_full, kv_waits & 1);
        ++kv_waits;
        prev_head = head;
      }
      const int st = j & 1;
      mbar_wait(&q_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t a = smem_u32(sQ + st * XattnSmem::kQ);
      const uint32_t b = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < kXaHd / 16; ++k)
        umma_ss(tmem_S, umma_desc_k_sw128(a + k * 32), umma_desc_k_sw128(b + k * 32), idesc_qk, k > 0 ? 1u : 0u);
      tc_commit(&q_empty[st]);
      tc_commit(s_full);
    };
    if (n_units > 0) issue_qk(0);
    for (int i = 0; i < n_units; ++i) {
      mbar_wait(p_ready, i & 1);      // P(i) in smem, S(i) fully read, O(i-1) fully read
      tc_fence_after();
      const uint32_t a = smem_u32(sP);
      const uint32_t b = smem_u32(sVt);
#pragma unroll
      for (int k = 0; k < kXaKeys / 16; ++k) {
        const uint32_t blk = k >> 2, sub = k & 3;
        umma_ss(tmem_O, umma_desc_k_sw128(a + blk * 16384 + sub * 32), umma_desc_k_sw128(b + blk * 8192 + sub * 32),
                idesc_pv, k > 0 ? 1u : 0u);
      }
      tc_commit(o_full);
      const int head = (u_begin + i) / p.m_tiles;
      const bool last = (i + 1 == n_units);
      if (!last && (u_begin + i + 1) / p.m_tiles != head) tc_commit(kv_empty);
      if (!last) issue_qk(i + 1);
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                         // row inside the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int full_words = p.L >> 5;
    const uint32_t tail_mask = (p.L & 31) ? ((1u << (p.L & 31)) - 1u) : 0u;
    for (int i = 0; i < n_units; ++i) {
      const int u = u_begin + i;
      const int head = u / p.m_tiles, mt = u % p.m_tiles;
      const int row = mt * 128 + r;
      const bool valid = row < p.rows;
      // pair mask = bits[i] | bits[j], restricted to the L real keys
      uint32_t m[8];
      bool empty = true;
      {
        int oi = 0, oj = 0;
        if (valid) {
          const int pair = row / p.n_query;
          const int pidx = p.pair_index ? p.pair_index[pair] : pair;
          oi = pidx / p.num_objects;
          oj = pidx % p.num_objects;
        }
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          uint32_t keyok = (w < full_words) ? 0xffffffffu : (w == full_words ? tail_mask : 0u);
          uint32_t v = 0;
          if (valid && w < p.words)
            v = __ldg(p.bits + static_cast<size_t>(oi) * p.words + w) | __ldg(p.bits + static_cast<size_t>(oj) * p.words + w);
          m[w] = v & keyok;
          empty = empty && (m[w] == 0);
        }
        if (empty) {   // finfo.min on every key -> uniform attention over the L real keys
#pragma unroll
          for (int w = 0; w < 8; ++w) m[w] = (w < full_words) ? 0xffffffffu : (w == full_words ? tail_mask : 0u);
        }
      }
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      // pass 1: row max over unmasked keys
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_ld_wait();
        const uint32_t mw = m[c];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((mw >> j) & 1u) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      if (empty || mx == -INFINITY) mx = 0.f;
      const float mxs = mx * p.scale_log2e;
      // pass 2: p = exp2(s*scale - max*scale), row sum, bf16 P -> swizzled smem
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off + c * 32, v);
        tmem_ld_wait();
        const uint32_t mw = m[c];
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = empty ? 0.f : __uint_as_float(v[j]);
          const float e = ex2_approx(fmaf(s, p.scale_log2e, -mxs));
          pv[j] = ((mw >> j) & 1u) ? e : 0.f;
          sum += pv[j];
        }
        uint8_t* blk = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c & 1) * 4 + g) ^ (r & 7);
          uint4 u4;
          u4.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]);
          u4.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
          u4.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]);
          u4.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
          *reinterpret_cast<uint4*>(blk + chunk * 16) = u4;
        }
      }
      const float inv = 1.f / sum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      // epilogue: O / rowsum -> bf16 ctx[row, head*64 : head*64+64]
      mbar_wait(o_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_off + c * 32, v);
        tmem_ld_wait();
        if (valid) {
          __nv_bfloat16* dst = p.out + static_cast<size_t>(row) * p.d_model + head * kXaHd + c * 32;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u4;
            u4.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
            u4.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
            u4.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
            u4.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[g] = u4;
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace xa1
}  // namespace opsg

using namespace opsg;
using namespace opsg::xa1;

extern "C" int opsg_xattn_pairs_v1(const opsg_bf16* q, const opsg_bf16* k, int ld_k, const opsg_bf16* vt, int ld_vt,
                                const uint32_t* bits, int words, const int32_t* pair_index, int num_objects, int B,
                                int n_query, int L, int num_heads, int head_dim, opsg_bf16* ctx_out, void* stream) {
  int rc = opsg_device_check();
  if (rc) return rc;
  OPSG_CHECK_ARG(q && k && vt && bits && ctx_out, "xattn_pairs: null pointer");
  OPSG_CHECK_ARG(B > 0 && n_query > 0 && L > 0 && num_heads > 0 && num_objects > 0, "xattn_pairs: bad shape");
  if (head_dim != kXaHd) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: head_dim %d unsupported (64 only)", head_dim);
  if (L > kXaKeys) return set_error(OPSG_E_UNSUPPORTED, "xattn_pairs: L=%d image tokens > %d unsupported", L, kXaKeys);
  OPSG_CHECK_ARG(words >= (L + 31) / 32 && words <= 8, "xattn_pairs: words=%d inconsistent with L=%d", words, L);
  const int d_model = num_heads * head_dim;
  OPSG_CHECK_ARG(ld_k >= d_model && ld_k % 8 == 0 && ld_vt >= L && ld_vt % 8 == 0, "xattn_pairs: bad leading dims");
  const int rows = B * n_query;
  CUtensorMap tmQ, tmK, tmVt;
  rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)rows, (uint64_t)d_model, (uint64_t)d_model, 128, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)L, (uint64_t)d_model, (uint64_t)ld_k, kXaKeys, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmVt, vt, (uint64_t)d_model, (uint64_t)L, (uint64_t)ld_vt, 64, 64);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    rc = check_cuda(cudaFuncSetAttribute(xattn_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XattnSmem::kTotal),
                    "cudaFuncSetAttribute(xattn)");
    if (rc) return rc;
    configured = true;
  }
  XattnParams p;
  p.bits = bits; p.pair_index = pair_index; p.out = reinterpret_cast<__nv_bfloat16*>(ctx_out);
  p.words = words; p.num_objects = num_objects; p.n_query = n_query; p.L = L; p.num_heads = num_heads; p.d_model = d_model;
  p.rows = rows; p.m_tiles = (rows + 127) / 128; p.total_units = p.m_tiles * num_heads;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  const int grid = p.total_units < opsg_num_sms() ? p.total_units : opsg_num_sms();
  launch_kernel(xattn_pairs_kernel, grid, kXaThreads, XattnSmem::kTotal, reinterpret_cast<cudaStream_t>(stream), tmQ, tmK, tmVt, p);
  OPSG_CHECK_LAUNCH("xattn_pairs_kernel");
  return OPSG_OK;
}
